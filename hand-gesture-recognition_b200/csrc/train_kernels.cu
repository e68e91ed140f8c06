// Element-wise and reduction kernels of the training step:
// train-mode BatchNorm (reference model/gelan.py:46,56 with nn.BatchNorm2d in training mode: batch statistics,
// running-stat update with momentum 0.1 and the unbiased variance), its backward, SiLU / GELU backward,
// LayerNorm backward, bias gradients, the loss of train.py:63-75 (0.001 * CE + JointsMSE, libs/loss.py:4-40),
// AdamW (train.py:50-51) and the fp32 -> bf16 weight re-layouts the tensor-core kernels consume.
//
// Every reduction is two-stage and ordered (per-CTA partial sums, then a fixed-order final sum), so the
// gradients are bitwise reproducible from run to run (the reference trains with deterministic=True,
// train.py:232).  All kernels are memory-bound; they use 128-bit loads over the channel-contiguous NHWC layout.
#include "hgr_internal.h"
#include "ptx.cuh"
#include "train.h"

namespace hgr {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = bf16_lo(w[i]);
    v[2 * i + 1] = bf16_hi(w[i]);
  }
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// d/du [u * sigmoid(u)]
__device__ __forceinline__ float silu_grad(float u) {
  const float s = sigmoid_f(u);
  return s * (1.0f + u * (1.0f - s));
}

// ---------------------------------------------------------------------------------------------
// Column reductions over a [rows][C] channel-contiguous matrix.  A CTA owns a contiguous row range; thread
// (cg, prow) walks rows prow, prow + nrow, ... of channel group cg (8 channels = one 16-byte load) and keeps
// K running sums per channel; the CTA then folds its threads through shared memory and writes
// partial[blockIdx][k][c].  `Body` supplies the per-element contributions.
// ---------------------------------------------------------------------------------------------
template <int K, typename Body>
__device__ __forceinline__ void column_partials(long long rows, int C, float* __restrict__ partial, Body body) {
  extern __shared__ float red[];  // [kThreads][K * 8]
  const int ncg = C >> 3;
  const int nrow = kThreads / ncg;
  const int cg = threadIdx.x % ncg, prow = threadIdx.x / ncg;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per;
  const long long r1 = r0 + per < rows ? r0 + per : rows;
  float acc[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  if (prow < nrow) {
#pragma unroll 4
    for (long long r = r0 + prow; r < r1; r += nrow) body(r, cg * 8, acc);
  }
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x * (K * 8) + k * 8 + e] = prow < nrow ? acc[k][e] : 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < K * C; idx += kThreads) {
    const int k = idx / C, c = idx % C;
    float s = 0.f;
    for (int r = 0; r < nrow; ++r) s += red[(r * ncg + (c >> 3)) * (K * 8) + k * 8 + (c & 7)];
    partial[((size_t)blockIdx.x * K + k) * C + c] = s;
  }
}

// Fixed-order sum of column c over the per-CTA partials by ONE CTA of 128 threads: thread t adds partials t, t + 128,
// ... (ascending, double), the lane sums are folded by a butterfly and the four warp sums in ascending order, so the
// order depends only on nblk and the result is reproducible.  True in thread 0, which holds the sums.  (A warp per
// channel and eight channels per CTA ran the 64-channel layers' 592 partials on 8 CTAs: 10 us per launch.)
constexpr int kFinalThreads = 128;
template <int K>
__device__ __forceinline__ bool block_sum_partials(const float* __restrict__ partial, int nblk, int C, int c,
                                                   double (&out)[K]) {
  __shared__ double fold[kFinalThreads / 32][K];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = 0.0;
  for (int b = threadIdx.x; b < nblk; b += kFinalThreads) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] += (double)partial[((size_t)b * K + k) * C + c];
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) out[k] += __shfl_xor_sync(0xffffffffu, out[k], o);
    if (lane == 0) fold[w][k] = out[k];
  }
  __syncthreads();
  if (threadIdx.x != 0) return false;
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = ((fold[0][k] + fold[1][k]) + fold[2][k]) + fold[3][k];
  return true;
}

int reduce_blocks(long long rows) {
  // at least 32 rows per CTA: the 12 x 12 maps of a 32-crop batch have 4608 rows, and at 128 rows per CTA their
  // reductions ran on 36 CTAs (bn_bwd_partial 28-52 us, as long as on the 96 x 96 maps)
  long long b = (rows + 31) / 32;
  if (b > 592) b = 592;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ BatchNorm forward (train) ----

__global__ void __launch_bounds__(kThreads)
bn_stats_partial_kernel(const __nv_bfloat16* __restrict__ z, long long rows, int C, float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  column_partials<2>(rows, C, partial, [&](long long r, int c0, float (&acc)[2][8]) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(z + r * C + c0)), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += v[e];
      acc[1][e] = fmaf(v[e], v[e], acc[1][e]);
    }
  });
}

// mean / biased variance of the batch -> scale = gamma * rstd, shift = beta - mean * scale, x-hat parameters
// (mean, rstd) for the backward pass, and the running statistics (unbiased variance, momentum).
__global__ void __launch_bounds__(kFinalThreads)
bn_stats_final_kernel(const float* __restrict__ partial, int nblk, int C, long long rows,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ scale,
                      float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                      float* __restrict__ running_mean, float* __restrict__ running_var, float momentum) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;  // one CTA per channel
  double s[2];
  if (!block_sum_partials<2>(partial, nblk, C, c, s)) return;
  const double n = (double)rows;
  const double mean = s[0] / n;
  double var = s[1] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + 1e-5));
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  mean_out[c] = (float)mean;
  rstd_out[c] = rstd;
  if (running_mean != nullptr) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// y = act(scale * z + shift (+ res)), written into a channel slice of the consumer's buffer
template <bool SILU, bool RES>
__global__ void __launch_bounds__(kThreads)
bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ z, long long rows, int C, const float* __restrict__ scale,
                  const float* __restrict__ shift, const __nv_bfloat16* __restrict__ res, int res_ctot,
                  __nv_bfloat16* __restrict__ y, int y_ctot, int cg_log2) {
  pdl_launch_dependents();
  pdl_wait();
  // channel groups per row are a power of two (C = 64 .. 512): row / group come from a shift and a mask
  const long long total = rows << cg_log2;
  const int cg_mask = (1 << cg_log2) - 1;
#pragma unroll 2
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long r = i >> cg_log2;
    const int c0 = ((int)i & cg_mask) * 8;
    float v[8], rr[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(z + r * C + c0)), v);
    if constexpr (RES) unpack8(__ldg(reinterpret_cast<const uint4*>(res + r * res_ctot + c0)), rr);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float u = fmaf(v[e], __ldg(scale + c0 + e), __ldg(shift + c0 + e));
      if constexpr (RES) u += rr[e];
      v[e] = SILU ? u * sigmoid_f(u) : u;
    }
    *reinterpret_cast<uint4*>(y + r * y_ctot + c0) = pack8(v);
  }
}

// ------------------------------------------------------------------ BatchNorm backward ----
// du = dy * act'(u), u = scale * z + shift (+ res);  x-hat = (z - mean) * rstd
// partial sums: S1 = sum du (= d beta), S2 = sum du * x-hat (= d gamma)
template <bool SILU, bool RES>
__global__ void __launch_bounds__(kThreads)
bn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ctot, const __nv_bfloat16* __restrict__ z,
                      long long rows, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ mean, const float* __restrict__ rstd,
                      const __nv_bfloat16* __restrict__ res, int res_ctot, float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  column_partials<2>(rows, C, partial, [&](long long r, int c0, float (&acc)[2][8]) {
    float g[8], v[8], rr[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + r * dy_ctot + c0)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(z + r * C + c0)), v);
    if constexpr (RES) unpack8(__ldg(reinterpret_cast<const uint4*>(res + r * res_ctot + c0)), rr);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float du = g[e];
      if constexpr (SILU) {
        float u = fmaf(v[e], __ldg(scale + c0 + e), __ldg(shift + c0 + e));
        if constexpr (RES) u += rr[e];
        du *= silu_grad(u);
      }
      const float xh = (v[e] - __ldg(mean + c0 + e)) * __ldg(rstd + c0 + e);
      acc[0][e] += du;
      acc[1][e] = fmaf(du, xh, acc[1][e]);
    }
  });
}

// d beta = S1, d gamma = S2; c1 = S1 / n, c2 = S2 / n for the apply pass
__global__ void __launch_bounds__(kFinalThreads)
bn_bwd_final_kernel(const float* __restrict__ partial, int nblk, int C, long long rows, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float* __restrict__ c1, float* __restrict__ c2) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;
  double s[2];
  if (!block_sum_partials<2>(partial, nblk, C, c, s)) return;
  dbeta[c] = (float)s[0];
  dgamma[c] = (float)s[1];
  c1[c] = (float)(s[0] / (double)rows);
  c2[c] = (float)(s[1] / (double)rows);
}

// dz = scale * (du - c1 - x-hat * c2)   [scale = gamma * rstd];   residual branch: dres += du
template <bool SILU, bool RES>
__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ctot, const __nv_bfloat16* __restrict__ z,
                    long long rows, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ c1,
                    const float* __restrict__ c2, const __nv_bfloat16* __restrict__ res, int res_ctot,
                    __nv_bfloat16* __restrict__ dres, int dres_ctot, __nv_bfloat16* __restrict__ dz, int cg_log2) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = rows << cg_log2;
  const int cg_mask = (1 << cg_log2) - 1;
#pragma unroll 2
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const long long r = i >> cg_log2;
    const int c0 = ((int)i & cg_mask) * 8;
    float g[8], v[8], rr[8], dr[8], out[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + r * dy_ctot + c0)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(z + r * C + c0)), v);
    if constexpr (RES) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(res + r * res_ctot + c0)), rr);
      unpack8(*reinterpret_cast<const uint4*>(dres + r * dres_ctot + c0), dr);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sc = __ldg(scale + c0 + e);
      float du = g[e];
      if constexpr (SILU) {
        float u = fmaf(v[e], sc, __ldg(shift + c0 + e));
        if constexpr (RES) u += rr[e];
        du *= silu_grad(u);
      }
      if constexpr (RES) dr[e] += du;
      const float xh = (v[e] - __ldg(mean + c0 + e)) * __ldg(rstd + c0 + e);
      out[e] = sc * (du - __ldg(c1 + c0 + e) - xh * __ldg(c2 + c0 + e));
    }
    *reinterpret_cast<uint4*>(dz + r * C + c0) = pack8(out);
    if constexpr (RES) *reinterpret_cast<uint4*>(dres + r * dres_ctot + c0) = pack8(dr);
  }
}

// ------------------------------------------------------------------ GELU ----
__global__ void __launch_bounds__(kThreads) gelu_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            __nv_bfloat16* __restrict__ y, long long n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kThreads) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = gelu_erf_f(v[e]);
    reinterpret_cast<uint4*>(y)[i] = pack8(v);
  }
}

// dpre = dh * (Phi(x) + x * phi(x))   (exact-erf GELU, transformer.py:35), in place on dh
__global__ void __launch_bounds__(kThreads) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ pre,
                                                            __nv_bfloat16* __restrict__ dh, long long n8) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n8; i += (long long)gridDim.x * kThreads) {
    float v[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(pre) + i), v);
    unpack8(reinterpret_cast<const uint4*>(dh)[i], g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float cdf = 0.5f * (1.0f + erff(v[e] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * v[e] * v[e]);
      g[e] *= fmaf(v[e], pdf, cdf);
    }
    reinterpret_cast<uint4*>(dh)[i] = pack8(g);
  }
}

// ------------------------------------------------------------------ bias gradients ----
__global__ void __launch_bounds__(kThreads)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ g, long long rows, int C, float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  column_partials<1>(rows, C, partial, [&](long long r, int c0, float (&acc)[1][8]) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(g + r * C + c0)), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] += v[e];
  });
}

// out_k[c] = sum_b partial[b][k][c]  (k < K <= 2; dst1 may be null)
__global__ void __launch_bounds__(kFinalThreads)
sums_final_kernel(const float* __restrict__ partial, int nblk, int K, int C, float* __restrict__ dst0,
                  float* __restrict__ dst1) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;
  if (K == 1) {
    double s[1];
    if (block_sum_partials<1>(partial, nblk, C, c, s) && dst0 != nullptr) dst0[c] = (float)s[0];
  } else {
    double s[2];
    if (block_sum_partials<2>(partial, nblk, C, c, s)) {
      if (dst0 != nullptr) dst0[c] = (float)s[0];
      if (dst1 != nullptr) dst1[c] = (float)s[1];
    }
  }
}

// ------------------------------------------------------------------ LayerNorm backward ----
// One warp per 256-wide row.  y = gamma * x-hat + beta;  dx = rstd * (dy*gamma - mean(dy*gamma) - x-hat * mean(dy*gamma*x-hat)).
// g_out = g_in + dx (the residual stream's gradient).  Per-CTA partial sums of d gamma = sum dy * x-hat and
// d beta = sum dy go to partial[blockIdx][2][256].
__global__ void __launch_bounds__(kThreads)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ g_in, __nv_bfloat16* __restrict__ g_out, long long rows,
              float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][2][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per;
  const long long r1 = r0 + per < rows ? r0 + per : rows;
  float gm[8];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * lane + 1);
    gm[0] = a.x; gm[1] = a.y; gm[2] = a.z; gm[3] = a.w; gm[4] = b.x; gm[5] = b.y; gm[6] = b.z; gm[7] = b.w;
  }
  float dgam[8], dbet[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) dgam[e] = dbet[e] = 0.f;
  for (long long r = r0 + warp; r < r1; r += 8) {
    float xv[8], dv[8], gi[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * 256) + lane), xv);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + r * 256) + lane), dv);
    if (g_in != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(g_in + r * 256) + lane), gi);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += xv[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 256);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      xv[e] -= mean;
      sq = fmaf(xv[e], xv[e], sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / 256) + 1e-5f);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      xv[e] *= rstd;  // x-hat
      dgam[e] = fmaf(dv[e], xv[e], dgam[e]);
      dbet[e] += dv[e];
      dv[e] *= gm[e];  // dy * gamma
      m1 += dv[e];
      m2 = fmaf(dv[e], xv[e], m2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= (1.0f / 256);
    m2 *= (1.0f / 256);
    float out[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      out[e] = rstd * (dv[e] - m1 - xv[e] * m2);
      if (g_in != nullptr) out[e] += gi[e];
    }
    reinterpret_cast<uint4*>(g_out + r * 256)[lane] = pack8(out);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[warp][0][lane * 8 + e] = dgam[e];
    red[warp][1][lane * 8 + e] = dbet[e];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 512; idx += kThreads) {
    const int k = idx >> 8, c = idx & 255;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][k][c];
    partial[((size_t)blockIdx.x * 2 + k) * 256 + c] = s;
  }
}

// ------------------------------------------------------------------ token assembly backward ----
// g: (B, T, 256) gradient of the token stream entering layer 0.  d cls_token = sum_b g[b, 0, :];
// dfeat[b * P + p, :] = g[b, 1 + p, :] (the position table is not a parameter).
__global__ void __launch_bounds__(kThreads)
token_bwd_kernel(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dcls, int B,
                 int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int P = T - 1;
  if (blockIdx.x == gridDim.x - 1) {
    const int c = threadIdx.x;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += __bfloat162float(g[(size_t)b * T * 256 + c]);
    dcls[c] = s;
    return;
  }
  const long long total = (long long)B * P * 32;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total;
       i += (long long)(gridDim.x - 1) * kThreads) {
    const long long row = i >> 5;
    const int ch = (int)(i & 31);
    const long long b = row / P, p = row % P;
    reinterpret_cast<uint4*>(dfeat)[i] = __ldg(reinterpret_cast<const uint4*>(g + ((size_t)b * T + 1 + p) * 256) + ch);
  }
}

// ------------------------------------------------------------------ weight re-layouts ----
// fp32 PyTorch layouts -> the bf16 K-major layouts of the implicit-GEMM kernels, ALL matrices of the network
// in one launch (blockIdx.y = job; see PackJob in train.h):
//  mode 0: conv forward   dst[co][kh][kw][ci]              = w[co][ci][kh][kw]
//  mode 1: conv dgrad s1  dst[ci][kh][kw][co]              = w[co][ci][k-1-kh][k-1-kw]
//  mode 2: conv dgrad s2  dst[ci][tap][co], tap of parity (ph, pw) in build_dgrad_s2_op's order
//  mode 3: conv1 forward  dst[co][32], k = (kh*3+kw)*3 + ci, zero for k >= 27
//  mode 4: matrix copy    dst[r][c] = w[r][c]        (Co = rows, Ci = cols)
//  mode 5: matrix transpose dst[c][r] = w[r][c]
__global__ void __launch_bounds__(kThreads)
pack_jobs_kernel(const PackJob* __restrict__ jobs, const float* __restrict__ params) {
  pdl_launch_dependents();
  pdl_wait();
  const PackJob jb = jobs[blockIdx.y];
  const float* __restrict__ w = params + jb.src_off;
  __nv_bfloat16* __restrict__ dst = jb.dst;
  const int Co = jb.Co, Ci = jb.Ci, mode = jb.mode, ph = jb.ph, pw = jb.pw;
  const int kk = jb.k * jb.k;
  const int nh = ph ? 2 : 1, nw = pw ? 2 : 1;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < jb.total; i += (long long)gridDim.x * kThreads) {
    float v = 0.f;
    if (mode == 0) {
      const int ci = (int)(i % Ci);
      const int t = (int)((i / Ci) % kk);
      const int co = (int)(i / ((long long)Ci * kk));
      v = w[((size_t)co * Ci + ci) * kk + t];
    } else if (mode == 1) {
      const int co = (int)(i % Co);
      const int t = (int)((i / Co) % kk);
      const int ci = (int)(i / ((long long)Co * kk));
      v = w[((size_t)co * Ci + ci) * kk + (kk - 1 - t)];
    } else if (mode == 2) {
      const int co = (int)(i % Co);
      const int t = (int)((i / Co) % (nh * nw));
      const int ci = (int)(i / ((long long)Co * nh * nw));
      const int a = t / nw, b = t % nw;
      const int kh = ph ? (a == 0 ? 0 : 2) : 1;
      const int kw = pw ? (b == 0 ? 0 : 2) : 1;
      v = w[((size_t)co * Ci + ci) * 9 + kh * 3 + kw];
    } else if (mode == 3) {
      const int kidx = (int)(i % 32), co = (int)(i / 32);
      if (kidx < 27) {
        const int ci = kidx % 3, t = kidx / 3;
        v = w[((size_t)co * 3 + ci) * 9 + t];
      }
    } else if (mode == 4) {
      v = w[i];
    } else {
      const int r = (int)(i % Co), c = (int)(i / Co);  // dst index i = c * Co + r, src [Co][Ci]
      v = w[(size_t)r * Ci + c];
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------ loss ----
// train.py:63-64: total = cls_weight * CE(logits, label) + JointsMSE(heat, target, weight)
// CE: mean over the batch of -log softmax(logits)[label]  (libs/loss.py:33-40)
// JointsMSE: (1/J) sum_j 0.5 * mean_{b,hw} (w[b,j] * (p - g))^2  (libs/loss.py:10-30)
// One CTA per image row of logits for CE (tiny) and a grid-stride pass over the heatmaps whose per-CTA
// partial sums are reduced in a fixed order.
__global__ void __launch_bounds__(kThreads)
loss_heat_kernel(const float* __restrict__ heat, const float* __restrict__ target, const float* __restrict__ weight,
                 long long n, int hw, float inv_norm, float* __restrict__ dheat, float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[kThreads / 32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float w = __ldg(weight + i / hw);
    const float d = w * (heat[i] - target[i]);
    s = fmaf(d, d, s);
    if (dheat != nullptr) dheat[i] = w * d * inv_norm;  // d/dp [0.5 (w (p-g))^2] * inv_norm
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// single CTA: CE over the batch + final sums.  out = {total, class_loss (weighted), joints_loss}
__global__ void __launch_bounds__(kThreads)
loss_final_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int C, float cls_weight,
                  float* __restrict__ dlogits, const float* __restrict__ partial, int nblk, float inv_norm,
                  float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += kThreads) {
    const float* row = logits + (size_t)b * C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, row[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(row[c] - m);
    const int y = (int)labels[b];
    acc += (logf(s) + m) - row[y];
    if (dlogits != nullptr)
      for (int c = 0; c < C; ++c)
        dlogits[(size_t)b * C + c] = cls_weight * (expf(row[c] - m) / s - (c == y ? 1.f : 0.f)) / (float)B;
  }
  // fixed-order block sums of the per-sample CE terms and of the heatmap partials: thread i adds partials i,
  // i + 256, ..., lanes fold by a butterfly, the eight warp sums in ascending order (one thread walking 256 + 592
  // values was 21 us on the step's critical path)
  __shared__ double red[kThreads / 32][2];
  double t = (double)acc, h = 0.0;
  for (int i = threadIdx.x; i < nblk; i += kThreads) h += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t += __shfl_xor_sync(0xffffffffu, t, o);
    h += __shfl_xor_sync(0xffffffffu, h, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = t;
    red[threadIdx.x >> 5][1] = h;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    t = 0.0;
    h = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) {
      t += red[w][0];
      h += red[w][1];
    }
    const float cl = cls_weight * (float)(t / (double)B);
    const float jl = 0.5f * (float)(h * (double)inv_norm);
    out[0] = cl + jl;
    out[1] = cl;
    out[2] = jl;
  }
}

// ------------------------------------------------------------------ AdamW ----
// torch.optim.AdamW (train.py:50-51): decoupled weight decay, bias-corrected moments.
// grad_scale multiplies the incoming gradient (1 / world_size after an all-reduce SUM).
__global__ void __launch_bounds__(kThreads)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
             float grad_scale) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float gi = g[i] * grad_scale;
    float pi = p[i];
    pi *= 1.0f - lr * wd;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

__global__ void __launch_bounds__(kThreads) zero_f32_kernel(float* __restrict__ p, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads)
    p[i] = 0.f;
}

int ew_blocks(long long n) {
  long long b = (n + kThreads - 1) / kThreads;
  const long long cap = 148 * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// ================================================================== host launchers ====

int train_partial_blocks(long long rows) { return reduce_blocks(rows); }

int launch_bn_stats(const __nv_bfloat16* z, long long rows, int C, const float* gamma, const float* beta, float* scale,
                    float* shift, float* mean, float* rstd, float* running_mean, float* running_var, float momentum,
                    float* partial, cudaStream_t st) {
  if (C % 8 != 0 || C > 2048 || kThreads % (C / 8) != 0) {
    set_error("bn_stats: unsupported channel count %d", C);
    return -1;
  }
  const int nblk = reduce_blocks(rows);
  HGR_CHECK_CUDA(launch_pdl(bn_stats_partial_kernel, dim3(nblk), dim3(kThreads), kThreads * 16 * sizeof(float), st,
                            z, rows, C, partial));
  HGR_CHECK_CUDA(launch_pdl(bn_stats_final_kernel, dim3(C), dim3(kFinalThreads), 0, st, partial, nblk, C, rows,
                            gamma, beta, scale, shift, mean, rstd, running_mean, running_var, momentum));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int channel_group_log2(int C) {
  int l = 0;
  while ((8 << l) < C) ++l;
  return (8 << l) == C ? l : -1;
}

int launch_bn_act_fwd(const __nv_bfloat16* z, long long rows, int C, const float* scale, const float* shift, int silu,
                      const __nv_bfloat16* res, int res_ctot, __nv_bfloat16* y, int y_ctot, cudaStream_t st) {
  const int cgl = channel_group_log2(C);
  if (cgl < 0) {
    set_error("bn_act_fwd: channel count %d is not a power of two >= 8", C);
    return -1;
  }
  const int blocks = ew_blocks(rows * (C / 8));
  if (silu && res)
    HGR_CHECK_CUDA(launch_pdl(bn_act_fwd_kernel<true, true>, dim3(blocks), dim3(kThreads), 0, st, z, rows, C, scale,
                              shift, res, res_ctot, y, y_ctot, cgl));
  else if (silu)
    HGR_CHECK_CUDA(launch_pdl(bn_act_fwd_kernel<true, false>, dim3(blocks), dim3(kThreads), 0, st, z, rows, C, scale,
                              shift, res, res_ctot, y, y_ctot, cgl));
  else if (res)
    HGR_CHECK_CUDA(launch_pdl(bn_act_fwd_kernel<false, true>, dim3(blocks), dim3(kThreads), 0, st, z, rows, C, scale,
                              shift, res, res_ctot, y, y_ctot, cgl));
  else
    HGR_CHECK_CUDA(launch_pdl(bn_act_fwd_kernel<false, false>, dim3(blocks), dim3(kThreads), 0, st, z, rows, C,
                              scale, shift, res, res_ctot, y, y_ctot, cgl));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_bn_bwd(const __nv_bfloat16* dy, int dy_ctot, const __nv_bfloat16* z, long long rows, int C, const float* scale,
                  const float* shift, const float* mean, const float* rstd, int silu, const __nv_bfloat16* res,
                  int res_ctot, __nv_bfloat16* dres, int dres_ctot, float* dgamma, float* dbeta, float* c1c2,
                  float* partial, __nv_bfloat16* dz, cudaStream_t st) {
  if (C % 8 != 0 || C > 2048 || kThreads % (C / 8) != 0) {
    set_error("bn_bwd: unsupported channel count %d", C);
    return -1;
  }
  const int nblk = reduce_blocks(rows);
  const size_t sm = kThreads * 16 * sizeof(float);
  const int blocks = ew_blocks(rows * (C / 8));
  const int cgl = channel_group_log2(C);
  if (cgl < 0) {
    set_error("bn_bwd: channel count %d is not a power of two >= 8", C);
    return -1;
  }
  float* c1 = c1c2;
  float* c2 = c1c2 + C;
#define HGR_BN_BWD(S, R)                                                                                           \
  do {                                                                                                             \
    HGR_CHECK_CUDA(launch_pdl(bn_bwd_partial_kernel<S, R>, dim3(nblk), dim3(kThreads), sm, st, dy, dy_ctot, z,         \
                              rows, C, scale, shift, mean, rstd, res, res_ctot, partial));                             \
    HGR_CHECK_CUDA(launch_pdl(bn_bwd_final_kernel, dim3(C), dim3(kFinalThreads), 0, st, partial, nblk, C, rows,        \
                              dgamma, dbeta, c1, c2));                                                                 \
    HGR_CHECK_CUDA(launch_pdl(bn_bwd_apply_kernel<S, R>, dim3(blocks), dim3(kThreads), 0, st, dy, dy_ctot, z,          \
                              rows, C, scale, shift, mean, rstd, c1, c2, res, res_ctot, dres, dres_ctot, dz,           \
                              cgl));                                                                                   \
  } while (0)
  if (silu && res) HGR_BN_BWD(true, true);
  else if (silu) HGR_BN_BWD(true, false);
  else if (res) HGR_BN_BWD(false, true);
  else HGR_BN_BWD(false, false);
#undef HGR_BN_BWD
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_gelu_fwd(const __nv_bfloat16* x, __nv_bfloat16* y, long long n, cudaStream_t st) {
  HGR_CHECK_CUDA(launch_pdl(gelu_fwd_kernel, dim3(ew_blocks(n / 8)), dim3(kThreads), 0, st, x, y, n / 8));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_gelu_bwd(const __nv_bfloat16* pre, __nv_bfloat16* dh, long long n, cudaStream_t st) {
  HGR_CHECK_CUDA(launch_pdl(gelu_bwd_kernel, dim3(ew_blocks(n / 8)), dim3(kThreads), 0, st, pre, dh, n / 8));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_colsum(const __nv_bfloat16* g, long long rows, int C, float* dst, float* partial, cudaStream_t st) {
  if (C % 8 != 0 || C > 2048 || kThreads % (C / 8) != 0) {
    set_error("colsum: unsupported width %d", C);
    return -1;
  }
  const int nblk = reduce_blocks(rows);
  HGR_CHECK_CUDA(launch_pdl(colsum_partial_kernel, dim3(nblk), dim3(kThreads), kThreads * 8 * sizeof(float), st, g,
                            rows, C, partial));
  HGR_CHECK_CUDA(launch_pdl(sums_final_kernel, dim3(C), dim3(kFinalThreads), 0, st, partial, nblk, 1, C, dst,
                            nullptr));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_ln_bwd(const __nv_bfloat16* dy, const __nv_bfloat16* x, const float* gamma, const __nv_bfloat16* g_in,
                  __nv_bfloat16* g_out, long long rows, float* dgamma, float* dbeta, float* partial, cudaStream_t st) {
  const int nblk = reduce_blocks(rows);
  HGR_CHECK_CUDA(launch_pdl(ln_bwd_kernel, dim3(nblk), dim3(kThreads), 0, st, dy, x, gamma, g_in, g_out, rows,
                            partial));
  HGR_CHECK_CUDA(launch_pdl(sums_final_kernel, dim3(256), dim3(kFinalThreads), 0, st, partial, nblk, 2, 256, dgamma,
                            dbeta));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_token_bwd(const __nv_bfloat16* g, __nv_bfloat16* dfeat, float* dcls, int B, int T, cudaStream_t st) {
  const int blocks = ew_blocks((long long)B * (T - 1) * 32) + 1;
  HGR_CHECK_CUDA(launch_pdl(token_bwd_kernel, dim3(blocks), dim3(kThreads), 0, st, g, dfeat, dcls, B, T));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_jobs(const PackJob* d_jobs, int njobs, const float* params, cudaStream_t st) {
  if (njobs <= 0) return 0;
  HGR_CHECK_CUDA(launch_pdl(pack_jobs_kernel, dim3(32, njobs), dim3(kThreads), 0, st, d_jobs, params));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_loss(const float* logits, const float* heat, const long long* labels, const float* target,
                const float* weight, int B, int J, int C, int hw, float cls_weight, float* dlogits, float* dheat,
                float* partial /* >= 592 floats */, float* out3, cudaStream_t st) {
  const long long n = (long long)B * J * hw;
  const float inv_norm = 1.0f / ((float)B * (float)hw * (float)J);
  int nblk = ew_blocks(n);
  if (nblk > 592) nblk = 592;
  HGR_CHECK_CUDA(launch_pdl(loss_heat_kernel, dim3(nblk), dim3(kThreads), 0, st, heat, target, weight, n, hw,
                            inv_norm, dheat, partial));
  HGR_CHECK_CUDA(launch_pdl(loss_final_kernel, dim3(1), dim3(kThreads), 0, st, logits, labels, B, C, cls_weight,
                            dlogits, partial, nblk, inv_norm, out3));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float wd, int step, float grad_scale, cudaStream_t st) {
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  HGR_CHECK_CUDA(launch_pdl(adamw_kernel, dim3(ew_blocks(n)), dim3(kThreads), 0, st, p, g, m, v, n, lr, beta1, beta2,
                            eps, wd, bc1, sqrtf(bc2), grad_scale));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_zero_f32(float* p, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  HGR_CHECK_CUDA(launch_pdl(zero_f32_kernel, dim3(ew_blocks(n)), dim3(kThreads), 0, st, p, n));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hgr
