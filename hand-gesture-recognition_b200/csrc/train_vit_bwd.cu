// Backward kernels of the ViT part of the training step (SURVEY.md 8a rows 10, 13, 14 in reverse):
//   attention core   (reference model/transformer.py:66-74)
//   class head       (transformer.py:113-116,142-144)
//   pose head        (transformer.py:118-127,146-150)
// They run at the 32-crop training batch, where the step is launch- and latency-bound rather than
// FLOP-bound, so these first versions use the CUDA cores with shared-memory staging, fp32 accumulation and
// ownership-based accumulation (no atomics: every output element has exactly one writer, so the gradients
// are bitwise reproducible).
#include "hgr_internal.h"
#include "ptx.cuh"
#include "train.h"

namespace hgr {

namespace {

constexpr int kHeads = 8;
constexpr int kHd = 32;
constexpr int kDim = 256;

// ---------------------------------------------------------------------------------------------
// Attention backward for one (image, head).  With P = softmax(S), S = scale * Q K^T, O = P V:
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - D),  D_i = sum_d dO_id O_id,  dQ = scale * dS K,  dK = scale * dS^T Q.
// P is the bf16 probability map the forward pass stored; query rows are processed in blocks of 32.
// Shared memory (fp32): Q, K, V, dO [T][33], dK, dV accumulators [T][32], one dS / P block [32][T + 1].
// ---------------------------------------------------------------------------------------------
constexpr int kRowBlock = 32;
constexpr int kAThreads = 256;

__global__ void __launch_bounds__(kAThreads)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ probs,
                     const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                     __nv_bfloat16* __restrict__ dqkv, int T, float scale) {
  extern __shared__ float sm[];
  const int P33 = 33;
  float* sq = sm;
  float* sk = sq + T * P33;
  float* sv = sk + T * P33;
  float* sdo = sv + T * P33;
  float* sdk = sdo + T * P33;          // [T][32]
  float* sdv = sdk + T * 32;           // [T][32]
  float* sds = sdv + T * 32;           // [32][T + 1]  dS block
  float* sp = sds + kRowBlock * (T + 1);  // [32][T + 1]  P block
  float* sD = sp + kRowBlock * (T + 1);   // [32]
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int tid = threadIdx.x;
  const size_t row_stride = 3 * kDim;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * row_stride + h * kHd;

  for (int i = tid; i < T * 32; i += kAThreads) {
    const int r = i >> 5, d = i & 31;
    sq[r * P33 + d] = __bfloat162float(qb[(size_t)r * row_stride + d]);
    sk[r * P33 + d] = __bfloat162float(qb[(size_t)r * row_stride + kDim + d]);
    sv[r * P33 + d] = __bfloat162float(qb[(size_t)r * row_stride + 2 * kDim + d]);
    sdo[r * P33 + d] = __bfloat162float(d_o[((size_t)b * T + r) * kDim + h * kHd + d]);
    sdk[i] = 0.f;
    sdv[i] = 0.f;
  }
  __syncthreads();

  const __nv_bfloat16* pb = probs + (size_t)(b * kHeads + h) * T * T;
  for (int i0 = 0; i0 < T; i0 += kRowBlock) {
    const int nr = T - i0 < kRowBlock ? T - i0 : kRowBlock;
    // D_i and the P block
    if (tid < nr) {
      const int r = i0 + tid;
      float dsum = 0.f;
      for (int d = 0; d < 32; ++d)
        dsum = fmaf(sdo[r * P33 + d], __bfloat162float(o[((size_t)b * T + r) * kDim + h * kHd + d]), dsum);
      sD[tid] = dsum;
    }
    for (int i = tid; i < nr * T; i += kAThreads) {
      const int r = i / T, j = i % T;
      sp[r * (T + 1) + j] = __bfloat162float(pb[(size_t)(i0 + r) * T + j]);
    }
    __syncthreads();
    // dS_ij = P_ij * (sum_d dO_id V_jd - D_i)
    for (int i = tid; i < nr * T; i += kAThreads) {
      const int r = i / T, j = i % T;
      float dp = 0.f;
#pragma unroll 8
      for (int d = 0; d < 32; ++d) dp = fmaf(sdo[(i0 + r) * P33 + d], sv[j * P33 + d], dp);
      sds[r * (T + 1) + j] = sp[r * (T + 1) + j] * (dp - sD[r]);
    }
    __syncthreads();
    // dQ rows of this block: thread (r = tid / 32 + 8 m, d = tid % 32)
    {
      const int d = tid & 31;
      for (int r = tid >> 5; r < nr; r += kAThreads / 32) {
        float acc = 0.f;
        for (int j = 0; j < T; ++j) acc = fmaf(sds[r * (T + 1) + j], sk[j * P33 + d], acc);
        dqkv[((size_t)b * T + i0 + r) * row_stride + h * kHd + d] = __float2bfloat16_rn(acc * scale);
      }
    }
    // dK_j += sum_r dS_rj Q_r,  dV_j += sum_r P_rj dO_r : thread owns (j = tid / 32 + 8 m, d = tid % 32)
    {
      const int d = tid & 31;
      for (int j = tid >> 5; j < T; j += kAThreads / 32) {
        float ak = sdk[j * 32 + d], av = sdv[j * 32 + d];
        for (int r = 0; r < nr; ++r) {
          ak = fmaf(sds[r * (T + 1) + j], sq[(i0 + r) * P33 + d], ak);
          av = fmaf(sp[r * (T + 1) + j], sdo[(i0 + r) * P33 + d], av);
        }
        sdk[j * 32 + d] = ak;
        sdv[j * 32 + d] = av;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < T * 32; i += kAThreads) {
    const int r = i >> 5, d = i & 31;
    dqkv[((size_t)b * T + r) * row_stride + kDim + h * kHd + d] = __float2bfloat16_rn(sdk[i] * scale);
    dqkv[((size_t)b * T + r) * row_stride + 2 * kDim + h * kHd + d] = __float2bfloat16_rn(sdv[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// Class head backward: logits = Wc (gamma * x-hat + beta) + bc on the class-token row of every image.
// One CTA, thread c = channel c; the batch is walked in order (fixed summation order).
// dWc accumulators live in shared memory [num_classes][256].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(256)
cls_head_bwd_kernel(const __nv_bfloat16* __restrict__ tokens, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ w, const float* __restrict__ dlogits,
                    int B, int T, int NC, __nv_bfloat16* __restrict__ dtokens, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float* __restrict__ dw, float* __restrict__ dbias) {
  extern __shared__ float smc[];
  float* sdw = smc;                 // [NC][256]
  float* sdl = sdw + NC * kDim;     // [NC] current image's dlogits
  __shared__ float red[8];
  const int c = threadIdx.x;
  for (int j = 0; j < NC; ++j) sdw[j * kDim + c] = 0.f;
  const float gm = gamma[c], bt = beta[c];
  float dg = 0.f, db = 0.f;
  for (int b = 0; b < B; ++b) {
    const float x = __bfloat162float(tokens[(size_t)b * T * kDim + c]);
    const float mean = block_sum_256(x, red) * (1.0f / kDim);
    const float xc = x - mean;
    const float var = block_sum_256(xc * xc, red) * (1.0f / kDim);
    const float rstd = rsqrtf(var + 1e-5f);
    const float xh = xc * rstd;
    const float y = fmaf(gm, xh, bt);
    __syncthreads();
    if (c < NC) sdl[c] = dlogits[(size_t)b * NC + c];
    __syncthreads();
    float dy = 0.f;
    for (int j = 0; j < NC; ++j) {
      const float dl = sdl[j];
      sdw[j * kDim + c] = fmaf(dl, y, sdw[j * kDim + c]);
      dy = fmaf(dl, w[(size_t)j * kDim + c], dy);
    }
    dg = fmaf(dy, xh, dg);
    db += dy;
    const float dyg = dy * gm;
    const float m1 = block_sum_256(dyg, red) * (1.0f / kDim);
    const float m2 = block_sum_256(dyg * xh, red) * (1.0f / kDim);
    dtokens[(size_t)b * T * kDim + c] = __float2bfloat16_rn(rstd * (dyg - m1 - xh * m2));
  }
  dgamma[c] = dg;
  dbeta[c] = db;
  for (int j = 0; j < NC; ++j) dw[(size_t)j * kDim + c] = sdw[j * kDim + c];
  if (c < NC) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dlogits[(size_t)b * NC + c];
    dbias[c] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Pose head backward.  Forward: up = relu(bilinear_x4(tokens[1:])), H = up W^T + bias.
//   dW[j][c] = sum_pix dH[pix][j] up[pix][c];  dup = (up > 0) * (dH W);  dtok = U^T dup.
// A CTA owns (image, token row ty): it walks the output rows whose interpolation touches token row ty
// (y0 == ty or y1 == ty), recomputes up and dup for every pixel of those rows, and accumulates ONLY the
// contributions to its own token row in shared memory -> one writer per gradient element.  dW is counted
// once per pixel (by the CTA with y0 == ty) into a per-CTA partial.  Thread c = channel c.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxJ = 24;
constexpr int kPhases = 4;  // pixel phases per CTA: thread (c, q) handles output columns ox = q (mod 4)

__global__ void __launch_bounds__(256 * kPhases, 1)
pose_head_bwd_kernel(const __nv_bfloat16* __restrict__ tokens, const float* __restrict__ w /*[J][256] fp32*/,
                     const float* __restrict__ dheat, int F, int J, __nv_bfloat16* __restrict__ dtokens,
                     float* __restrict__ dw_partial) {
  extern __shared__ float smp[];
  const int So = 4 * F;
  float* sdx = smp;                        // [kPhases][F][256]  per-phase accumulators of this token row
  float* sdh = sdx + kPhases * F * kDim;   // [J][So]            dheat of the current output row
  float* sw = sdh + kMaxJ * So;            // [J][256]           bf16-rounded weights (as the forward used them)
  const int b = blockIdx.y, ty = blockIdx.x;
  const int c = threadIdx.x & 255, q = threadIdx.x >> 8;
  const int T = F * F + 1;
  const float scale = (float)(F - 1) / (float)(So - 1);
  const __nv_bfloat16* tok = tokens + ((size_t)b * T + 1) * kDim;

  float dwacc[kMaxJ];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) dwacc[j] = 0.f;
  for (int i = threadIdx.x; i < J * kDim; i += 256 * kPhases) sw[i] = __bfloat162float(__float2bfloat16_rn(w[i]));
  float* mydx = sdx + q * F * kDim;
  for (int x = 0; x < F; ++x) mydx[x * kDim + c] = 0.f;

  for (int oy = 0; oy < So; ++oy) {
    const float sy = scale * (float)oy;
    const int y0 = (int)sy;
    const int y1 = y0 + (y0 < F - 1 ? 1 : 0);
    if (y0 != ty && y1 != ty) continue;  // block-uniform
    const float l1 = sy - (float)y0, l0 = 1.0f - l1;
    const float wy = (y0 == ty ? l0 : 0.f) + (y1 == ty ? l1 : 0.f);
    const bool owner = y0 == ty;
    __syncthreads();
    for (int i = threadIdx.x; i < J * So; i += 256 * kPhases) {
      const int j = i / So, ox = i % So;
      sdh[i] = dheat[(((size_t)b * J + j) * So + oy) * So + ox];
    }
    __syncthreads();
    for (int ox = q; ox < So; ox += kPhases) {
      const float sx = scale * (float)ox;
      const int x0 = (int)sx;
      const int x1 = x0 + (x0 < F - 1 ? 1 : 0);
      const float m1 = sx - (float)x0, m0 = 1.0f - m1;
      const float t00 = __bfloat162float(tok[((size_t)y0 * F + x0) * kDim + c]);
      const float t01 = __bfloat162float(tok[((size_t)y0 * F + x1) * kDim + c]);
      const float t10 = __bfloat162float(tok[((size_t)y1 * F + x0) * kDim + c]);
      const float t11 = __bfloat162float(tok[((size_t)y1 * F + x1) * kDim + c]);
      const float up = l0 * (m0 * t00 + m1 * t01) + l1 * (m0 * t10 + m1 * t11);
      if (up > 0.f) {
        float dup = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j)
          if (j < J) {
            const float dh = sdh[j * So + ox];
            dup = fmaf(dh, sw[j * kDim + c], dup);
            if (owner) dwacc[j] = fmaf(dh, up, dwacc[j]);
          }
        const float v = wy * dup;
        mydx[x0 * kDim + c] = fmaf(m0, v, mydx[x0 * kDim + c]);
        mydx[x1 * kDim + c] = fmaf(m1, v, mydx[x1 * kDim + c]);
      }
    }
  }
  __syncthreads();
  // fold the phases in a fixed order: token-row gradient, then the weight-gradient partial (through sdh / sw space)
  if (q == 0)
    for (int x = 0; x < F; ++x) {
      float s = 0.f;
#pragma unroll
      for (int p = 0; p < kPhases; ++p) s += sdx[(p * F + x) * kDim + c];
      dtokens[((size_t)b * T + 1 + (size_t)ty * F + x) * kDim + c] = __float2bfloat16_rn(s);
    }
  __syncthreads();
  float* sacc = sdx;  // [J][256], re-used after the token rows have been written
  for (int p = 0; p < kPhases; ++p) {
    if (q == p) {
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j)
        if (j < J) sacc[j * kDim + c] = (p == 0 ? 0.f : sacc[j * kDim + c]) + dwacc[j];
    }
    __syncthreads();
  }
  float* out = dw_partial + ((size_t)b * gridDim.x + ty) * J * kDim;
  for (int i = threadIdx.x; i < J * kDim; i += 256 * kPhases) out[i] = sacc[i];
}

// dbias[j] = sum_{b, pix} dheat[b][j][pix]: one CTA per joint, fixed order
__global__ void __launch_bounds__(256)
heat_bias_grad_kernel(const float* __restrict__ dheat, int B, int J, int hw, float* __restrict__ dbias) {
  __shared__ float red[256];
  const int j = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* src = dheat + ((size_t)b * J + j) * hw;
    for (int i = threadIdx.x; i < hw; i += 256) s += src[i];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dbias[j] = red[0];
}

}  // namespace

int launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* probs, const __nv_bfloat16* o,
                         const __nv_bfloat16* d_o, __nv_bfloat16* dqkv, int B, int T, cudaStream_t st) {
  const size_t floats = (size_t)4 * T * 33 + (size_t)2 * T * 32 + (size_t)2 * kRowBlock * (T + 1) + kRowBlock;
  const size_t smem = floats * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("attention_bwd: %d tokens do not fit one CTA's shared memory", T);
    return -1;
  }
  HGR_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_bwd_kernel<<<B * kHeads, kAThreads, smem, st>>>(qkv, probs, o, d_o, dqkv, T, 0.17677669529663687f);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_cls_head_bwd(const __nv_bfloat16* tokens, const float* gamma, const float* beta, const float* w,
                        const float* dlogits, int B, int T, int NC, __nv_bfloat16* dtokens, float* dgamma,
                        float* dbeta, float* dw, float* dbias, cudaStream_t st) {
  if (NC < 1 || NC > 128) {
    set_error("cls_head_bwd: num_classes %d unsupported in training (1..128)", NC);
    return -1;
  }
  const size_t smem = ((size_t)NC * kDim + NC) * sizeof(float);
  HGR_CHECK_CUDA(cudaFuncSetAttribute(cls_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cls_head_bwd_kernel<<<1, 256, smem, st>>>(tokens, gamma, beta, w, dlogits, B, T, NC, dtokens, dgamma, dbeta, dw, dbias);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pose_head_bwd(const __nv_bfloat16* tokens, const float* w, const float* dheat, int B, int F, int J,
                         __nv_bfloat16* dtokens, float* dw_partial, float* dw, float* dbias, cudaStream_t st) {
  if (J < 1 || J > kMaxJ) {
    set_error("pose_head_bwd: unsupported J=%d", J);
    return -1;
  }
  const int So = 4 * F;
  size_t accf = (size_t)kPhases * F * kDim;
  if (accf < (size_t)kMaxJ * kDim) accf = (size_t)kMaxJ * kDim;  // the phase fold re-uses this space as [J][256]
  const size_t smem = (accf + (size_t)kMaxJ * So + (size_t)kMaxJ * kDim) * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("pose_head_bwd: feature side %d does not fit shared memory", F);
    return -1;
  }
  HGR_CHECK_CUDA(cudaFuncSetAttribute(pose_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pose_head_bwd_kernel<<<dim3(F, B), 256 * kPhases, smem, st>>>(tokens, w, dheat, F, J, dtokens, dw_partial);
  if (int rc = launch_partial_sum(dw_partial, B * F, J * kDim, dw, st)) return rc;
  heat_bias_grad_kernel<<<J, 256, 0, st>>>(dheat, B, J, So * So, dbias);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hgr
