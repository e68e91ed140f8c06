// Backward kernels of the ViT part of the training step (SURVEY.md 8a rows 10, 13, 14 in reverse):
//   attention core   (reference model/transformer.py:66-74)
//   class head       (transformer.py:113-116,142-144)
//   pose head        (transformer.py:118-127,146-150)
// They run at the 32-crop training batch, where the step is launch- and latency-bound rather than
// FLOP-bound: the attention backward runs on mma.sync tiles, the two heads on the CUDA cores with shared-memory
// staging; all use fp32 accumulation and ownership-based accumulation (no atomics: every output element has
// exactly one writer, so the gradients are bitwise reproducible).
#include "hgr_internal.h"
#include "ptx.cuh"
#include "train.h"

namespace hgr {

namespace {

constexpr int kHeads = 8;
constexpr int kHd = 32;
constexpr int kDim = 256;

// ---------------------------------------------------------------------------------------------
// Attention backward for one (image, head) on mma.sync m16n8k16.  With P = softmax(S), S = scale * Q K^T, O = P V:
//   dP = dO V^T,  D_i = sum_d dO_id O_id,  dS = P o (dP - D),  dQ = scale * dS K,  dK = scale * dS^T Q,  dV = P^T dO.
// Q, K, V, dO (T x 32 each) and the bf16 probability map P the forward pass stored (T x T) are staged in shared
// memory once.  Pass A: warp per 16-query tile, dP -> dS in registers (accumulator layout == A-fragment layout,
// the forward kernel's P.V trick) -> dQ.  Pass B: warp per 16-key tile, the TRANSPOSED products
// dP^T = V dO^T, dS^T = P^T o (dP^T - D) -> dK = dS^T Q and dV = P^T dO, so every output row has one owner and no
// cross-warp reduction (bitwise reproducible).
// ---------------------------------------------------------------------------------------------
constexpr int kBPitch = 40;   // bf16 elements per staged Q/K/V/dO row (80 B: conflict-free ldmatrix)
constexpr int kBWarps = 5;
constexpr int kBThreads = kBWarps * 32;

// 16 x 32 block of A(16 x 32) . B^T where B rows ([n][k], pitch kBPitch) start at b_addr_lane
__device__ __forceinline__ void mm_block_nt(const uint32_t (&a)[2][4], uint32_t b_addr_lane, float (&c)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    uint32_t bf[4];
    ldmatrix_x4(bf, b_addr_lane + nt * 8 * kBPitch * 2);
    c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
    mma_bf16_16816(c[nt], a[0], bf[0], bf[1]);
    mma_bf16_16816(c[nt], a[1], bf[2], bf[3]);
  }
}

// acc(16 x 32) += A(16 x 32, fragments built from `v`) . B(32 x 32) with B rows ([k][n], pitch kBPitch) at b_addr_lane
__device__ __forceinline__ void mm_block_nn(const float (&v)[4][4], uint32_t b_addr_lane, float (&acc)[4][4]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint32_t pa[4];
    pa[0] = pack_bf16x2(v[2 * j][0], v[2 * j][1]);
    pa[1] = pack_bf16x2(v[2 * j][2], v[2 * j][3]);
    pa[2] = pack_bf16x2(v[2 * j + 1][0], v[2 * j + 1][1]);
    pa[3] = pack_bf16x2(v[2 * j + 1][2], v[2 * j + 1][3]);
    const uint32_t addr = b_addr_lane + j * 16 * kBPitch * 2;
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bf[4];
      ldmatrix_x4_trans(bf, addr + np * 32);
      mma_bf16_16816(acc[2 * np], pa, bf[0], bf[1]);
      mma_bf16_16816(acc[2 * np + 1], pa, bf[2], bf[3]);
    }
  }
}

// PSMEM = false (T > 256 tokens, e.g. 257 at 256x256): the probability map does not fit next to Q/K/V/dO, so it is
// read straight from global memory (L2-resident, 150 KB per head) with explicit bounds instead of zero padding.
template <bool PSMEM>
__global__ void __launch_bounds__(kBThreads)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ probs, int pp,
                     const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                     __nv_bfloat16* __restrict__ dqkv, int T, int Tp, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smraw[];
  const int spp = Tp + 8;  // staged P pitch (elements)
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* sk = sq + Tp * kBPitch;
  __nv_bfloat16* sv = sk + Tp * kBPitch;
  __nv_bfloat16* sdo = sv + Tp * kBPitch;
  __nv_bfloat16* sp = sdo + Tp * kBPitch;  // [Tp][spp] (PSMEM only)
  float* sD = reinterpret_cast<float*>(sp + (PSMEM ? Tp * spp : 0));  // [Tp]
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t rs = 3 * kDim;
  const __nv_bfloat16* qb = qkv + (size_t)b * T * rs + h * kHd;
  const __nv_bfloat16* dob = d_o + (size_t)b * T * kDim + h * kHd;
  const __nv_bfloat16* ob = o + (size_t)b * T * kDim + h * kHd;

  // ---- stage Q, K, V, dO (rows >= T zero) ----
  for (int i = tid; i < 4 * Tp * 4; i += kBThreads) {
    const int c = i & 3, row = (i >> 2) % Tp, part = (i >> 2) / Tp;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < T)
      v = part < 3 ? __ldg(reinterpret_cast<const uint4*>(qb + (size_t)row * rs + part * kDim) + c)
                   : __ldg(reinterpret_cast<const uint4*>(dob + (size_t)row * kDim) + c);
    *reinterpret_cast<uint4*>(sq + (part * Tp + row) * kBPitch + c * 8) = v;
  }
  // ---- stage P (rows / columns >= T zero) ----
  const __nv_bfloat16* pb = probs + (size_t)(b * kHeads + h) * T * pp;
  if constexpr (!PSMEM) {
    // nothing to stage
  } else if ((pp & 7) == 0 && pp >= Tp) {
    const int cpr = Tp >> 3;  // the buffer's pad columns are zero (cleared once by the plan)
    for (int i = tid; i < Tp * cpr; i += kBThreads) {
      const int row = i / cpr, c = i % cpr;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < T) v = __ldg(reinterpret_cast<const uint4*>(pb + (size_t)row * pp) + c);
      *reinterpret_cast<uint4*>(sp + row * spp + c * 8) = v;
    }
  } else {
    for (int i = tid; i < Tp * Tp; i += kBThreads) {
      const int row = i / Tp, c = i % Tp;
      sp[row * spp + c] = (row < T && c < T) ? pb[(size_t)row * pp + c] : __float2bfloat16_rn(0.f);
    }
  }
  // ---- D_i = sum_d dO_id O_id ----
  for (int r = tid; r < Tp; r += kBThreads) {
    float dsum = 0.f;
    if (r < T) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a4 = __ldg(reinterpret_cast<const uint4*>(dob + (size_t)r * kDim) + c);
        const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(ob + (size_t)r * kDim) + c);
        const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          dsum = fmaf(bf16_lo(aw[e]), bf16_lo(bw[e]), fmaf(bf16_hi(aw[e]), bf16_hi(bw[e]), dsum));
      }
    }
    sD[r] = dsum;
  }
  __syncthreads();

  auto ldp = [&](int row, int col) -> float {
    return (row < T && col < T) ? __bfloat162float(pb[(size_t)row * pp + col]) : 0.f;
  };
  const int ntiles = (T + 15) >> 4;
  const int nblocks = Tp >> 5;  // 32-wide blocks
  // lane addresses: A fragments of 16 rows (ldmatrix), "n-major" B (rows = n, ldmatrix), "k-major" B (rows = k, .trans)
  const uint32_t a_off = ((lane & 15) * kBPitch + (lane >> 4) * 8) * 2;
  const uint32_t bn_off = ((lane & 7) * kBPitch + (lane >> 3) * 8) * 2;
  const uint32_t bk_off = ((((lane >> 3) & 1) * 8 + (lane & 7)) * kBPitch + (lane >> 4) * 8) * 2;
  const uint32_t uq = smem_u32(sq), uk = smem_u32(sk), uv = smem_u32(sv), udo = smem_u32(sdo);

  // ================= pass A: dQ, warp per 16-query tile =================
  for (int mt = warp; mt < ntiles; mt += kBWarps) {
    uint32_t da[2][4];
    ldmatrix_x4(da[0], udo + a_off + mt * 16 * kBPitch * 2);
    ldmatrix_x4(da[1], udo + a_off + mt * 16 * kBPitch * 2 + 32);
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const float D0 = sD[r0], D1 = sD[r1];
    float acc[4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
    for (int kb = 0; kb < nblocks; ++kb) {
      float ds[4][4];
      mm_block_nt(da, uv + bn_off + kb * 32 * kBPitch * 2, ds);  // dP block = dO V^T
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = kb * 32 + nt * 8 + 2 * t;
        float p00, p01, p10, p11;
        if constexpr (PSMEM) {
          const uint32_t p0 = *reinterpret_cast<const uint32_t*>(sp + r0 * spp + key);
          const uint32_t p1 = *reinterpret_cast<const uint32_t*>(sp + r1 * spp + key);
          p00 = bf16_lo(p0);
          p01 = bf16_hi(p0);
          p10 = bf16_lo(p1);
          p11 = bf16_hi(p1);
        } else {
          p00 = ldp(r0, key);
          p01 = ldp(r0, key + 1);
          p10 = ldp(r1, key);
          p11 = ldp(r1, key + 1);
        }
        ds[nt][0] = p00 * (ds[nt][0] - D0);
        ds[nt][1] = p01 * (ds[nt][1] - D0);
        ds[nt][2] = p10 * (ds[nt][2] - D1);
        ds[nt][3] = p11 * (ds[nt][3] - D1);
      }
      mm_block_nn(ds, uk + bk_off + kb * 32 * kBPitch * 2, acc);  // dQ += dS K
    }
    __nv_bfloat16* q0 = dqkv + ((size_t)b * T + r0) * rs + h * kHd + 2 * t;
    __nv_bfloat16* q1 = dqkv + ((size_t)b * T + r1) * rs + h * kHd + 2 * t;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      if (r0 < T) *reinterpret_cast<uint32_t*>(q0 + nd * 8) = pack_bf16x2(acc[nd][0] * scale, acc[nd][1] * scale);
      if (r1 < T) *reinterpret_cast<uint32_t*>(q1 + nd * 8) = pack_bf16x2(acc[nd][2] * scale, acc[nd][3] * scale);
    }
  }

  // ================= pass B: dK, dV, warp per 16-key tile =================
  const unsigned short* spu = reinterpret_cast<const unsigned short*>(sp);
  for (int kt = warp; kt < ntiles; kt += kBWarps) {
    uint32_t va[2][4];
    ldmatrix_x4(va[0], uv + a_off + kt * 16 * kBPitch * 2);
    ldmatrix_x4(va[1], uv + a_off + kt * 16 * kBPitch * 2 + 32);
    const int c0 = kt * 16 + g, c1 = c0 + 8;
    float ak[4][4], av[4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      ak[nd][0] = ak[nd][1] = ak[nd][2] = ak[nd][3] = 0.f;
      av[nd][0] = av[nd][1] = av[nd][2] = av[nd][3] = 0.f;
    }
    for (int qblk = 0; qblk < nblocks; ++qblk) {
      float dst[4][4], pt[4][4];
      mm_block_nt(va, udo + bn_off + qblk * 32 * kBPitch * 2, dst);  // dP^T block = V dO^T
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int q = qblk * 32 + nt * 8 + 2 * t;
        const float Dq0 = sD[q], Dq1 = sD[q + 1];
        if constexpr (PSMEM) {
          pt[nt][0] = __uint_as_float((uint32_t)spu[q * spp + c0] << 16);
          pt[nt][1] = __uint_as_float((uint32_t)spu[(q + 1) * spp + c0] << 16);
          pt[nt][2] = __uint_as_float((uint32_t)spu[q * spp + c1] << 16);
          pt[nt][3] = __uint_as_float((uint32_t)spu[(q + 1) * spp + c1] << 16);
        } else {
          pt[nt][0] = ldp(q, c0);
          pt[nt][1] = ldp(q + 1, c0);
          pt[nt][2] = ldp(q, c1);
          pt[nt][3] = ldp(q + 1, c1);
        }
        dst[nt][0] = pt[nt][0] * (dst[nt][0] - Dq0);
        dst[nt][1] = pt[nt][1] * (dst[nt][1] - Dq1);
        dst[nt][2] = pt[nt][2] * (dst[nt][2] - Dq0);
        dst[nt][3] = pt[nt][3] * (dst[nt][3] - Dq1);
      }
      mm_block_nn(dst, uq + bk_off + qblk * 32 * kBPitch * 2, ak);  // dK += dS^T Q
      mm_block_nn(pt, udo + bk_off + qblk * 32 * kBPitch * 2, av);  // dV += P^T dO
    }
    __nv_bfloat16* k0 = dqkv + ((size_t)b * T + c0) * rs + kDim + h * kHd + 2 * t;
    __nv_bfloat16* k1 = dqkv + ((size_t)b * T + c1) * rs + kDim + h * kHd + 2 * t;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      if (c0 < T) {
        *reinterpret_cast<uint32_t*>(k0 + nd * 8) = pack_bf16x2(ak[nd][0] * scale, ak[nd][1] * scale);
        *reinterpret_cast<uint32_t*>(k0 + kDim + nd * 8) = pack_bf16x2(av[nd][0], av[nd][1]);
      }
      if (c1 < T) {
        *reinterpret_cast<uint32_t*>(k1 + nd * 8) = pack_bf16x2(ak[nd][2] * scale, ak[nd][3] * scale);
        *reinterpret_cast<uint32_t*>(k1 + kDim + nd * 8) = pack_bf16x2(av[nd][2], av[nd][3]);
      }
    }
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
  pdl_launch_dependents();
  pdl_wait();
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
constexpr int kPhases = 2;  // pixel phases per CTA: thread (c, q) handles output columns [q So / 2, (q + 1) So / 2)

// Shared-memory traffic decides this kernel (42 loads per pixel and channel in the naive form), so the weight
// column of a thread lives in registers and the dheat row is staged TRANSPOSED ([ox][24 joints]) so that the 21
// joint values of a pixel arrive as six broadcast 128-bit loads.  The three token rows a CTA interpolates from
// (ty - 1, ty, ty + 1) are staged once: four dependent global loads per pixel were the kernel's latency chain
// (0.31 ms at batch 32 before, one L2 round trip per inner iteration).
__global__ void __launch_bounds__(256 * kPhases, 1)
pose_head_bwd_kernel(const __nv_bfloat16* __restrict__ tokens, const float* __restrict__ w /*[J][256] fp32*/,
                     const float* __restrict__ dheat, int F, int J, __nv_bfloat16* __restrict__ dtokens,
                     float* __restrict__ dw_partial) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float smp[];
  const int So = 4 * F;
  float* sdx = smp;                       // [kPhases][F][256]  per-phase accumulators of this token row
  float* sdh = sdx + kPhases * F * kDim;  // [So][kMaxJ]        dheat of the current output row, joint-contiguous
  __nv_bfloat16* stok = reinterpret_cast<__nv_bfloat16*>(sdh + So * kMaxJ);  // [3][F][256] token rows ty-1 .. ty+1
  const int b = blockIdx.y, ty = blockIdx.x;
  const int c = threadIdx.x & 255, q = threadIdx.x >> 8;
  const int T = F * F + 1;
  const float scale = (float)(F - 1) / (float)(So - 1);
  const __nv_bfloat16* tok = tokens + ((size_t)b * T + 1) * kDim;

  float wr[kMaxJ], dwacc[kMaxJ];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) {
    wr[j] = j < J ? __bfloat162float(__float2bfloat16_rn(w[(size_t)j * kDim + c])) : 0.f;  // the forward used bf16 weights
    dwacc[j] = 0.f;
  }
  for (int i = threadIdx.x; i < 3 * F * (kDim / 8); i += 256 * kPhases) {
    const int r = i / (F * (kDim / 8)), y = ty - 1 + r;
    if (y >= 0 && y < F)
      reinterpret_cast<uint4*>(stok)[i] =
          __ldg(reinterpret_cast<const uint4*>(tok + (size_t)y * F * kDim) + i % (F * (kDim / 8)));
  }
  float* mydx = sdx + q * F * kDim;
  for (int x = 0; x < F; ++x) mydx[x * kDim + c] = 0.f;
  for (int i = threadIdx.x; i < So * kMaxJ; i += 256 * kPhases) sdh[i] = 0.f;  // joints >= J stay zero

  for (int oy = 0; oy < So; ++oy) {
    const float sy = scale * (float)oy;
    const int y0 = (int)sy;
    const int y1 = y0 + (y0 < F - 1 ? 1 : 0);
    if (y0 != ty && y1 != ty) continue;  // block-uniform
    const float l1 = sy - (float)y0, l0 = 1.0f - l1;
    const float wy = (y0 == ty ? l0 : 0.f) + (y1 == ty ? l1 : 0.f);
    const bool owner = y0 == ty;
    const __nv_bfloat16* r0 = stok + (size_t)(y0 - ty + 1) * F * kDim + c;
    const __nv_bfloat16* r1 = stok + (size_t)(y1 - ty + 1) * F * kDim + c;
    __syncthreads();
    for (int i = threadIdx.x; i < J * So; i += 256 * kPhases) {
      const int j = i / So, ox = i % So;
      sdh[ox * kMaxJ + j] = dheat[(((size_t)b * J + j) * So + oy) * So + ox];
    }
    __syncthreads();
    // Phase q walks the output columns [q So / 2, (q + 1) So / 2) in order: the source column x0 advances by at most
    // one per pixel (scale < 1), so the two row-interpolated token values around the pixel (a0 at x0, a1 at x0 + 1)
    // and the gradients pending for those two tokens (g0, g1) live in registers and touch shared memory only when
    // x0 moves - per pixel that leaves the six dheat loads and the 48 FMAs.
    const int ox_begin = q * (So / kPhases), ox_end = ox_begin + So / kPhases;
    auto colval = [&](int x) { return l0 * __bfloat162float(r0[x * kDim]) + l1 * __bfloat162float(r1[x * kDim]); };
    int cx = (int)(scale * (float)ox_begin);
    float a0 = colval(cx), a1 = colval(cx + 1 < F ? cx + 1 : F - 1), g0 = 0.f, g1 = 0.f;
    for (int ox = ox_begin; ox < ox_end; ++ox) {
      const float sx = scale * (float)ox;
      const int x0 = (int)sx;
      const float m1 = sx - (float)x0, m0 = 1.0f - m1;
      if (x0 != cx) {
        mydx[cx * kDim + c] += g0;
        g0 = g1;
        g1 = 0.f;
        a0 = a1;
        cx = x0;
        a1 = colval(cx + 1 < F ? cx + 1 : F - 1);  // at the last column m1 is 0
      }
      const float up = m0 * a0 + m1 * a1;
      if (up > 0.f) {
        float dup = 0.f;
        const float4* dh4 = reinterpret_cast<const float4*>(sdh + ox * kMaxJ);
#pragma unroll
        for (int j4 = 0; j4 < kMaxJ / 4; ++j4) {
          const float4 d = dh4[j4];
          const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            dup = fmaf(dv[e], wr[4 * j4 + e], dup);
            if (owner) dwacc[4 * j4 + e] = fmaf(dv[e], up, dwacc[4 * j4 + e]);
          }
        }
        const float v = wy * dup;
        g0 = fmaf(m0, v, g0);
        g1 = fmaf(m1, v, g1);
      }
    }
    mydx[cx * kDim + c] += g0;
    if (cx + 1 < F) mydx[(cx + 1) * kDim + c] += g1;
  }
  __syncthreads();
  // fold the phases in a fixed order: token-row gradient, then the weight-gradient partial
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
  pdl_launch_dependents();
  pdl_wait();
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

int launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* probs, int probs_pitch,
                         const __nv_bfloat16* o, const __nv_bfloat16* d_o, __nv_bfloat16* dqkv, int B, int T,
                         cudaStream_t st) {
  const int Tp = (T + 31) / 32 * 32;
  const int pp = probs_pitch > 0 ? probs_pitch : T;
  const size_t base = (size_t)4 * Tp * kBPitch * 2 + (size_t)Tp * sizeof(float);
  const size_t with_p = base + (size_t)Tp * (Tp + 8) * 2;
  if (base > 227 * 1024) {
    set_error("attention_bwd: %d tokens do not fit one CTA's shared memory (%zu bytes)", T, base);
    return -1;
  }
  const float scale = 0.17677669529663687f;
  if (with_p <= 227 * 1024) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)with_p));
    HGR_CHECK_CUDA(launch_pdl(attention_bwd_kernel<true>, dim3(B * kHeads), dim3(kBThreads), with_p, st, qkv, probs,
                              pp, o, d_o, dqkv, T, Tp, scale));
  } else {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base));
    HGR_CHECK_CUDA(launch_pdl(attention_bwd_kernel<false>, dim3(B * kHeads), dim3(kBThreads), base, st, qkv, probs,
                              pp, o, d_o, dqkv, T, Tp, scale));
  }
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
  HGR_CHECK_CUDA(launch_pdl(cls_head_bwd_kernel, dim3(1), dim3(256), smem, st, tokens, gamma, beta, w, dlogits, B, T,
                            NC, dtokens, dgamma, dbeta, dw, dbias));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pose_head_bwd(const __nv_bfloat16* tokens, const float* w, const float* dheat, int B, int F, int J,
                         __nv_bfloat16* dtokens, float* dw_partial, float* dw, cudaStream_t st) {
  if (J < 1 || J > kMaxJ) {
    set_error("pose_head_bwd: unsupported J=%d", J);
    return -1;
  }
  const int So = 4 * F;
  size_t accf = (size_t)kPhases * F * kDim;
  if (accf < (size_t)kMaxJ * kDim) accf = (size_t)kMaxJ * kDim;  // the phase fold re-uses this space as [J][256]
  const size_t smem = (accf + (size_t)kMaxJ * So) * sizeof(float) + (size_t)3 * F * kDim * sizeof(__nv_bfloat16);
  if (smem > 227 * 1024) {
    set_error("pose_head_bwd: feature side %d does not fit shared memory", F);
    return -1;
  }
  HGR_CHECK_CUDA(cudaFuncSetAttribute(pose_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HGR_CHECK_CUDA(launch_pdl(pose_head_bwd_kernel, dim3(F, B), dim3(256 * kPhases), smem, st, tokens, w, dheat, F, J,
                            dtokens, dw_partial));
  return launch_partial_sum(dw_partial, B * F, J * kDim, dw, st);
}

// d bias of the pose head: independent of the kernels above (the trainer's plan runs it beside them)
int launch_heat_bias_grad(const float* dheat, int B, int J, int hw, float* dbias, cudaStream_t st) {
  HGR_CHECK_CUDA(launch_pdl(heat_bias_grad_kernel, dim3(J), dim3(256), 0, st, dheat, B, J, hw, dbias));
  return 0;
}

}  // namespace hgr
