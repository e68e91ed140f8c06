// Fused short-sequence attention core for the ViT
// (reference model/transformer.py:66-74):
//     q,k,v = split(to_qkv(LN(x)));  P = softmax(q k^T * d^-0.5);  out = P v
// with 8 heads of width 32 and T = 145 (192x192) or 257 (256x256) tokens.
//
// One CTA owns one (image, head): its Q, K and V slices (T x 32 bf16 each)
// live in shared memory for the whole kernel, the score matrix never leaves
// registers, and the result is written already in the 'b n (h d)' layout the
// output projection consumes, so the reference's three rearrange copies and
// the HBM round trip of the (B, 8, T, T) score tensor disappear.  The last
// layer optionally writes the normalised probabilities, because
// Transformer.forward returns them (transformer.py:90-96).
//
// The contraction runs on mma.sync m16n8k16 (bf16 in, fp32 accumulate): the
// whole attention core is 2 % of the network's FLOPs and a (145 x 145 x 32)
// problem does not fill a 128-row tcgen05 tile.  Softmax is two-pass: pass 1
// keeps only the running row max / sum, pass 2 recomputes the (bitwise
// identical) scores, normalises, and feeds P straight back into the P.V MMAs
// from registers.
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kHeads = 8;
constexpr int kHd = 32;        // head width
constexpr int kPitch = 40;     // smem row pitch in bf16 (80 B): conflict-free ldmatrix
constexpr int kWarps = 5;
constexpr int kKeyBlock = 32;

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename TP>
__device__ __forceinline__ void store_prob(TP* p, float v);
template <>
__device__ __forceinline__ void store_prob<float>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void store_prob<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// Scores of a 16-query x 32-key block: s[nt][0..1] = row g, keys nt*8+2t,+1; s[nt][2..3] = row g+8.
__device__ __forceinline__ void score_block(const uint32_t (&qa)[2][4], uint32_t k_addr_lane, float (&s)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    uint32_t kb[4];
    ldmatrix_x4(kb, k_addr_lane + nt * 8 * kPitch * 2);
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    mma_bf16_16816(s[nt], qa[0], kb[0], kb[1]);
    mma_bf16_16816(s[nt], qa[1], kb[2], kb[3]);
  }
}

template <typename TP>
__global__ void __launch_bounds__(kWarps * 32)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, TP* __restrict__ probs, int T,
                 int Tp, float scale_log2e, int pp) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sk = sq + Tp * kPitch;
  __nv_bfloat16* sv = sk + Tp * kPitch;

  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  pdl_launch_dependents();
  pdl_wait();

  // ---- stage Q, K, V (zero rows beyond T) --------------------------------
  const __nv_bfloat16* base = qkv + (size_t)b * T * (3 * kHeads * kHd) + h * kHd;
  for (int i = tid; i < 3 * Tp * 4; i += kWarps * 32) {
    const int c = i & 3;
    const int row = (i >> 2) % Tp;
    const int part = (i >> 2) / Tp;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < T) v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)row * (3 * kHeads * kHd) + part * kHeads * kHd) + c);
    *reinterpret_cast<uint4*>(sq + (part * Tp + row) * kPitch + c * 8) = v;
  }
  __syncthreads();

  const int mtiles = (T + 15) >> 4;
  const int kblocks = Tp / kKeyBlock;
  // ldmatrix lane addresses
  const uint32_t q_lane = smem_u32(sq) + ((lane & 15) * kPitch + (lane >> 4) * 8) * 2;
  const uint32_t k_lane = smem_u32(sk) + ((lane & 7) * kPitch + (lane >> 3) * 8) * 2;
  const uint32_t v_lane = smem_u32(sv) + ((((lane >> 3) & 1) * 8 + (lane & 7)) * kPitch + (lane >> 4) * 8) * 2;

  for (int mt = warp; mt < mtiles; mt += kWarps) {
    uint32_t qa[2][4];
    ldmatrix_x4(qa[0], q_lane + mt * 16 * kPitch * 2);
    ldmatrix_x4(qa[1], q_lane + mt * 16 * kPitch * 2 + 32);

    // ---- pass 1: row max and sum of exponentials ----
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kb = 0; kb < kblocks; ++kb) {
      float s[4][4];
      score_block(qa, k_lane + kb * kKeyBlock * kPitch * 2, s);
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = kb * kKeyBlock + nt * 8 + 2 * t;
        if (key >= T) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= T) s[nt][1] = s[nt][3] = -INFINITY;
        bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
        bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
      }
      // key 0 is always valid, so the running max is finite from the first block on
      const float n0 = fmaxf(m0, quad_max(bm0)), n1 = fmaxf(m1, quad_max(bm1));
      l0 *= ex2((m0 - n0) * scale_log2e);
      l1 *= ex2((m1 - n1) * scale_log2e);
      m0 = n0;
      m1 = n1;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        l0 += ex2((s[nt][0] - m0) * scale_log2e) + ex2((s[nt][1] - m0) * scale_log2e);
        l1 += ex2((s[nt][2] - m1) * scale_log2e) + ex2((s[nt][3] - m1) * scale_log2e);
      }
    }
    const float inv0 = 1.0f / quad_sum(l0), inv1 = 1.0f / quad_sum(l1);

    // ---- pass 2: normalised probabilities and P.V ----
    float o[4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
    const int row0 = mt * 16 + g, row1 = row0 + 8;
    for (int kb = 0; kb < kblocks; ++kb) {
      float s[4][4];
      score_block(qa, k_lane + kb * kKeyBlock * kPitch * 2, s);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = kb * kKeyBlock + nt * 8 + 2 * t;
        s[nt][0] = key < T ? ex2((s[nt][0] - m0) * scale_log2e) * inv0 : 0.f;
        s[nt][1] = key + 1 < T ? ex2((s[nt][1] - m0) * scale_log2e) * inv0 : 0.f;
        s[nt][2] = key < T ? ex2((s[nt][2] - m1) * scale_log2e) * inv1 : 0.f;
        s[nt][3] = key + 1 < T ? ex2((s[nt][3] - m1) * scale_log2e) * inv1 : 0.f;
        if (probs != nullptr) {
          TP* pr = probs + ((size_t)(b * kHeads + h) * T) * pp;
          if (row0 < T) {
            if (key < T) store_prob<TP>(pr + (size_t)row0 * pp + key, s[nt][0]);
            if (key + 1 < T) store_prob<TP>(pr + (size_t)row0 * pp + key + 1, s[nt][1]);
          }
          if (row1 < T) {
            if (key < T) store_prob<TP>(pr + (size_t)row1 * pp + key, s[nt][2]);
            if (key + 1 < T) store_prob<TP>(pr + (size_t)row1 * pp + key + 1, s[nt][3]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {  // two 16-key k-steps per block
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
        pa[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
        pa[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
        pa[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
        const uint32_t vaddr = v_lane + (kb * kKeyBlock + j * 16) * kPitch * 2;
#pragma unroll
        for (int np = 0; np < 2; ++np) {  // two pairs of 8-wide d tiles
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, vaddr + np * 32);
          mma_bf16_16816(o[2 * np], pa, vb[0], vb[1]);
          mma_bf16_16816(o[2 * np + 1], pa, vb[2], vb[3]);
        }
      }
    }
    // ---- write 'b n (h d)' ----
    __nv_bfloat16* orow0 = out + ((size_t)b * T + row0) * (kHeads * kHd) + h * kHd + 2 * t;
    __nv_bfloat16* orow1 = out + ((size_t)b * T + row1) * (kHeads * kHd) + h * kHd + 2 * t;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      if (row0 < T) *reinterpret_cast<uint32_t*>(orow0 + nd * 8) = pack_bf16x2(o[nd][0], o[nd][1]);
      if (row1 < T) *reinterpret_cast<uint32_t*>(orow1 + nd * 8) = pack_bf16x2(o[nd][2], o[nd][3]);
    }
  }
}

// Single-pass variant for T <= 32 * NKB keys: the whole 16 x T score strip of a
// warp stays in registers (NKB * 16 fp32), so Q.K^T and the exponentials are
// evaluated once instead of twice.  Used for the 145-token (192x192) case.
template <typename TP, int NKB>
__global__ void __launch_bounds__(kWarps * 32, 3)
attention_kernel_1pass(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, TP* __restrict__ probs,
                       int T, float scale_log2e, int reverse, int pp) {
  constexpr int Tp = NKB * kKeyBlock;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sk = sq + Tp * kPitch;
  __nv_bfloat16* sv = sk + Tp * kPitch;

  pdl_launch_dependents();
  pdl_wait();
  const int bid = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int b = bid / kHeads, h = bid % kHeads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  const __nv_bfloat16* base = qkv + (size_t)b * T * (3 * kHeads * kHd) + h * kHd;
  {
    // all 16-byte requests of this thread are issued before the first one is consumed
    constexpr int kPer = 3 * Tp * 4 / (kWarps * 32);
    static_assert(kPer * kWarps * 32 == 3 * Tp * 4, "chunk count must divide evenly");
    uint4 v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * kWarps * 32;
      const int c = i & 3, row = (i >> 2) % Tp, part = (i >> 2) / Tp;
      v[k] = make_uint4(0, 0, 0, 0);
      if (row < T)
        v[k] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)row * (3 * kHeads * kHd) + part * kHeads * kHd) + c);
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * kWarps * 32;
      const int c = i & 3, row = (i >> 2) % Tp, part = (i >> 2) / Tp;
      *reinterpret_cast<uint4*>(sq + (part * Tp + row) * kPitch + c * 8) = v[k];
    }
  }
  __syncthreads();

  const int mtiles = (T + 15) >> 4;
  const uint32_t q_lane = smem_u32(sq) + ((lane & 15) * kPitch + (lane >> 4) * 8) * 2;
  const uint32_t k_lane = smem_u32(sk) + ((lane & 7) * kPitch + (lane >> 3) * 8) * 2;
  const uint32_t v_lane = smem_u32(sv) + ((((lane >> 3) & 1) * 8 + (lane & 7)) * kPitch + (lane >> 4) * 8) * 2;

  for (int mt = warp; mt < mtiles; mt += kWarps) {
    uint32_t qa[2][4];
    ldmatrix_x4(qa[0], q_lane + mt * 16 * kPitch * 2);
    ldmatrix_x4(qa[1], q_lane + mt * 16 * kPitch * 2 + 32);

    float s[NKB][4][4];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      score_block(qa, k_lane + kb * kKeyBlock * kPitch * 2, s[kb]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (kb == NKB - 1) {  // only the last key block can hold padding keys (T > 32 * (NKB - 1) is checked at launch)
          const int key = kb * kKeyBlock + nt * 8 + 2 * t;
          if (key >= T) s[kb][nt][0] = s[kb][nt][2] = -INFINITY;
          if (key + 1 >= T) s[kb][nt][1] = s[kb][nt][3] = -INFINITY;
        }
        m0 = fmaxf(m0, fmaxf(s[kb][nt][0], s[kb][nt][1]));
        m1 = fmaxf(m1, fmaxf(s[kb][nt][2], s[kb][nt][3]));
      }
    }
    m0 = quad_max(m0) * scale_log2e;
    m1 = quad_max(m1) * scale_log2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[kb][nt][0] = ex2(fmaf(s[kb][nt][0], scale_log2e, -m0));
        s[kb][nt][1] = ex2(fmaf(s[kb][nt][1], scale_log2e, -m0));
        s[kb][nt][2] = ex2(fmaf(s[kb][nt][2], scale_log2e, -m1));
        s[kb][nt][3] = ex2(fmaf(s[kb][nt][3], scale_log2e, -m1));
        l0 += s[kb][nt][0] + s[kb][nt][1];
        l1 += s[kb][nt][2] + s[kb][nt][3];
      }
    const float inv0 = 1.0f / quad_sum(l0), inv1 = 1.0f / quad_sum(l1);

    // P.V on the un-normalised exponentials (each in (0, 1]); 1/l is applied to the 16 x 32 output instead of
    // the 16 x T probabilities.  The probabilities themselves are only normalised when they are written out.
    float o[4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
    const int row0 = mt * 16 + g, row1 = row0 + 8;
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      if (probs != nullptr) {
        TP* pr = probs + ((size_t)(b * kHeads + h) * T) * pp;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int key = kb * kKeyBlock + nt * 8 + 2 * t;
          if (row0 < T) {
            if (key < T) store_prob<TP>(pr + (size_t)row0 * pp + key, s[kb][nt][0] * inv0);
            if (key + 1 < T) store_prob<TP>(pr + (size_t)row0 * pp + key + 1, s[kb][nt][1] * inv0);
          }
          if (row1 < T) {
            if (key < T) store_prob<TP>(pr + (size_t)row1 * pp + key, s[kb][nt][2] * inv1);
            if (key + 1 < T) store_prob<TP>(pr + (size_t)row1 * pp + key + 1, s[kb][nt][3] * inv1);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[kb][2 * j][0], s[kb][2 * j][1]);
        pa[1] = pack_bf16x2(s[kb][2 * j][2], s[kb][2 * j][3]);
        pa[2] = pack_bf16x2(s[kb][2 * j + 1][0], s[kb][2 * j + 1][1]);
        pa[3] = pack_bf16x2(s[kb][2 * j + 1][2], s[kb][2 * j + 1][3]);
        const uint32_t vaddr = v_lane + (kb * kKeyBlock + j * 16) * kPitch * 2;
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, vaddr + np * 32);
          mma_bf16_16816(o[2 * np], pa, vb[0], vb[1]);
          mma_bf16_16816(o[2 * np + 1], pa, vb[2], vb[3]);
        }
      }
    }
    __nv_bfloat16* orow0 = out + ((size_t)b * T + row0) * (kHeads * kHd) + h * kHd + 2 * t;
    __nv_bfloat16* orow1 = out + ((size_t)b * T + row1) * (kHeads * kHd) + h * kHd + 2 * t;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      if (row0 < T) *reinterpret_cast<uint32_t*>(orow0 + nd * 8) = pack_bf16x2(o[nd][0] * inv0, o[nd][1] * inv0);
      if (row1 < T) *reinterpret_cast<uint32_t*>(orow1 + nd * 8) = pack_bf16x2(o[nd][2] * inv1, o[nd][3] * inv1);
    }
  }
}

// Online-softmax variant for launches that do not return the probabilities (every layer but the last, and the
// last one too when the attention map is switched off).  A warp owns MT m16 query tiles and walks the key blocks
// once: only one 32-key block of scores is live at a time, so the rescaled running output (flash-attention
// recurrence) costs far fewer registers than the 16 x T strip of the kernel above.
//   MT = 1 (default): 72 registers, five CTAs per SM at 145 tokens - 0.156 ms per layer at batch 1024;
//   MT = 2: the K and V fragments of a block are fetched from shared memory once for two tiles (half the
//           ldmatrix traffic), but 128 registers leave three CTAs per SM - 0.166 ms, like the strip kernel.
// NW = warps per CTA: 5 at 145 tokens (10 tiles), 9 at 257 tokens (17 tiles in two rounds, 0.593 -> 0.505 ms).
// Q, K and V are staged by cp.async (0.163 -> 0.154 ms against staging through registers).
template <int NKB, int MT, int NW>
__global__ void __launch_bounds__(NW * 32, NW > 5 ? (MT == 1 ? 3 : 1) : (MT == 2 ? 3 : 5))
attention_kernel_online(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T,
                        float scale_log2e, int reverse, int attn_cp_async) {
  constexpr int Tp = NKB * kKeyBlock;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sk = sq + Tp * kPitch;
  __nv_bfloat16* sv = sk + Tp * kPitch;

  pdl_launch_dependents();
  pdl_wait();
  const int bid = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int b = bid / kHeads, h = bid % kHeads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  const __nv_bfloat16* base = qkv + (size_t)b * T * (3 * kHeads * kHd) + h * kHd;
  {
    // 16-byte chunks of Q, K and V go straight from global to shared memory (cp.async, zero-fill for the padding
    // rows): no register staging, no second pass of shared-memory stores
    constexpr int kTotal = 3 * Tp * 4;
    if (attn_cp_async) {
      for (int i = tid; i < kTotal; i += NW * 32) {
        const int c = i & 3, row = (i >> 2) % Tp, part = (i >> 2) / Tp;
        const int srow = row < T ? row : 0;
        cp_async_16(sq + (part * Tp + row) * kPitch + c * 8,
                    base + (size_t)srow * (3 * kHeads * kHd) + part * kHeads * kHd + c * 8, row < T ? 16u : 0u);
      }
      cp_async_commit();
      cp_async_wait<0>();
    } else {
      constexpr int kBatch = 12;                    // requests in flight per thread
      for (int i0 = 0; i0 < kTotal; i0 += kBatch * NW * 32) {
        uint4 v[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
          const int i = i0 + tid + k * NW * 32;
          const int c = i & 3, row = (i >> 2) % Tp, part = (i >> 2) / Tp;
          v[k] = make_uint4(0, 0, 0, 0);
          if (i < kTotal && row < T)
            v[k] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)row * (3 * kHeads * kHd) + part * kHeads * kHd) + c);
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
          const int i = i0 + tid + k * NW * 32;
          const int c = i & 3, row = (i >> 2) % Tp, part = (i >> 2) / Tp;
          if (i < kTotal) *reinterpret_cast<uint4*>(sq + (part * Tp + row) * kPitch + c * 8) = v[k];
        }
      }
    }
  }
  __syncthreads();

  const int mgroups = (T + 16 * MT - 1) / (16 * MT);
  const uint32_t q_lane = smem_u32(sq) + ((lane & 15) * kPitch + (lane >> 4) * 8) * 2;
  const uint32_t k_lane = smem_u32(sk) + ((lane & 7) * kPitch + (lane >> 3) * 8) * 2;
  const uint32_t v_lane = smem_u32(sv) + ((((lane >> 3) & 1) * 8 + (lane & 7)) * kPitch + (lane >> 4) * 8) * 2;

  for (int mp = warp; mp < mgroups; mp += NW) {
    uint32_t qa[MT][2][4];
#pragma unroll
    for (int u = 0; u < MT; ++u) {
      ldmatrix_x4(qa[u][0], q_lane + (mp * 16 * MT + u * 16) * kPitch * 2);
      ldmatrix_x4(qa[u][1], q_lane + (mp * 16 * MT + u * 16) * kPitch * 2 + 32);
    }
    float o[MT][4][4];
    float m[MT][2], l[MT][2];  // running max (already in exp2 units) and partial row sums: [tile][row g / g + 8]
#pragma unroll
    for (int u = 0; u < MT; ++u) {
      m[u][0] = m[u][1] = -INFINITY;
      l[u][0] = l[u][1] = 0.f;
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) o[u][nd][0] = o[u][nd][1] = o[u][nd][2] = o[u][nd][3] = 0.f;
    }

#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      float s[MT][4][4];
      {
        uint32_t kf[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) ldmatrix_x4(kf[nt], k_lane + (kb * kKeyBlock + nt * 8) * kPitch * 2);
#pragma unroll
        for (int u = 0; u < MT; ++u)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            s[u][nt][0] = s[u][nt][1] = s[u][nt][2] = s[u][nt][3] = 0.f;
            mma_bf16_16816(s[u][nt], qa[u][0], kf[nt][0], kf[nt][1]);
            mma_bf16_16816(s[u][nt], qa[u][1], kf[nt][2], kf[nt][3]);
          }
      }
#pragma unroll
      for (int u = 0; u < MT; ++u) {
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (kb == NKB - 1) {  // only the last key block can hold padding keys (checked at launch)
            const int key = kb * kKeyBlock + nt * 8 + 2 * t;
            if (key >= T) s[u][nt][0] = s[u][nt][2] = -INFINITY;
            if (key + 1 >= T) s[u][nt][1] = s[u][nt][3] = -INFINITY;
          }
          bm0 = fmaxf(bm0, fmaxf(s[u][nt][0], s[u][nt][1]));
          bm1 = fmaxf(bm1, fmaxf(s[u][nt][2], s[u][nt][3]));
        }
        // key 0 is always valid, so the running max is finite from the first block on
        const float n0 = fmaxf(m[u][0], quad_max(bm0) * scale_log2e);
        const float n1 = fmaxf(m[u][1], quad_max(bm1) * scale_log2e);
        const float a0 = ex2(m[u][0] - n0), a1 = ex2(m[u][1] - n1);
        m[u][0] = n0;
        m[u][1] = n1;
        float p0 = 0.f, p1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          s[u][nt][0] = ex2(fmaf(s[u][nt][0], scale_log2e, -n0));
          s[u][nt][1] = ex2(fmaf(s[u][nt][1], scale_log2e, -n0));
          s[u][nt][2] = ex2(fmaf(s[u][nt][2], scale_log2e, -n1));
          s[u][nt][3] = ex2(fmaf(s[u][nt][3], scale_log2e, -n1));
          p0 += s[u][nt][0] + s[u][nt][1];
          p1 += s[u][nt][2] + s[u][nt][3];
        }
        l[u][0] = fmaf(l[u][0], a0, p0);
        l[u][1] = fmaf(l[u][1], a1, p1);
        if (kb > 0) {
#pragma unroll
          for (int nd = 0; nd < 4; ++nd) {
            o[u][nd][0] *= a0;
            o[u][nd][1] *= a0;
            o[u][nd][2] *= a1;
            o[u][nd][3] *= a1;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {  // two 16-key k-steps per block
        uint32_t vb[2][4];
        const uint32_t vaddr = v_lane + (kb * kKeyBlock + j * 16) * kPitch * 2;
        ldmatrix_x4_trans(vb[0], vaddr);
        ldmatrix_x4_trans(vb[1], vaddr + 32);
#pragma unroll
        for (int u = 0; u < MT; ++u) {
          uint32_t pa[4];
          pa[0] = pack_bf16x2(s[u][2 * j][0], s[u][2 * j][1]);
          pa[1] = pack_bf16x2(s[u][2 * j][2], s[u][2 * j][3]);
          pa[2] = pack_bf16x2(s[u][2 * j + 1][0], s[u][2 * j + 1][1]);
          pa[3] = pack_bf16x2(s[u][2 * j + 1][2], s[u][2 * j + 1][3]);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            mma_bf16_16816(o[u][2 * np], pa, vb[np][0], vb[np][1]);
            mma_bf16_16816(o[u][2 * np + 1], pa, vb[np][2], vb[np][3]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < MT; ++u) {
      const float inv0 = 1.0f / quad_sum(l[u][0]), inv1 = 1.0f / quad_sum(l[u][1]);
      const int row0 = mp * 16 * MT + u * 16 + g, row1 = row0 + 8;
      __nv_bfloat16* orow0 = out + ((size_t)b * T + row0) * (kHeads * kHd) + h * kHd + 2 * t;
      __nv_bfloat16* orow1 = out + ((size_t)b * T + row1) * (kHeads * kHd) + h * kHd + 2 * t;
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        if (row0 < T) *reinterpret_cast<uint32_t*>(orow0 + nd * 8) = pack_bf16x2(o[u][nd][0] * inv0, o[u][nd][1] * inv0);
        if (row1 < T) *reinterpret_cast<uint32_t*>(orow1 + nd * 8) = pack_bf16x2(o[u][nd][2] * inv1, o[u][nd][3] * inv1);
      }
    }
  }
}

}  // namespace

// Q, K and V of one (image, head) live in shared memory as three [Tp][kPitch] bf16 arrays.
int attention_max_tokens() { return (227 * 1024 / (3 * kPitch * 2)) / kKeyBlock * kKeyBlock; }
bool attention_tokens_supported(int T) { return T >= 1 && T <= attention_max_tokens(); }

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, void* attn_probs, int probs_dtype, int B, int T,
                     cudaStream_t stream, int reverse, int probs_pitch) {
  const int pp = probs_pitch > 0 ? probs_pitch : T;
  const int Tp = (T + kKeyBlock - 1) / kKeyBlock * kKeyBlock;
  const size_t smem = (size_t)3 * Tp * kPitch * 2;
  if (smem > 227 * 1024) {
    set_error("attention: %d tokens do not fit one CTA's shared memory", T);
    return -1;
  }
  // softmax(x * d^-0.5) evaluated as exp2((x - max) * d^-0.5 * log2(e))
  const float scale_log2e = 0.17677669529663687f * 1.4426950408889634f;
  const unsigned grid = (unsigned)B * kHeads;
  if (attn_probs == nullptr && attention_tc_enabled() && attention_tc_supported(T))
    return launch_attention_tc(qkv, out, B, T, scale_log2e, device_sm_count(), stream, reverse);
  if (attn_probs == nullptr && attention_online_enabled()) {
    // no probabilities to return: online-softmax kernel, K/V fragments shared by two query tiles per warp
    const int nkb = Tp / kKeyBlock;
    const bool one = attention_tiles_per_warp() == 1;  // 1: fewer registers, five CTAs per SM; 2: shared K/V fragments
    if (nkb == 5) {
      HGR_CHECK_CUDA(launch_pdl(one ? attention_kernel_online<5, 1, kWarps> : attention_kernel_online<5, 2, kWarps>,
                                dim3(grid), dim3(kWarps * 32), smem, stream, qkv, out, T, scale_log2e, reverse,
                                attention_cp_async_enabled() ? 1 : 0));
      return 0;
    }
    if (nkb == 9) {
      // 257 tokens = 17 query tiles: nine warps walk them in two rounds (five warps need four)
      auto kern = one ? attention_kernel_online<9, 1, 9> : attention_kernel_online<9, 2, 9>;
      HGR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      HGR_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(9 * 32), smem, stream, qkv, out, T, scale_log2e, reverse,
                                attention_cp_async_enabled() ? 1 : 0));
      return 0;
    }
  }
  if (T <= 5 * kKeyBlock && T > 4 * kKeyBlock) {
    const size_t smem1 = (size_t)3 * 5 * kKeyBlock * kPitch * 2;
    if (attn_probs != nullptr && probs_dtype == DT_BF16)
      HGR_CHECK_CUDA(launch_pdl(attention_kernel_1pass<__nv_bfloat16, 5>, dim3(grid), dim3(kWarps * 32), smem1, stream, qkv,
                                out, static_cast<__nv_bfloat16*>(attn_probs), T, scale_log2e, reverse, pp));
    else
      HGR_CHECK_CUDA(launch_pdl(attention_kernel_1pass<float, 5>, dim3(grid), dim3(kWarps * 32), smem1, stream, qkv, out,
                                static_cast<float*>(attn_probs), T, scale_log2e, reverse, pp));
    return 0;
  }
  if (attn_probs != nullptr && probs_dtype == DT_BF16) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    HGR_CHECK_CUDA(launch_pdl(attention_kernel<__nv_bfloat16>, dim3(grid), dim3(kWarps * 32), smem, stream, qkv, out,
                              static_cast<__nv_bfloat16*>(attn_probs), T, Tp, scale_log2e, pp));
  } else {
    HGR_CHECK_CUDA(
        cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HGR_CHECK_CUDA(launch_pdl(attention_kernel<float>, dim3(grid), dim3(kWarps * 32), smem, stream, qkv, out,
                              static_cast<float*>(attn_probs), T, Tp, scale_log2e, pp));
  }
  return 0;
}

}  // namespace hgr
