// Attention core on the 5th-gen tensor cores for the 145-token case (192 x 192 inputs)
// (reference model/transformer.py:66-74: softmax(q k^T * d^-0.5) v per head, heads of width 32).
//
// The mma.sync kernels of attention.cu spend their time feeding the legacy tensor path (ldmatrix traffic, one CTA
// per (image, head), ~4 700 SM cycles per head).  Here a persistent CTA walks (image, head pair) items:
//   * TMA brings the Q, K and V columns of TWO heads (64 columns = one 128-byte swizzle row) as three
//     [160 tokens x 64] boxes; rows beyond the image's 145 tokens are zero-filled by the tensor map;
//   * S = Q K^T is one M = 128, N = 160, K = 32 tcgen05.mma per (head, 128-query tile) with the fp32 scores in TMEM;
//   * softmax warps (one thread per query row) read the scores twice out of TMEM - row maximum, then
//     exp2 / row sum / bf16 - and write P into shared memory as the K-major SWIZZLE_128B A operand of the next MMA;
//   * O = P V is M = 128, N = 32, K = 160 with V as the MN-major B operand (the [tokens x 64] box as it was loaded:
//     no transpose), accumulated in TMEM; the softmax warps scale by 1 / sum and store 'b n (h d)'.
// The probabilities are not returned by this kernel: launches that need them (the last layer with
// return_attention=True) use attention.cu.
//
// Status (round 1): parity-green (tests/test_gpu_ops.py::test_attention_tc, 1.9e-3 rel-L2 like the mma.sync kernels),
// 0.223 ms per layer with one softmax warp per lane quarter, 0.170-0.184 ms with two (this version) against 0.156-0.171
// ms for the mma.sync online-softmax kernel on the same boxes - so it is OPT-IN (HGR_ATTN_TC=1).  What bounds it: a
// unit's chain S -> max -> exp -> P -> P V -> O is ~4 500 cycles of mostly latency, only two units fit TMEM
// (2 x 160 score columns + outputs), and the 17-row second query tile keeps one lane quarter busy while three idle.
// Keeping a row's 80 scores in registers across the two passes spilled (0.233 ms).
//
// Units (head hh of the pair, query tile mt) are issued in the order mt-major, so softmax group g (8 warps, two per TMEM
// lane quarter) always owns head hh = g and both groups see the same mix of full (rows 0-127) and short
// (rows 128-144) tiles.  TMEM: two score buffers of 160 columns and two output buffers of 32 columns.
#include <cstdio>
#include <cstring>

#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 640;  // 4 service warps + 2 softmax groups of 8 warps
constexpr int kTp = 160;                          // padded tokens (keys per MMA, rows per box)
constexpr int kHeads = 8;
constexpr int kBoxBytes = kTp * 128;              // [160 rows][64 cols] bf16
constexpr int kStageBytes = 3 * kBoxBytes;        // Q | K | V of one head pair
constexpr int kPChunkBytes = 128 * 128;           // [128 queries][64 keys] bf16
constexpr int kPBytes = 3 * kPChunkBytes;         // keys 0-63 | 64-127 | 128-159 (+ unused tail)
constexpr int kOffQkv = 0;
constexpr int kOffP = 2 * kStageBytes;
constexpr int kOffExch = kOffP + 2 * kPBytes;    // per group: row maxima and row sums of the two key halves, 2 KiB
constexpr int kOffBars = kOffExch + 2 * 2048;
constexpr int kNumBars = 10;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
constexpr int kSCols = 160;                       // TMEM columns of one score buffer
constexpr int kOBase = 2 * kSCols;                // output buffers start here, 32 columns each
static_assert(kBoxBytes % 1024 == 0 && kOffP % 1024 == 0, "operand tiles need 1024-byte alignment");
static_assert(kSmemBytes <= 227 * 1024, "attention_tc shared-memory plan exceeds one CTA");

// MN-major, 128-byte-swizzled operand: rows of the K dimension are 128 B apart, 8-row groups `sbo_bytes` apart; the
// 64 MN elements of a row are contiguous (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>:
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).  N <= 64 here, so the leading byte offset is never used.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;  // LBO (unused), non-zero like CUTLASS emits
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnTcParams {
  __nv_bfloat16* out;  // (B, T, 256)
  int B, T;
  float scale_log2e;
  int reverse;
};

__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQkv, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* qkv_full = bars;       // [2]
  uint64_t* qkv_empty = bars + 2;  // [2]
  uint64_t* s_full = bars + 4;     // [2] scores of a unit are in TMEM
  uint64_t* p_ready = bars + 6;    // [2] probabilities of a unit are in shared memory (128 arrivals)
  uint64_t* o_full = bars + 8;     // [2] P V of a unit is in TMEM
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmQkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qkv_full[i], 1);
      mbar_init(&qkv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  const int total_items = p.B * (kHeads / 2);  // (image, head pair)
  const int my_items = (int)blockIdx.x < total_items ? (total_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int first = p.reverse ? total_items - 1 - (int)blockIdx.x : (int)blockIdx.x;
  const int step = p.reverse ? -(int)gridDim.x : (int)gridDim.x;

  if (warp == 0) {
    // ================= TMA producer: Q | K | V boxes of one head pair per item =================
    if (elect_one_sync()) {
      int item = first;
      for (int i = 0; i < my_items; ++i, item += step) {
        const int s = i & 1;
        const int b = item >> 2, pair = item & 3;
        mbar_wait(&qkv_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qkv_full[s], kStageBytes);
#pragma unroll
        for (int part = 0; part < 3; ++part)
          tma_load_5d(smem + kOffQkv + s * kStageBytes + part * kBoxBytes, &tmQkv, &qkv_full[s], part * 256 + pair * 64,
                      0, b, 0, 0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // unit u = 4 i + 2 mt + hh:  S(u) = Q[mt] K^T into score buffer u & 1, then P V of unit u - 1
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kTp);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);  // B is MN-major
    auto issue_pv = [&](int v) {
      const int g = v & 1;            // = hh
      const int s = (v >> 2) & 1;     // shared-memory stage of the unit's item
      mbar_wait(&p_ready[g], (v >> 1) & 1);
      tc_fence_after();
      const uint32_t pbuf = smem_u32(smem + kOffP + g * kPBytes);
      const uint32_t vbuf = smem_u32(smem + kOffQkv + s * kStageBytes + 2 * kBoxBytes) + g * 64;
      const uint32_t tmem_d = tmem_base + kOBase + g * 32;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < kTp / 16; ++k) {
          const uint64_t a = umma_desc_sw128(pbuf + (k >> 2) * kPChunkBytes, 1024) + 2 * (k & 3);
          const uint64_t bdesc = umma_desc_mn_sw128(vbuf + k * 2048, 1024);
          umma_bf16_ss(tmem_d, a, bdesc, idesc_o, k != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[g]);
        if ((v & 3) == 3) umma_commit(&qkv_empty[s]);  // last unit of the item: Q, K, V may be overwritten
      }
      __syncwarp();
    };
    int u = 0;
    for (int i = 0; i < my_items; ++i) {
      const int s = i & 1;
      mbar_wait(&qkv_full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t qbuf = smem_u32(smem + kOffQkv + s * kStageBytes);
      const uint32_t kbuf = qbuf + kBoxBytes;
#pragma unroll
      for (int r = 0; r < 4; ++r, ++u) {
        const int mt = r >> 1, hh = r & 1;
        // the score buffer u & 1 is free: P V of unit u - 2 was issued after its probabilities were complete
        const uint32_t tmem_d = tmem_base + (u & 1) * kSCols;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t a = umma_desc_sw128(qbuf + mt * 128 * 128 + hh * 64, 1024) + 2 * k;
            const uint64_t bdesc = umma_desc_sw128(kbuf + hh * 64, 1024) + 2 * k;
            umma_bf16_ss(tmem_d, a, bdesc, idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[u & 1]);
        }
        __syncwarp();
        if (u >= 1) issue_pv(u - 1);
      }
    }
    if (u >= 1) issue_pv(u - 1);
  } else if (warp >= 4) {
    // ================= softmax groups: group g (8 warps) owns head hh = g of every pair; the two warps of a TMEM
    // lane quarter split the 160 keys of a row (80 each) and exchange row maximum and row sum through shared memory,
    // so four softmax warps per scheduler keep the MUFU pipe (one ex2 per score) busy =================
    const int e4 = warp - 4;
    const int q = e4 & 3;
    const int half = (e4 >> 2) & 1;
    const int g = e4 >> 3;
    const int lrow = q * 32 + lane;  // row inside the 128-query tile == TMEM lane
    const uint32_t sw = static_cast<uint32_t>(lrow & 7);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* pbuf = smem + kOffP + g * kPBytes + lrow * 128;
    float* xmax = reinterpret_cast<float*>(smem + kOffExch) + g * 512;  // [2 halves][128 rows]
    float* xsum = xmax + 256;
    const uint32_t bar_id = 1 + g;
    const float c = p.scale_log2e;
    const int col_base = half * 80;
    int item = first;
    int n = 0;  // units this group has processed
    for (int i = 0; i < my_items; ++i, item += step) {
      const int b = item >> 2, pair = item & 3;
      const int h = pair * 2 + g;
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt, ++n) {
        const uint32_t ph = n & 1;
        const int row = mt * 128 + lrow;
        const bool active = mt == 0 || q == 0;  // rows 128-159 live in lane quarter 0 of the second tile
        mbar_wait(&s_full[g], ph);
        tc_fence_after();
        const uint32_t s_addr = t_lane + g * kSCols + col_base;
        float m = -INFINITY, l = 0.f;
        if (active) {
          // ---- pass 1: maximum over this warp's 80 keys; all five TMEM loads in flight, one wait ----
          {
            uint32_t acc[80];
#pragma unroll
            for (int cb = 0; cb < 5; ++cb) tmem_ld_32x32b_x16(s_addr + cb * 16, acc + cb * 16);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 80; ++e) {
              const float v = __uint_as_float(acc[e]);
              if (half == 0 || col_base + e < p.T) m = fmaxf(m, v);  // T > 128: keys 0-79 are always valid
            }
          }
          xmax[half * 128 + lrow] = m;
        }
        bar_sync(bar_id, 256);
        if (active) {
          m = fmaxf(m, xmax[(half ^ 1) * 128 + lrow]) * c;
          // ---- pass 2: exponentials, partial row sum, bf16 probabilities into the A-operand layout; the next
          // block's scores are on their way out of TMEM while this block is computed ----
          uint32_t accb[2][16];
          tmem_ld_32x32b_x16(s_addr, accb[0]);
#pragma unroll
          for (int cb = 0; cb < 5; ++cb) {
            const uint32_t(&acc)[16] = accb[cb & 1];
            tmem_ld_wait();
            if (cb < 4) tmem_ld_32x32b_x16(s_addr + (cb + 1) * 16, accb[(cb + 1) & 1]);
            uint32_t packed[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float p0 = ex2f(fmaf(__uint_as_float(acc[e]), c, -m));
              float p1 = ex2f(fmaf(__uint_as_float(acc[e + 1]), c, -m));
              if (half == 1) {
                if (col_base + cb * 16 + e >= p.T) p0 = 0.f;
                if (col_base + cb * 16 + e + 1 >= p.T) p1 = 0.f;
              }
              l += p0 + p1;
              packed[e >> 1] = pack_bf16x2(p0, p1);
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const int col0 = col_base + cb * 16 + v * 8;
              const uint32_t piece = static_cast<uint32_t>((col0 & 63) >> 3) ^ sw;
              *reinterpret_cast<uint4*>(pbuf + (col0 >> 6) * kPChunkBytes + piece * 16) =
                  make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
            }
          }
          xsum[half * 128 + lrow] = l;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_ready[g]);
        // ---- output: O / l -> 'b n (h d)', 16 of the 32 columns per warp ----
        mbar_wait(&o_full[g], ph);
        tc_fence_after();
        bar_sync(bar_id, 256);  // both halves' partial sums are visible
        if (active) {
          const float inv = 1.0f / (xsum[lrow] + xsum[128 + lrow]);
          uint32_t o[16];
          tmem_ld_32x32b_x16(t_lane + kOBase + g * 32 + half * 16, o);
          tmem_ld_wait();
          if (row < p.T) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.T + row) * (kHeads * 32) + h * 32 + half * 16);
#pragma unroll
            for (int v = 0; v < 2; ++v)
              dst[v] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * v]) * inv, __uint_as_float(o[8 * v + 1]) * inv),
                                  pack_bf16x2(__uint_as_float(o[8 * v + 2]) * inv, __uint_as_float(o[8 * v + 3]) * inv),
                                  pack_bf16x2(__uint_as_float(o[8 * v + 4]) * inv, __uint_as_float(o[8 * v + 5]) * inv),
                                  pack_bf16x2(__uint_as_float(o[8 * v + 6]) * inv, __uint_as_float(o[8 * v + 7]) * inv));
          }
        }
        tc_fence_before();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool attention_tc_supported(int T) { return T > 128 && T <= kTp; }

int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, float scale_log2e, int num_sms,
                        cudaStream_t stream, int reverse) {
  if (!attention_tc_supported(T)) {
    set_error("attention_tc: built for 129..160 tokens, got %d", T);
    return -1;
  }
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap tm;
  {
    const uint64_t dims[5] = {768, (uint64_t)T, (uint64_t)B, 1, 1};
    const uint64_t row = 768 * 2;
    const uint64_t strides[4] = {row, row * T, row * T * B, row * T * B};
    const uint32_t box[5] = {64, (uint32_t)kTp, 1, 1, 1};
    if (int r = make_tensor_map_bf16(&tm, qkv, 5, dims, strides, box)) return r;
  }
  AttnTcParams p;
  p.out = out;
  p.B = B;
  p.T = T;
  p.scale_log2e = scale_log2e;
  p.reverse = reverse;
  const int items = B * (kHeads / 2);
  const int grid = items < num_sms ? items : num_sms;
  if (grid <= 0) return 0;
  HGR_CHECK_CUDA(launch_pdl(attention_tc_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, tm, p));
  return 0;
}

}  // namespace hgr
