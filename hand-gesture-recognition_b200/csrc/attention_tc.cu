// Attention core on the 5th-gen tensor cores for the 145-token case (192 x 192 inputs)
// (reference model/transformer.py:66-74: softmax(q k^T * d^-0.5) v per head, heads of width 32).
//
// A persistent CTA walks (image, head pair) items.  Per item:
//   * TMA brings the Q, K and V columns of TWO heads (64 columns = one 128-byte swizzle row): Q rows 0-127, K and V as
//     [160 tokens x 64] boxes whose rows beyond the image's 145 tokens are zero-filled by the tensor map;
//   * S = Q K^T is an M = 128, N = 160, K = 32 tcgen05.mma per unit with the fp32 scores in TMEM;
//   * a softmax group (4 warps, ONE THREAD PER QUERY ROW, no cross-thread exchange) reads the scores twice out of
//     TMEM - row maximum, then exp2 / row sum - and writes the bf16 probabilities back INTO TENSOR MEMORY over the
//     scores it has consumed (tcgen05.st, two keys per 32-bit column);
//   * O = P V has P as the TMEM A operand and V as the MN-major shared-memory B operand (the [tokens x 64] box as it
//     was loaded: no transpose); the group scales by 1 / sum and stores 'b n (h d)'.
// The probabilities never touch shared memory, so a unit costs exactly 160 TMEM columns (scores, then P in columns
// 0-79 and O behind them) and THREE units are in flight (3 x 160 of the 512 columns), one per softmax group.
//
// The 145 query rows of a head are a full tile (rows 0-127) and a 17-row tile.  A warp can only read its own quarter
// of the TMEM lanes and lives on the scheduler of the same index, so a unit made of ONE short tile would keep one
// warp of its group busy and three idle (and that warp's MUFU pipe - one ex2 per score, the bound of this kernel -
// would carry the whole unit).  The two short tiles of an item therefore share ONE unit: a fourth box holds the
// rows 128.. of both heads at a NEGATIVE row coordinate of a tensor map that starts at row 128, so that everything
// around the 17 real rows is zero-filled by TMA; head 0's MMA starts its A operand 32 rows into that box, head 1's at
// its first row, both ACCUMULATE into the same 160 columns: lane quarter j0 receives head 0's rows, quarter j0 + 1
// head 1's, every other lane adds zeros.  Their P V is one N = 64 MMA chain against the whole V box (both heads'
// columns); each quarter reads its own head's 32 output columns.  An item is three units (full head 0, full head 1, the two
// short tiles), rotated over the three groups, and j0 alternates between 0 and 2, so that the four schedulers carry
// the same load.
//
// The probabilities are not returned by this kernel: launches that need them (the last layer with
// return_attention=True) use attention.cu, and so do token counts outside 129..160.
//
// Measured on B200 (tools/tmem_probe.cu, tools/mufu_probe.cu): tcgen05.ld.32x32b.x32 sustains 173 B/clk per warp and
// ~800 B/clk per SM, so reading the scores twice costs ~200 cycles per unit; MUFU.EX2 issues one warp instruction per
// 8 cycles per scheduler (9.3 inside the softmax instruction mix), a lone warp reaches 11.2; the TMEM-A MMA reproduces
// P V exactly.  Round 1's two-unit variant with P in shared memory ran at 0.17-0.18 ms per layer, the first three-unit
// version (short tiles as units of their own, one thread per row) at 0.138 ms, with the combined short unit (this
// version) at 0.108 ms, mma.sync at 0.150 ms.  A variant with two warps per lane quarter splitting a row's keys (24
// softmax warps, maximum / sum exchanged through shared memory) was SLOWER (0.122-0.133 ms): the cycle timeline
// (hgr_attention_tc_trace, tools/attention_trace.py) shows the MUFU pipe saturated only while two or three groups are
// in their exponential pass at the same time (40 % of a unit's life); the maximum pass, the MMA hand-over and the
// output pass of a unit are stretched by the other groups' queued MUFU work, and more warps add queueing, not
// throughput.  Next step: two softmax groups over three buffers with the output drained by separate
// warps, so that a group never waits for its own hand-over.
#include <cstdio>
#include <cstring>

#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kBufs = 3;                          // TMEM buffers = units in flight = units per item
constexpr int kSoftGroups = 2;                    // softmax groups of four warps (one thread per query row)
constexpr int kThreads = (8 + 4 * kSoftGroups) * 32;  // 4 service warps + 4 drain warps + the softmax groups
constexpr int kTp = 160;                          // padded tokens (keys per MMA, rows per K / V box)
constexpr int kHeads = 8;
// Q, K and the short-row box are dead once the three score MMAs of their item have retired, V lives until the item's
// last P V: two rings, so that the (long) loads of item i + 2 start while item i's probabilities are still computed
constexpr int kStages = 2;                        // Q | K | short-row boxes
constexpr int kVStages = 4;                       // V boxes
constexpr int kQBytes = 128 * 128;                // Q rows 0-127:  [128 rows][64 cols] bf16
constexpr int kBoxBytes = kTp * 128;              // K, V, short-row box: [160 rows][64 cols] bf16
constexpr int kOffK = kQBytes;
constexpr int kOffQs = kOffK + kBoxBytes;         // rows 128.. of both heads inside a zero-filled box
constexpr int kStageBytes = kOffQs + kBoxBytes;
constexpr int kOffVRing = kStages * kStageBytes;
constexpr int kOffSum = kOffVRing + kVStages * kBoxBytes;  // row sums of the units in flight: [kBufs][128] fp32
constexpr int kOffBars = kOffSum + kBufs * 128 * 4;
constexpr int kNumBars = 2 * kStages + 2 * kVStages + 4 * kBufs;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
constexpr int kUnitCols = 160;  // TMEM columns of one unit: S [0,160), then P [0,80) and O behind it
constexpr int kOColFull = 128;  // full tile: O (N = 32) in columns [128,160)
constexpr int kOColShort = 80;  // short tiles: O (N = 64) in columns [80,144)
static_assert(kQBytes % 1024 == 0 && kBoxBytes % 1024 == 0, "operand tiles need 1024-byte alignment");
static_assert(kSmemBytes <= 227 * 1024, "attention_tc shared-memory plan exceeds one CTA");
static_assert(kBufs * kUnitCols <= 512, "units in flight exceed tensor memory");

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

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

struct AttnTcParams {
  __nv_bfloat16* out;  // (B, T, 256)
  int B, T;
  float scale_log2e;
  int reverse;
  // optional timeline of CTA 0 (hgr_attention_tc_trace): clock64 at the hand-over points, [item][warp][8 marks];
  // softmax warps (8..): scores ready, maximum done, probabilities written at marks 0-2 (marks 4-6 for a group's
  // second unit of the same item); drain warps (4-7): output ready / stored at marks 2 r, 2 r + 1 of position r;
  // MMA warp (1): scores issued at marks 0-2 (position r), P V issued at marks 3-5
  long long* trace;
  int trace_items;
};

// Unit `r` (0..2, == its TMEM buffer == its softmax group) of the CTA's i-th item: the types rotate with the item so
// that every group sees full and short units alike.  type 0 / 1: rows 0-127 of head 0 / 1; type 2: both short tiles.
__device__ __forceinline__ int unit_type(int i, int r) {
  const int t = r + i % 3;
  return t >= 3 ? t - 3 : t;
}
// lane quarter of head 0's short rows in the i-th item (head 1's are in the next quarter)
__device__ __forceinline__ int short_quarter(int i) { return (i & 1) * 2; }

// kTokens > 0: token count known at compile time (the padded keys cost neither masks nor exponentials); 0: p.T
template <int kTokens>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmQs, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T = kTokens > 0 ? kTokens : p.T;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* qkv_full = bars;              // [kStages]
  uint64_t* qkv_empty = bars + kStages;   // [kStages]
  uint64_t* v_full = bars + 2 * kStages;              // [kVStages]
  uint64_t* v_empty = v_full + kVStages;              // [kVStages]
  uint64_t* s_full = v_empty + kVStages;  // [kBufs] scores of a unit are in TMEM
  uint64_t* p_ready = s_full + kBufs;     // [kBufs] probabilities are in TMEM, row sums in shared memory (128 arrivals)
  uint64_t* o_full = p_ready + kBufs;     // [kBufs] P V of a unit is in TMEM
  uint64_t* buf_free = o_full + kBufs;    // [kBufs] the unit's output has been read (128 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  float* lsum = reinterpret_cast<float*>(smem + kOffSum);  // [kBufs][128] row sums, softmax -> drain

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmKV);
    prefetch_tensormap(&tmQs);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&qkv_full[i], 1);
      mbar_init(&qkv_empty[i], 1);
    }
    for (int i = 0; i < kVStages; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < kBufs; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&buf_free[i], 128);
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
  const int num_units = kBufs * my_items;  // unit u = (item u / 3, position / TMEM buffer u % 3)

  if (warp == 0) {
    // ================= TMA producer: Q | K | V | short-row boxes of one head pair per item =================
    if (elect_one_sync()) {
      int item = first;
      for (int i = 0; i < my_items; ++i, item += step) {
        const int b = item >> 2, pair = item & 3;
        const int s = i % kStages, sv = i % kVStages;
        uint8_t* stage = smem + s * kStageBytes;
        mbar_wait(&qkv_empty[s], ((i / kStages) & 1) ^ 1);
        mbar_expect_tx(&qkv_full[s], kStageBytes);
        tma_load_5d(stage, &tmQ, &qkv_full[s], pair * 64, 0, b, 0, 0);
        tma_load_5d(stage + kOffK, &tmKV, &qkv_full[s], 256 + pair * 64, 0, b, 0, 0);
        // box row r holds token 128 + r - 32 (j0 + 1); everything outside [128, T) is zero-filled
        tma_load_5d(stage + kOffQs, &tmQs, &qkv_full[s], pair * 64, -32 * (short_quarter(i) + 1), b, 0, 0);
        mbar_wait(&v_empty[sv], ((i / kVStages) & 1) ^ 1);
        mbar_expect_tx(&v_full[sv], kBoxBytes);
        tma_load_5d(smem + kOffVRing + sv * kBoxBytes, &tmKV, &v_full[sv], 512 + pair * 64, 0, b, 0, 0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the scores of the first three units, then P V of unit v and the scores of the
    // unit that inherits its buffer (v + 3), in the order the softmax groups finish =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kTp);
    constexpr uint32_t idesc_o32 = umma_idesc_bf16(128, 32) | (1u << 16);  // B is MN-major
    constexpr uint32_t idesc_o64 = umma_idesc_bf16(128, 64) | (1u << 16);
    int i_s = 0, r_s = 0;  // item / position of the unit whose scores are issued next
    auto issue_scores = [&]() {
      const int st = i_s % kStages;
      if (r_s == 0) {
        mbar_wait(&qkv_full[st], (i_s / kStages) & 1);
        tc_fence_after();
      }
      if (i_s >= 1) {
        // the previous unit of this buffer has been drained (its O has been read, which implies its P V has retired)
        mbar_wait(&buf_free[r_s], (i_s - 1) & 1);
        tc_fence_after();
      }
      const uint32_t stage = smem_u32(smem + st * kStageBytes);
      const uint32_t tmem_d = tmem_base + r_s * kUnitCols;
      const int ty = unit_type(i_s, r_s);
      if (elect_one_sync()) {
        if (ty < 2) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t a = umma_desc_sw128(stage + ty * 64, 1024) + 2 * k;
            const uint64_t bdesc = umma_desc_sw128(stage + kOffK + ty * 64, 1024) + 2 * k;
            umma_bf16_ss(tmem_d, a, bdesc, idesc_s, k != 0 ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            // head 0: its rows are 32 (j0 + 1) rows into the box and belong in quarter j0 -> start 32 rows in;
            // head 1: quarter j0 + 1 -> start at the box's first row
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t a = umma_desc_sw128(stage + kOffQs + (hh == 0 ? 32 * 128 : 0) + hh * 64, 1024) + 2 * k;
              const uint64_t bdesc = umma_desc_sw128(stage + kOffK + hh * 64, 1024) + 2 * k;
              umma_bf16_ss(tmem_d, a, bdesc, idesc_s, (hh | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&s_full[r_s]);
        // the item's last score MMA: its Q, K and short-row boxes may be overwritten (not signalled when no later
        // item will use the stage, so that no arrive is in flight when the CTA exits)
        if (r_s == kBufs - 1 && i_s + kStages < my_items) umma_commit(&qkv_empty[st]);
        if (p.trace != nullptr && blockIdx.x == 0 && i_s < p.trace_items)
          p.trace[((size_t)i_s * (kThreads / 32) + 1) * 8 + r_s] = clock64();
      }
      __syncwarp();
      if (++r_s == kBufs) {
        r_s = 0;
        ++i_s;
      }
    };
    for (int u = 0; u < kBufs && u < num_units; ++u) issue_scores();
    int i_o = 0, r_o = 0;
    for (int v = 0; v < num_units; ++v) {
      const int sv = i_o % kVStages;
      if (r_o == 0) mbar_wait(&v_full[sv], (i_o / kVStages) & 1);
      mbar_wait(&p_ready[r_o], i_o & 1);
      tc_fence_after();
      const uint32_t vbuf = smem_u32(smem + kOffVRing + sv * kBoxBytes);
      const uint32_t tmem_u = tmem_base + r_o * kUnitCols;
      const int ty = unit_type(i_o, r_o);
      if (elect_one_sync()) {
        if (ty < 2) {
#pragma unroll
          for (int k = 0; k < kTp / 16; ++k)
            umma_bf16_ts(tmem_u + kOColFull, tmem_u + 8 * k, umma_desc_mn_sw128(vbuf + ty * 64 + k * 2048, 1024),
                         idesc_o32, k != 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < kTp / 16; ++k)
            umma_bf16_ts(tmem_u + kOColShort, tmem_u + 8 * k, umma_desc_mn_sw128(vbuf + k * 2048, 1024), idesc_o64,
                         k != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[r_o]);
        // last unit of the item: its V box may be overwritten
        if (r_o == kBufs - 1 && i_o + kVStages < my_items) umma_commit(&v_empty[sv]);
        if (p.trace != nullptr && blockIdx.x == 0 && i_o < p.trace_items)
          p.trace[((size_t)i_o * (kThreads / 32) + 1) * 8 + 3 + r_o] = clock64();
      }
      __syncwarp();
      if (++r_o == kBufs) {
        r_o = 0;
        ++i_o;
      }
      if (v + kBufs < num_units) issue_scores();
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= drain warps (one per lane quarter): O / l -> 'b n (h d)', in unit order =================
    const int q = warp & 3;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int item = first;
    for (int i = 0; i < my_items; ++i, item += step) {
      const int b = item >> 2, pair = item & 3;
      const int j0 = short_quarter(i);
      const uint32_t ph = i & 1;
#pragma unroll 1
      for (int r = 0; r < kBufs; ++r) {
        const int ty = unit_type(i, r);
        const bool active = ty < 2 || q == j0 || q == j0 + 1;
        const int hh = ty < 2 ? ty : q - j0;
        const int h = pair * 2 + hh;
        const int row = ty < 2 ? q * 32 + lane : 128 + lane;
        const uint32_t ocol = ty < 2 ? kOColFull : kOColShort + 32 * hh;
        mbar_wait(&p_ready[r], ph);  // acquires the row sums the softmax threads left in shared memory
        mbar_wait(&o_full[r], ph);
        tc_fence_after();
        long long* tr = (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && i < p.trace_items)
                            ? p.trace + ((size_t)i * (kThreads / 32) + warp) * 8 : nullptr;
        if (tr) tr[2 * r] = clock64();
        if (active) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(t_lane + r * kUnitCols + ocol, o);
          const float l = lsum[r * 128 + q * 32 + lane];
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&buf_free[r]);  // the registers hold the output: the MMA warp may reuse the buffer
          if (row < T) {
            const float inv = 1.0f / l;
            uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * T + row) * (kHeads * 32) + h * 32);
#pragma unroll
            for (int v = 0; v < 4; ++v)
              dst[v] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * v]) * inv, __uint_as_float(o[8 * v + 1]) * inv),
                                  pack_bf16x2(__uint_as_float(o[8 * v + 2]) * inv, __uint_as_float(o[8 * v + 3]) * inv),
                                  pack_bf16x2(__uint_as_float(o[8 * v + 4]) * inv, __uint_as_float(o[8 * v + 5]) * inv),
                                  pack_bf16x2(__uint_as_float(o[8 * v + 6]) * inv, __uint_as_float(o[8 * v + 7]) * inv));
          }
        } else {
          tc_fence_before();
          mbar_arrive(&buf_free[r]);
        }
        if (tr) tr[2 * r + 1] = clock64();
      }
    }
  } else if (warp >= 8) {
    // ================= softmax groups: group G takes every second unit (u = G, G + 2, ...), whatever buffer it is
    // in; it never waits for a P V or an output: as soon as its probabilities are written it moves on =================
    const int e8 = warp - 8;
    const int q = e8 & 3;  // == warp % 4: the TMEM lane quarter this warp may access
    const int grp = e8 >> 2;
    const float c = p.scale_log2e;
    const int nlast = T - 128;  // valid keys of the last 32-key chunk
    int i = 0, r = grp;         // item and position of this group's current unit
    for (int u = grp; u < num_units; u += kSoftGroups) {
      const int ty = unit_type(i, r);
      const int j0 = short_quarter(i);
      const bool active = ty < 2 || q == j0 || q == j0 + 1;
      const uint32_t t_unit = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + r * kUnitCols;
      mbar_wait(&s_full[r], i & 1);
      tc_fence_after();
      long long* tr = (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && i < p.trace_items)
                          ? p.trace + ((size_t)i * (kThreads / 32) + warp) * 8 : nullptr;
      const int tb = r == 2 ? 4 : 0;  // a group's second unit inside the same item
      if (tr) tr[tb] = clock64();
      if (active) {
        // ---- pass 1: row maximum over the valid keys; chunk c + 1 is on its way while chunk c is reduced ----
        uint32_t sb[2][32];
        float m0 = -INFINITY, m1 = -INFINITY;
        tmem_ld_32x32b_x32(t_unit, sb[0]);
#pragma unroll
        for (int cb = 0; cb < 5; ++cb) {
          const uint32_t(&s)[32] = sb[cb & 1];
          tmem_ld_wait();
          if (cb < 4) tmem_ld_32x32b_x32(t_unit + (cb + 1) * 32, sb[(cb + 1) & 1]);
          if (cb < 4) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              m0 = max3f(m0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
              m1 = max3f(m1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e < nlast) m0 = fmaxf(m0, __uint_as_float(s[e]));
          }
        }
        const float mc = fmaxf(m0, m1) * c;
        if (tr) tr[tb + 1] = clock64();
        // ---- pass 2: exponentials, row sum, bf16 probabilities over the consumed score columns ----
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        tmem_ld_32x32b_x32(t_unit, sb[0]);
#pragma unroll
        for (int cb = 0; cb < 5; ++cb) {
          const uint32_t(&s)[32] = sb[cb & 1];
          tmem_ld_wait();
          if (cb < 4) tmem_ld_32x32b_x32(t_unit + (cb + 1) * 32, sb[(cb + 1) & 1]);
          uint32_t packed[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            float pe[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              if (cb < 4 || (kTokens > 0 && e + x < kTokens - 128)) {
                pe[x] = ex2f(fmaf(__uint_as_float(s[e + x]), c, -mc));
              } else if (kTokens > 0) {
                pe[x] = 0.f;  // padded key: V's row is zero and so is the probability
              } else {
                pe[x] = e + x < nlast ? ex2f(fmaf(__uint_as_float(s[e + x]), c, -mc)) : 0.f;
              }
            }
            l0 += pe[0];
            l1 += pe[1];
            l2 += pe[2];
            l3 += pe[3];
            packed[e >> 1] = pack_bf16x2(pe[0], pe[1]);
            packed[(e >> 1) + 1] = pack_bf16x2(pe[2], pe[3]);
          }
          // keys 32 cb .. 32 cb + 31 -> columns 16 cb .. 16 cb + 15: inside the score columns already consumed
          tmem_st_32x32b_x16(t_unit + cb * 16, packed);
        }
        lsum[r * 128 + q * 32 + lane] = (l0 + l1) + (l2 + l3);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&p_ready[r]);
      if (tr) tr[tb + 2] = clock64();
      r += kSoftGroups;
      if (r >= kBufs) {
        r -= kBufs;
        ++i;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool attention_tc_supported(int T) { return T > 128 && T <= kTp; }

int attention_tc_warps() { return kThreads / 32; }

int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, float scale_log2e, int num_sms,
                        cudaStream_t stream, int reverse, long long* trace, int trace_items) {
  if (!attention_tc_supported(T)) {
    set_error("attention_tc: built for 129..160 tokens, got %d", T);
    return -1;
  }
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(
        cudaFuncSetAttribute(attention_tc_kernel<145>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    HGR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap tmQ, tmKV, tmQs;
  {
    const uint64_t row = 768 * 2;
    const uint64_t dims[5] = {768, (uint64_t)T, (uint64_t)B, 1, 1};
    const uint64_t strides[4] = {row, row * T, row * T * B, row * T * B};
    const uint32_t box_q[5] = {64, 128, 1, 1, 1};
    const uint32_t box_kv[5] = {64, (uint32_t)kTp, 1, 1, 1};
    if (int r = make_tensor_map_bf16(&tmQ, qkv, 5, dims, strides, box_q)) return r;
    if (int r = make_tensor_map_bf16(&tmKV, qkv, 5, dims, strides, box_kv)) return r;
    // the same tensor seen from token 128 on: T - 128 rows per image, everything else is out of bounds (= zero)
    const uint64_t dims_s[5] = {768, (uint64_t)(T - 128), (uint64_t)B, 1, 1};
    if (int r = make_tensor_map_bf16(&tmQs, qkv + (size_t)128 * 768, 5, dims_s, strides, box_kv)) return r;
  }
  AttnTcParams p;
  p.out = out;
  p.B = B;
  p.T = T;
  p.scale_log2e = scale_log2e;
  p.reverse = reverse;
  p.trace = trace;
  p.trace_items = trace_items;
  const int items = B * (kHeads / 2);
  const int grid = items < num_sms ? items : num_sms;
  if (grid <= 0) return 0;
  if (T == 145)
    HGR_CHECK_CUDA(launch_pdl(attention_tc_kernel<145>, dim3(grid), dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, tmQs, p));
  else
    HGR_CHECK_CUDA(launch_pdl(attention_tc_kernel<0>, dim3(grid), dim3(kThreads), kSmemBytes, stream, tmQ, tmKV, tmQs, p));
  return 0;
}

}  // namespace hgr
