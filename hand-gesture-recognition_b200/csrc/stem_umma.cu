// encoder.conv1 -> encoder.conv2 -> encoder.cspelan1.cv1 as ONE kernel, all three contractions on tcgen05: the
// 64-channel map between the first two layers (a1, 1.15 GB at batch 1024 - the largest HBM round trip of the forward)
// and the 128-channel map between the last two are never written
// (reference model/gelan.py:155 `conv1 = Conv(3, 64, 3, 2)`, :156 `conv2 = Conv(64, 128, 3, 2)`, :127
// GELANBlock.cv1 = Conv(128, 128, 1, 1); Conv.forward :56 = SiLU(BN(conv(x)))):
//
//     a1 = SiLU(BN0(conv3x3_s2(x)))         G0 (K = 27 -> 32, im2col rows built in shared memory), E0 -> parity planes
//     a2 = SiLU(BN1(conv3x3_s2(a1)))        G1 (K = 9 taps x 64 ch, A = the parity planes),        E1 -> tensor memory
//     g  = SiLU(BN2(conv1x1(a2)))           G2 (K = 128, A = a2 in tensor memory),                 E2 -> TMA store
//
// stem_chain.cu stages the (33 x 17)-pixel a1 patch of an 8 x 16 output tile as the four parity planes of the
// space-to-depth view and reads the nine taps through shifted UMMA descriptors; all nine weight taps of conv2 are
// resident here (72 KiB per CTA of the pair).  The planes are PRODUCED in place: a TMA box brings the (67 x 40)-pixel,
// 3-channel patch of the bf16 NCHW input (zero-filled outside the image = conv1's padding); the patch's 561 pixels
// are five blocks of 128 rows (planes P11, P10, P01, P00 back to back).  Two builder warps write each pixel's 27 input
// values (+ two 1.0 slots that multiply the BN shift / 2 as a bf16 hi + lo pair, so the accumulator IS h = x / 2 of
// SiLU(x) = h + h tanh(h)) as a 64-byte K-major row (SWIZZLE_64B), two `cta_group::2` MMAs (M = 256, N = 64, K = 16)
// produce the block in tensor memory, and the G0 warps (two groups of eight: two warps per tensor-memory lane
// quarter, 32 channels each; block B of the pair-wide sequence belongs to group, A buffer and accumulator B & 1) read
// their half rows, release the accumulator, apply SiLU and store the bf16 half pixels at the SWIZZLE_128B position of
// their parity plane.  The patch is single-buffered with one full / empty barrier pair PER PLANE and G1 walks its
// taps plane by plane (P11: 4 taps, P10: 2, P01: 2, P00: 1), so the planes are refilled while the tensor core works
// on the others.  Pixels of a1 outside the map (row / column -1 = conv2's padding) are stored as zeros.
//
// A first version ran conv1 on mma.sync in producer warps (HMMA.16816 occupies its scheduler's tensor sub-pipe for
// ~22 cycles on B200 and building A fragments on the fly costs ~250 issue slots per 16 pixels: 0.67 ms); the way from
// there was read off the kernel's own timeline (HGR_STEM_TRACE, tools/stem_trace.py): every role of this kernel is a
// chain of short dependent steps around shared-memory, tensor-memory and mbarrier round trips of 50-350 cycles each,
// so the roles were split until none of them paces the tile alone: G0 in channel halves (0.66 -> 0.59 ms), im2col
// building moved to the two otherwise idle control warps (-> 0.555 ms).  At that point the E1 / E2 groups (~16 k
// cycles per item each, ~8 k per tile), the builders (~1.5 k per block) and the G0 groups all sit near the tile time,
// the MUFU pipe (one tanh per a1, a2 and g element) is 60 % busy and shared memory, tensor memory (512 columns) and
// the thread count leave no room for another stage.  Measured negatives: E2 with direct global stores instead of the
// staging buffer + TMA store (0.62 ms); one E group + three G0 warps per scheduler (0.72 ms); E1 spread over the G0
// warps with a2 in its own tensor-memory buffer and G2 accumulating into the G1 stage (0.58 ms: E2's reads then sit
// inside the stage-reuse loop); __nanosleep back-off in the waits (no change).
//
// Tensor memory (512 columns): G1 stages at 0 and 128 (a2 over the first 64 columns of its stage), G2's single
// stage at 256, G0's two 64-column accumulators at 384 and 448.  CTA pairs as in stem_chain.cu; every arrival that
// goes to the pair's leader is a plain (CTA-scope release) remote arrive - what it announces (shared memory behind
// fence.proxy.async, tensor memory behind tcgen05.fence) never leaves the arriving CTA.  The leader's MMA warp is
// event driven: it polls the three in-order streams (G0 blocks, G1 planes, G2 items) and issues whatever is ready.
// 896 threads, 72 registers: warp 0 TMA, 1 MMA issue (leader), 2-3 im2col builders (warp 2 also allocates tensor
// memory), 4-11 two E1 / E2 groups (alternating items), 12-27 the two G0 groups.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "epilogue_math.cuh"
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 896;                  // see the header for the roles; 65536 / 896 -> 72 registers per thread
constexpr int kG0Warps = 8;                    // warps per G0 group
constexpr int kC = 128;                        // channels of a2 and of g
constexpr int kC0 = 64;                        // channels of a1
constexpr int kTW = 8, kTH = 16;               // output tile
// parity planes of the a1 patch: rows x columns of 128-byte pixels; G0's row order is P11, P10, P01, P00
constexpr int kOffP00 = 0;                                   // 16 x 8
constexpr int kOffP01 = kOffP00 + 16 * 8 * 128;              // 16 x 9
constexpr int kOffP10 = kOffP01 + 16 * 9 * 128;              // 17 x 8
constexpr int kOffP11 = kOffP10 + 17 * 8 * 128;              // 17 x 9 (padded to 20 KiB)
constexpr int kPatchBytes = kOffP11 + 20 * 1024;
constexpr int kRows = 153 + 136 + 144 + 128;   // 561 pixels of the patch
constexpr int kBlocks = 5;                     // blocks of 128 rows
constexpr int kTapBytes = 64 * 128;            // this CTA's 64 weight rows of one tap / of one k-block of cv1
constexpr int kOutBytes = 128 * 64;            // one 32-channel chunk of an output tile (SWIZZLE_64B rows)
// input patch: x rows 4 h0 - 3 .., columns 4 w0 - 8 .. (TMA wants the innermost start on a 16-byte boundary; the
// first column a tap reads is 4 w0 - 3 = patch column 5, the last 4 w0 + 31 = patch column 39)
constexpr int kXCols = 40, kXRows = 67;
constexpr int kXLoadBytes = 3 * kXRows * kXCols * 2;
constexpr int kXBytes = 16128;
constexpr int kABytes = 128 * 64;              // one im2col block: 128 rows x 32 k, SWIZZLE_64B
constexpr int kW0Bytes = 32 * 64;              // this CTA's 32 rows of conv1's weights, SWIZZLE_64B
constexpr int kOffPatch = 0;
constexpr int kOffW1 = kOffPatch + kPatchBytes;
constexpr int kOffW2 = kOffW1 + 9 * kTapBytes;
constexpr int kOffOut = kOffW2 + 2 * kTapBytes;
constexpr int kOffA = kOffOut + 2 * kOutBytes;
constexpr int kOffW0 = kOffA + 2 * kABytes;
constexpr int kOffX = kOffW0 + kW0Bytes;
constexpr int kOffAffine = kOffX + 2 * kXBytes;  // scale1, shift1, scale2, shift2: 4 x 128 floats, pre-halved
constexpr int kOffBars = kOffAffine + 4 * kC * 4;
constexpr int kNumBars = 4 + 4 + 2 + 2 + 2 + 2 + 2 + 2 + 3 + 2;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kPatchBytes % 1024 == 0 && kOffP01 % 1024 == 0 && kOffP10 % 1024 == 0 && kOffP11 % 1024 == 0,
              "planes start on swizzle-atom boundaries");
static_assert(kOffW1 % 1024 == 0 && kOffW2 % 1024 == 0 && kOffOut % 1024 == 0 && kOffA % 1024 == 0 &&
                  kOffW0 % 1024 == 0 && kOffX % 128 == 0,
              "operand alignment");
static_assert(kXLoadBytes <= kXBytes && kXBytes % 128 == 0, "input patch buffer");
static_assert(kSmemBytes <= 227 * 1024, "stem_umma shared-memory plan exceeds one CTA");
static_assert(kRows <= kBlocks * 128, "five blocks hold the patch");

__device__ __forceinline__ int plane_offset(int pr, int pc) {
  return pr == 0 ? (pc == 0 ? kOffP00 : kOffP01) : (pc == 0 ? kOffP10 : kOffP11);
}

// K-major operand tile with 64-byte rows (K = 32 bf16), SWIZZLE_64B: 8-row groups 512 bytes apart
// (cute/arch/mma_sm100_desc.hpp: layout type 4)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((512 >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

// No __noinline__ call in this kernel (see stem_fused.cu): a timed-out wait leaves its marks in the mapped debug
// buffer, if one is set, and traps.
__device__ __forceinline__ void wait_timeout(uint32_t addr, uint32_t parity) {
  if (hgr_dbg_ptr) {
    hgr_dbg_ptr[1] = blockIdx.x;
    hgr_dbg_ptr[2] = threadIdx.x;
    hgr_dbg_ptr[3] = addr;
    hgr_dbg_ptr[4] = parity;
    hgr_dbg_ptr[0] = 0xDEADu;
    __threadfence_system();
  }
  __trap();
}

__device__ __forceinline__ void wait_cta(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > (1ll << 31)) wait_timeout(smem_u32(bar), parity);
  }
}

// Non-blocking test for the MMA warp's polling loop (test_wait: try_wait may suspend the thread for a system-dependent
// time).  Default semantics (acquire at CTA scope), as CUTLASS's ClusterBarrier waits: an acquire at cluster scope
// invalidates L1 on every poll (measured: ~1400 cycles from the last arrival to the retired MMA, 30 % of the G0 warps'
// time), and nothing the leader's thread reads itself is written by the peer - the arrivals only announce that the
// peer's shared / tensor memory is ready for the peer's own tensor core.
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// arrive on the LEADER's copy of a barrier
__device__ __forceinline__ void arrive_leader(uint64_t* bar, uint32_t cta_rank) {
  if (cta_rank == 0) mbar_arrive(bar);
  else mbar_arrive_remote(bar, 0);
}

__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

constexpr int kTraceSlots = 512;

struct StemUmmaParams {
  int tiles_w, tiles_h, total_tiles;  // per image: tiles_w x tiles_h tiles of 8 x 16; total over the batch
  int out_c_off;
  int reverse;
  int debug;                // HGR_STEM_DEBUG=1: a timed-out wait leaves its marks in mapped host memory
  long long* trace;         // HGR_STEM_TRACE=<file>: [8 roles][kTraceSlots][2] = {event << 32 | index, clock64} of CTA 0
  const __nv_bfloat16* w0;  // conv1: [64][32] bf16, k = (kh * 3 + kw) * 3 + c, BN scale folded in, k >= 27 zero
  const float* shift0;
  const float* scale1;
  const float* shift1;
  const float* scale2;
  const float* shift2;
};

// 32 values of one pixel row -> SiLU(affine) -> 16 packed bf16 pairs
__device__ __forceinline__ void activate_pack32(const uint32_t (&acc)[32], const float* s_scale, const float* s_shift,
                                                uint32_t (&packed)[16]) {
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + e);
    const float4 sh = *reinterpret_cast<const float4*>(s_shift + e);
    const float v0 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e]), sc.x, sh.x));
    const float v1 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e + 1]), sc.y, sh.y));
    const float v2 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e + 2]), sc.z, sh.z));
    const float v3 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e + 3]), sc.w, sh.w));
    packed[e >> 1] = pack_bf16x2(v0, v1);
    packed[(e >> 1) + 1] = pack_bf16x2(v2, v3);
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
stem_umma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
                 const StemUmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* plane_full = bars;               // [4] leader's: every G0 warp of both CTAs has stored its share
  uint64_t* plane_empty = bars + 4;          // [4] the taps that read the plane have retired (both CTAs)
  uint64_t* x_full = bars + 8;               // [2] this CTA's input patch has landed
  uint64_t* x_empty = bars + 10;             // [2] every G0 warp of this CTA has built its last block from the buffer
  uint64_t* a_full = bars + 12;              // [2] leader's: both CTAs' im2col blocks are in shared memory
  uint64_t* c1_full = bars + 14;             // [2] G0 of a block has retired (both CTAs)
  uint64_t* acc_full = bars + 16;            // [2] G1 of an item has retired (both CTAs)
  uint64_t* a2_ready = bars + 18;            // [2] leader's: both CTAs' a2 tiles are in tensor memory
  uint64_t* acc2_full = bars + 20;           // G2 of an item has retired (both CTAs)
  uint64_t* acc2_empty = bars + 21;          // leader's: both CTAs' E2 have read G2's accumulator
  uint64_t* w_bar = bars + 22;               // leader's: both halves of conv2's and cv1's weights are resident
  uint64_t* c1_empty = bars + 23;            // [2] leader's: every G0 warp of both CTAs has read the accumulator
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  float* s_aff = reinterpret_cast<float*>(smem + kOffAffine);

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // dynamic smem base not 1024-byte aligned
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmW2);
    prefetch_tensormap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&plane_full[i], 2 * 2 * kG0Warps);  // one arrival per G0 warp, both CTAs
      mbar_init(&plane_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 2);       // the two builder warps
      mbar_init(&a_full[i], 2 * 2);    // the two builder warps, both CTAs
      mbar_init(&c1_empty[i], 2 * kG0Warps);  // one arrival per warp of the group, both CTAs
      mbar_init(&c1_full[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&a2_ready[i], 2 * 4);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 2 * 4);
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  const uint32_t cta_rank = cluster_ctarank();
  // SiLU is evaluated on h = x / 2, so the 1/2 is folded into both affines (as in gemm_tcgen05.cu's load_affine)
  for (int i = threadIdx.x; i < kC; i += kThreads) {
    s_aff[i] = 0.5f * (p.scale1 ? p.scale1[i] : 1.0f);
    s_aff[kC + i] = 0.5f * (p.shift1 ? p.shift1[i] : 0.0f);
    s_aff[2 * kC + i] = 0.5f * (p.scale2 ? p.scale2[i] : 1.0f);
    s_aff[3 * kC + i] = 0.5f * (p.shift2 ? p.shift2[i] : 0.0f);
  }
  // conv1's weights -> this CTA's half of G0's B operand (rows n = 32 rank .. + 31, K-major, SWIZZLE_64B): halved
  // (exact in bf16), k = 27 / 28 carry shift / 2 as a bf16 hi + lo pair against A slots that hold 1.0
  for (int e = threadIdx.x; e < 32 * 4; e += kThreads) {
    const int nl = e >> 2, ch = e & 3;  // local row, 16-byte chunk = k 8 ch .. 8 ch + 7
    const int n = (int)cta_rank * 32 + nl;
    const float hs = 0.5f * __ldg(p.shift0 + n);
    const __nv_bfloat16 hi = __float2bfloat16_rn(hs);
    const __nv_bfloat16 lo = __float2bfloat16_rn(hs - __bfloat162float(hi));
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned short w[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = ch * 8 + 2 * j + h;
        __nv_bfloat16 val = __float2bfloat16_rn(0.0f);
        if (k < 27) val = __float2bfloat16_rn(0.5f * __bfloat162float(p.w0[n * 32 + k]));
        else if (k == 27) val = hi;
        else if (k == 28) val = lo;
        w[h] = __bfloat16_as_ushort(val);
      }
      v[j] = (uint32_t)w[0] | ((uint32_t)w[1] << 16);
    }
    *reinterpret_cast<uint4*>(smem + kOffW0 + nl * 64 + ((ch ^ ((nl >> 1) & 3)) << 4)) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  fence_proxy_async_smem();  // W0 is read by the tensor core
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  const int total_items = (p.total_tiles + 1) / 2;  // work items of the pair-wide walk
  const int first = blockIdx.x / 2, stride = gridDim.x / 2;
  const int n_items = first < total_items ? (total_items - first + stride - 1) / stride : 0;
  // item -> this CTA's tile origin (it may lie beyond the batch: TMA clips loads and stores)
  auto coords = [&](int item, int& w0, int& h0, int& n0) {
    if (p.reverse) item = total_items - 1 - item;
    int mt = item * 2 + (int)cta_rank;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    n0 = mt / p.tiles_h;
    w0 = tw * kTW;
    h0 = th * kTH;
  };

  // development timeline (HGR_STEM_TRACE, its own instantiation: the stamps cost the G0 loop ~10 %): lane 0 of one
  // warp per role of CTA 0 stamps its events
  int trace_n = 0;
  auto mark = [&](int role, int event, int index) {
    if constexpr (!TRACE) return;
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && trace_n < kTraceSlots) {
      long long* t = p.trace + ((size_t)role * kTraceSlots + trace_n) * 2;
      t[0] = ((long long)event << 32) | (unsigned)index;
      t[1] = clock64();
      ++trace_n;
    }
  };

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own half of the weights once, then the input patches =================
    if (elect_one_sync()) {
      if (cta_rank == 0) mbar_expect_tx(w_bar, 2 * 11 * kTapBytes);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d_2sm(smem + kOffW1 + tap * kTapBytes, &tmW, w_bar, tap * 64, (int)cta_rank * (kC / 2));
      for (int kb = 0; kb < 2; ++kb)
        tma_load_2d_2sm(smem + kOffW2 + kb * kTapBytes, &tmW2, w_bar, kb * 64, (int)cta_rank * (kC / 2));
      int iter = 0;
      for (int item = first; item < total_items; item += stride, ++iter) {
        int w0, h0, n0;
        coords(item, w0, h0, n0);
        const int xb = iter & 1;
        wait_cta(&x_empty[xb], ((iter >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[xb], kXLoadBytes);
        tma_load_4d(smem + kOffX + xb * kXBytes, &tmX, &x_full[xb], 4 * w0 - 8, 4 * h0 - 3, 0, n0);
        mark(0, 0, iter);
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0) {
      // ================= MMA issuer (leader), event driven over three in-order streams =================
      constexpr uint32_t idesc0 = umma_idesc_bf16(256, kC0);
      constexpr uint32_t idesc = umma_idesc_bf16(256, kC);
      const uint32_t patch = smem_u32(smem + kOffPatch);
      wait_cta(w_bar, 0);
      const int total_blocks = kBlocks * n_items;
      int cb = 0;          // G0: blocks issued
      int gi = 0, gp = 0;  // G1: item and plane to issue next
      int g2 = 0;          // G2: items issued
      long long t0 = clock64();
      while (cb < total_blocks || gi < n_items || g2 < n_items) {
        bool progress = false;
        // ---- G0: block cb from A buffer cb & 1 into accumulator cb & 1 ----
        if (cb < total_blocks && __all_sync(0xffffffffu, test_wait(&a_full[cb & 1], (cb >> 1) & 1) &&
                                                            test_wait(&c1_empty[cb & 1], ((cb >> 1) & 1) ^ 1))) {
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t tmem_d = tmem_base + 3 * kC + (cb & 1) * kC0;
            const uint64_t a_base = umma_desc_sw64(smem_u32(smem + kOffA + (cb & 1) * kABytes));
            const uint64_t b_base = umma_desc_sw64(smem_u32(smem + kOffW0));
            umma_bf16_ss_2sm(tmem_d, a_base, b_base, idesc0, 0u);
            umma_bf16_ss_2sm(tmem_d, a_base + 2, b_base + 2, idesc0, 1u);
            umma_commit_2sm(&c1_full[cb & 1], 0b11);
          }
          __syncwarp();
          mark(1, 0, cb);
          ++cb;
          progress = true;
        }
        // ---- G1: plane gp of item gi; its stage gi & 1 held a2 of item gi - 2, read by G2(gi - 2) ----
        if (gi < n_items && (gp != 0 || g2 >= gi - 1) && __all_sync(0xffffffffu, test_wait(&plane_full[gp], gi & 1))) {
          tc_fence_after();
          const int pr = gp < 2 ? 1 : 0, pc = (gp & 1) ? 0 : 1;
          const int pw = 8 + pc;  // plane width in pixels
          const uint32_t tmem_d = tmem_base + (gi & 1) * kC;
          if (elect_one_sync()) {
            for (int a = 0; a < (pr ? 2 : 1); ++a)
              for (int b = 0; b < (pc ? 2 : 1); ++b) {
                const int kh = pr ? 2 * a : 1, kw = pc ? 2 * b : 1;
                const uint32_t a_addr = patch + plane_offset(pr, pc) + ((kh == 2 ? pw : 0) + (kw == 2 ? 1 : 0)) * 128;
                const uint64_t a_base = umma_desc_sw128(a_addr, pw * 128);
                const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffW1 + (kh * 3 + kw) * kTapBytes), 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ss_2sm(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (gp | a | b | k) != 0 ? 1u : 0u);
              }
            umma_commit_2sm(&plane_empty[gp], 0b11);
            if (gp == 3) umma_commit_2sm(&acc_full[gi & 1], 0b11);
          }
          __syncwarp();
          mark(1, 1, gi * 4 + gp);
          if (++gp == 4) {
            gp = 0;
            ++gi;
          }
          progress = true;
        }
        // ---- G2: item g2 (its G1 has been issued), once E1 has written a2 and E2 of the item before has left ----
        if (g2 < gi && __all_sync(0xffffffffu, test_wait(&a2_ready[g2 & 1], (g2 >> 1) & 1) &&
                                              test_wait(acc2_empty, (g2 & 1) ^ 1))) {
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t tmem_a = tmem_base + (g2 & 1) * kC;  // a2 over the first 64 columns of the G1 stage
            const uint32_t tmem_d = tmem_base + 2 * kC;
#pragma unroll
            for (int k = 0; k < kC / 16; ++k)
              umma_bf16_ts_2sm(tmem_d, tmem_a + 8 * k,
                               umma_desc_sw128(smem_u32(smem + kOffW2 + (k >> 2) * kTapBytes), 1024) + 2 * (k & 3), idesc,
                               k != 0 ? 1u : 0u);
            umma_commit_2sm(acc2_full, 0b11);
          }
          __syncwarp();
          mark(1, 2, g2);
          ++g2;
          progress = true;
        }
        if (progress) {
          t0 = clock64();
        } else {
          if (clock64() - t0 > (1ll << 31)) wait_timeout((uint32_t)(cb | (gi << 12) | (gp << 20) | (g2 << 22)), 0x300u);
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ================= im2col builders: two rows of every block per thread, blocks in pair-wide order =================
    const int tb = (warp - 2) * 32 + lane;
    uint32_t tab_x[kBlocks][2];
#pragma unroll
    for (int b = 0; b < kBlocks; ++b)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int L = 128 * b + tb + 64 * j;
        const int pl = L < 153 ? 0 : (L < 289 ? 1 : (L < 433 ? 2 : 3));
        const int pr = pl < 2 ? 1 : 0, pc = (pl & 1) ? 0 : 1;
        const int pw = 8 + pc;
        const int idx = L < kRows ? L - (pl == 0 ? 0 : (pl == 1 ? 153 : (pl == 2 ? 289 : 433))) : 0;
        const int y = pc ? (idx * 57) >> 9 : idx >> 3;
        const int x = idx - y * pw;
        tab_x[b][j] = 2 * ((4 * y - 2 * pr + 2) * kXCols + 4 * x - 2 * pc + 7);
      }
    const uint32_t smem_base = smem_u32(smem);
    for (int it = 0; it < n_items; ++it) {
      const int xb = it & 1;
      wait_cta(&x_full[xb], (it >> 1) & 1);
      const uint32_t xs = smem_base + kOffX + xb * kXBytes;
#pragma unroll
      for (int b = 0; b < kBlocks; ++b) {
        const int B = kBlocks * it + b;
        const int buf = B & 1, use = B >> 1;
        wait_cta(&c1_full[buf], (use & 1) ^ 1);  // G0 of the block before in this A buffer has retired
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int r = tb + 64 * j;
          const uint32_t src = xs + tab_x[b][j];
          uint32_t v[16];
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            uint32_t lo = 0, hi = 0;
            const int k = 2 * k2;
            if (k < 27) lo = lds16(src + 2 * (((k % 3) * kXRows + k / 9) * kXCols + (k / 3) % 3));
            if (k + 1 < 27) hi = lds16(src + 2 * ((((k + 1) % 3) * kXRows + (k + 1) / 9) * kXCols + ((k + 1) / 3) % 3));
            v[k2] = lo | (hi << 16);
          }
          v[13] |= 0x3F800000u;  // k = 27: 1.0 against hi(shift / 2)
          v[14] = 0x00003F80u;   // k = 28: 1.0 against lo(shift / 2); k = 29: 0
          v[15] = 0u;
          const uint32_t dst = smem_base + kOffA + buf * kABytes + r * 64;
          const uint32_t sw = static_cast<uint32_t>((r >> 1) & 3);
#pragma unroll
          for (int c = 0; c < 4; ++c) sts128(dst + (((uint32_t)c ^ sw) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        fence_proxy_async_smem();  // read by the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) arrive_leader(&a_full[buf], cta_rank);
        if (tb == 0) mark(6, 0, B);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[xb]);
    }
  } else if (warp >= 4 && warp < 12) {
    // ================= E1 / E2 groups: group g owns the G1 stage g; G2's stage and the staging chunks alternate =================
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - 128 - group * 128;
    const uint32_t bar_id = 1 + group;
    const uint32_t sw = static_cast<uint32_t>((row >> 1) & 3);  // SWIZZLE_64B: 16-byte chunk ^= bits 7-8 of the address
    uint8_t* stage_out = smem + kOffOut + group * kOutBytes;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int iter = 0;
    for (int item = first; item < total_items; item += stride, ++iter) {
      if ((iter & 1) != group) continue;
      const uint32_t ph = (iter >> 1) & 1;
      int w0, h0, n0;
      coords(item, w0, h0, n0);

      // ---------------- E1: a2 tile -> tensor memory (A operand of G2), 32 channels at a time ----------------
      if (q == 0) mark(2 + group, 0, iter);
      wait_cta(&acc_full[group], ph);
      tc_fence_after();
      if (q == 0) mark(2 + group, 1, iter);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t acc[32], packed[16];
        tmem_ld_32x32b_x32(t_row + group * kC + j * 32, acc);
        tmem_ld_wait();
        activate_pack32(acc, s_aff + j * 32, s_aff + kC + j * 32, packed);
        // channels 32 j .. 32 j + 31 -> columns 16 j .. 16 j + 15 of the stage: inside what this thread has consumed
        tmem_st_32x32b_x16(t_row + group * kC + j * 16, packed);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(&a2_ready[group], cta_rank);
      if (q == 0) mark(2 + group, 2, iter);

      // ---------------- E2: g tile, one 32-channel chunk at a time -> staging -> TMA store ----------------
      wait_cta(acc2_full, iter & 1);
      tc_fence_after();
      if (q == 0) mark(2 + group, 3, iter);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t acc[32], packed[16];
        tmem_ld_32x32b_x32(t_row + 2 * kC + j * 32, acc);
        tmem_ld_wait();
        if (j == 3) {  // G2's accumulator has been read: the next item's G2 may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_leader(acc2_empty, cta_rank);
        }
        activate_pack32(acc, s_aff + 2 * kC + j * 32, s_aff + 3 * kC + j * 32, packed);
        if (gtid == 0) tma_store_wait_read<0>();  // the previous chunk has left the staging buffer
        bar_sync(bar_id, 128);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          *reinterpret_cast<uint4*>(stage_out + row * 64 + ((static_cast<uint32_t>(v) ^ sw) << 4)) =
              make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
        fence_proxy_async_smem();
        bar_sync(bar_id, 128);
        if (gtid == 0) {
          tma_store_4d(&tmO, stage_out, p.out_c_off + j * 32, w0, h0, n0);
          tma_store_commit();
        }
      }
      if (q == 0) mark(2 + group, 4, iter);
    }
    if (gtid == 0) tma_store_wait_all();
  } else if (warp >= 12) {
    // ================= G0 groups: im2col rows in, SiLU'd a1 pixels out to the parity planes =================
    const int group = (warp - 12) >> 3;
    const int hf = ((warp - 12) >> 2) & 1;  // this warp's half of the pixel: channels 32 hf .. + 31
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of every block this thread post-processes = its tensor-memory lane
    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 3 * kC + group * kC0 + hf * 32;
    // Per block b of an item, fixed for the whole kernel: row L = 128 b + r of the patch in plane order.  tab_p = byte
    // offset of the pixel in the a1 patch | swizzle phase << 20 | flags << 24 (1 = row exists, 2 = a1 row -1 when
    // h0 == 0, 4 = a1 column -1 when w0 == 0).
    uint32_t tab_p[kBlocks];
#pragma unroll
    for (int b = 0; b < kBlocks; ++b) {
      const int L = 128 * b + r;
      const int pl = L < 153 ? 0 : (L < 289 ? 1 : (L < 433 ? 2 : 3));
      const int pr = pl < 2 ? 1 : 0, pc = (pl & 1) ? 0 : 1;
      const int pw = 8 + pc;
      const bool exists = L < kRows;
      const int idx = exists ? L - (pl == 0 ? 0 : (pl == 1 ? 153 : (pl == 2 ? 289 : 433))) : 0;
      const int y = pc ? (idx * 57) >> 9 : idx >> 3;
      const int x = idx - y * pw;
      tab_p[b] = (uint32_t)(kOffPatch + plane_offset(pr, pc) + idx * 128) | ((uint32_t)(idx & 7) << 20) |
                 ((exists ? 1u : 0u) | ((pr && y == 0) ? 2u : 0u) | ((pc && x == 0) ? 4u : 0u)) << 24;
    }
    auto pick = [](const uint32_t (&t)[kBlocks], int b) {
      return b == 0 ? t[0] : (b == 1 ? t[1] : (b == 2 ? t[2] : (b == 3 ? t[3] : t[4])));
    };
    const uint32_t smem_base = smem_u32(smem);
    // The group's blocks are B = group, group + 2, ... of the pair-wide sequence (item B / 5, block B % 5), through
    // its own accumulator: as soon as G0(B) has retired the half row is pulled into registers and the accumulator is
    // released, so that the round trip of G0(B + 2) (the builders run ahead) lies under this block's arithmetic.
    auto advance = [](int& it, int& b, bool& first_of_item) {
      b += 2;
      first_of_item = b >= kBlocks;
      if (first_of_item) {
        b -= kBlocks;
        ++it;
      }
    };
    int cit = 0, cb = group;
    bool cfirst = true;
    uint32_t edge = 0;
    int use = 0;  // uses of this group's A buffer / accumulator so far (barrier phase)
    while (cit < n_items) {
      // ---------------- this warp's half of the accumulator row of the block -> registers ----------------
      const bool tr = q == 0 && hf == 0;
      const int B = kBlocks * cit + cb;
      if (tr) mark(4 + group, 0, B);
      wait_cta(&c1_full[group], use & 1);
      ++use;
      tc_fence_after();
      if (tr) mark(4 + group, 1, B);
      uint32_t acc[32];
      tmem_ld_32x32b_x32(t_acc, acc);
      tmem_ld_wait();
      tc_fence_before();  // the reads are done: the group's next G0 may overwrite the accumulator
      __syncwarp();
      if (lane == 0) arrive_leader(&c1_empty[group], cta_rank);
      if (tr) mark(4 + group, 2, B);
      // ---------------- the block: SiLU -> bf16 half pixel in its parity plane ----------------
      if (cfirst) {
        int w0, h0, n0;
        coords(first + cit * stride, w0, h0, n0);
        edge = (h0 == 0 ? 2u : 0u) | (w0 == 0 ? 4u : 0u);
      }
      const uint32_t eph = (cit & 1) ^ 1;
      // block b overlaps planes b - 1 and b: the previous item's taps on them must have retired
      if (cb >= 1) wait_cta(&plane_empty[cb - 1], eph);
      if (cb <= 3) wait_cta(&plane_empty[cb], eph);
      if (tr) mark(4 + group, 4, B);
      const uint32_t tps = pick(tab_p, cb);
      const uint32_t flags = tps >> 24;
      const bool exists = flags & 1u;
      const bool zero = (flags & edge) != 0;  // a1 pixel outside the map = conv2's zero padding
      const uint32_t prow = smem_base + (tps & 0xFFFFFu);
      const uint32_t psw = (tps >> 20) & 7u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d0 = __uint_as_float(acc[8 * c + 2 * e]), d1 = __uint_as_float(acc[8 * c + 2 * e + 1]);
          pk[e] = zero ? 0u : pack_bf16x2(fmaf(d0, tanh_approx(d0), d0), fmaf(d1, tanh_approx(d1), d1));  // d = x / 2
        }
        if (exists) sts128(prow + (((uint32_t)(4 * hf + c) ^ psw) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();  // the pixels are read by the tensor core
      __syncwarp();
      if (lane == 0) {
        if (cb >= 1) arrive_leader(&plane_full[cb - 1], cta_rank);
        if (cb <= 3) arrive_leader(&plane_full[cb], cta_rank);
      }
      if (tr) mark(4 + group, 5, B);
      advance(cit, cb, cfirst);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while its peer may still read its shared memory or arrive on its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace

bool stem_umma_supported(int S) { return S % 64 == 0 && S >= 64 && stem_chain_supported(S / 2, S / 2); }

// x: (B, 3, S, S) bf16 NCHW; w0 [64][32] bf16 + shift0 (conv1, packed as for conv1.cu); w1 [128][3][3][64],
// w2 [128][128] bf16; out: channel slice [out_coff, +128) of a (B, S / 4, S / 4, out_ctot) buffer.
int run_stem_umma(const void* x, int B, int S, const void* w0, const float* shift0, const void* w1, const float* scale1,
                  const float* shift1, const void* w2, const float* scale2, const float* shift2, void* out,
                  int out_ctot, int out_coff, int reverse, int num_sms, cudaStream_t stream) {
  if (!stem_umma_supported(S)) {
    set_error("stem_umma: image side %d must be a multiple of 64", S);
    return -1;
  }
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) {
    set_error("stem_umma: the input batch must be 16-byte aligned");
    return -1;
  }
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(stem_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    HGR_CHECK_CUDA(cudaFuncSetAttribute(stem_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap tx, tw, tw2, to;
  const int Ho = S / 4, Wo = S / 4;
  {
    const uint64_t dims[4] = {(uint64_t)S, (uint64_t)S, 3, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)S * 2, (uint64_t)S * S * 2, (uint64_t)3 * S * S * 2};
    const uint32_t box[4] = {(uint32_t)kXCols, (uint32_t)kXRows, 3, 1};
    if (int r = make_tensor_map_bf16(&tx, x, 4, dims, strides, box, 0)) return r;
  }
  {
    const uint64_t dims[2] = {576, (uint64_t)kC};
    const uint64_t strides[1] = {576 * 2};
    const uint32_t box[2] = {64, (uint32_t)(kC / 2)};
    if (int r = make_tensor_map_bf16(&tw, w1, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)kC, (uint64_t)kC};
    const uint64_t strides[1] = {(uint64_t)kC * 2};
    const uint32_t box[2] = {64, (uint32_t)(kC / 2)};
    if (int r = make_tensor_map_bf16(&tw2, w2, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)out_ctot, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t row = (uint64_t)out_ctot * 2;
    const uint64_t strides[3] = {row, row * Wo, row * Wo * Ho};
    const uint32_t box[4] = {32, (uint32_t)kTW, (uint32_t)kTH, 1};
    if (int r = make_tensor_map_bf16(&to, out, 4, dims, strides, box, 64)) return r;
  }
  StemUmmaParams p;
  p.tiles_w = Wo / kTW;
  p.tiles_h = Ho / kTH;
  p.total_tiles = p.tiles_w * p.tiles_h * B;
  p.out_c_off = out_coff;
  p.reverse = reverse;
  p.debug = getenv("HGR_STEM_DEBUG") ? atoi(getenv("HGR_STEM_DEBUG")) : 0;
  p.trace = nullptr;
  const char* trace_path = getenv("HGR_STEM_TRACE");
  const size_t trace_bytes = (size_t)8 * kTraceSlots * 2 * sizeof(long long);
  if (trace_path) {
    HGR_CHECK_CUDA(cudaMalloc(&p.trace, trace_bytes));
    HGR_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, trace_bytes, stream));
  }
  p.w0 = static_cast<const __nv_bfloat16*>(w0);
  p.shift0 = shift0;
  p.scale1 = scale1;
  p.shift1 = shift1;
  p.scale2 = scale2;
  p.shift2 = shift2;
  const int items = (p.total_tiles + 1) / 2;
  int grid = items * 2 < num_sms ? items * 2 : num_sms;
  grid -= grid % 2;
  if (grid <= 0) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  static unsigned int* dbg_host = nullptr;
  if (p.debug && !dbg_host) {
    unsigned int* dptr = nullptr;
    HGR_CHECK_CUDA(cudaHostAlloc(&dbg_host, 256, cudaHostAllocMapped));
    memset(dbg_host, 0, 256);
    HGR_CHECK_CUDA(cudaHostGetDevicePointer(&dptr, dbg_host, 0));
    HGR_CHECK_CUDA(cudaMemcpyToSymbol(hgr_dbg_ptr, &dptr, sizeof(dptr)));
  }
  if (trace_path) HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, stem_umma_kernel<true>, tx, tw, tw2, to, p));
  else HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, stem_umma_kernel<false>, tx, tw, tw2, to, p));
  if (trace_path) {
    std::vector<long long> h(trace_bytes / sizeof(long long));
    HGR_CHECK_CUDA(cudaStreamSynchronize(stream));
    HGR_CHECK_CUDA(cudaMemcpy(h.data(), p.trace, trace_bytes, cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int role = 0; role < 8; ++role)
        for (int i = 0; i < kTraceSlots; ++i) {
          const long long a = h[((size_t)role * kTraceSlots + i) * 2], t = h[((size_t)role * kTraceSlots + i) * 2 + 1];
          if (t != 0) fprintf(f, "%d %d %d %lld\n", role, (int)(a >> 32), (int)(a & 0xffffffff), t);
        }
      fclose(f);
    }
  }
  if (p.debug & 1) {
    cudaError_t e = cudaStreamSynchronize(stream);
    fprintf(stderr, "hgr stem_umma debug %d: sync -> %s; marks %x block %u thread %u %x %x\n", p.debug, cudaGetErrorName(e),
            dbg_host[0], dbg_host[1], dbg_host[2], dbg_host[3], dbg_host[4]);
  }
  return 0;
}

}  // namespace hgr
