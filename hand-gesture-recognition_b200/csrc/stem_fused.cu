// encoder.conv1 -> encoder.conv2 -> encoder.cspelan1.cv1 as ONE kernel: the 64-channel map between the first two
// layers (a1, 1.15 GB at batch 1024 - the largest HBM round trip of the forward) is never written
// (reference model/gelan.py:155 `conv1 = Conv(3, 64, 3, 2)`, :156 `conv2 = Conv(64, 128, 3, 2)`, :127
// GELANBlock.cv1 = Conv(128, 128, 1, 1); Conv.forward :56 = SiLU(BN(conv(x)))):
//
//     a1 = SiLU(BN0(conv3x3_s2(x)))         producer warps, mma.sync, written into shared memory
//     a2 = SiLU(BN1(conv3x3_s2(a1)))        G1 (K = 9 taps x 64 ch), E1
//     g  = SiLU(BN2(conv1x1(a2)))           G2 (K = 128),            E2
//
// stem_chain.cu stages the (33 x 17)-pixel a1 patch of an 8 x 16 output tile as the four parity planes of the
// space-to-depth view and reads the nine taps through shifted UMMA descriptors.  Here the same planes are PRODUCED in
// place: a TMA box brings the (67 x 40)-pixel, 3-channel patch of the bf16 NCHW input (zero-filled outside the
// image = conv1's padding), eight producer warps run conv1 on it with mma.sync m16n8k16 on m-tiles of 16 consecutive
// plane pixels and store the bf16 results at the SWIZZLE_128B position the tensor core expects; pixels of a1 outside
// the map (row / column -1 = conv2's padding) are stored as zeros.  The patch is single-buffered with one full /
// empty barrier pair PER PLANE, and G1 walks its taps plane by plane in the order the planes are produced (P11: 4
// taps, P10: 2, P01: 2, P00: 1), so the producers refill a plane while the tensor core works on the others.  What
// the second patch buffer of stem_chain.cu occupied now holds all nine weight taps of conv2 (72 KiB per CTA of the
// pair, loaded once), so a tile's only operand traffic is its 16 KiB input patch.
//
// conv1 in the producers.  A scheduler has to issue every instruction of its producer warps, so the contraction is
// laid out for the fewest instructions rather than the fewest MMAs: K is 3 channels x 4 input rows x 4 input columns
// = 48 (k = 16 c + 4 kh + slot, slot = kw + 1; row kh = 3 and column slot 0 carry zero weights), which makes every
// A-fragment register ONE aligned 32-bit shared-memory load at a compile-time offset from the pixel's base address
// (12 loads per m-tile instead of 32 16-bit loads + 16 permutes + their address arithmetic).  The B fragments (48 x
// 64 weights, halved, BN scale folded in) sit in shared memory in fragment order; the BN shift / 2 rides in the unused
// (c = 0, kh = 3) slots 0, 1 as a bf16 hi + lo pair against A slots forced to 1.0, so the accumulator IS h = x / 2 of
// SiLU(x) = h + h tanh(h): one MUFU, one FMA and half a pack per output.
//
// CTA pairs (cta_group::2, M = 256) and tensor-memory plan as in stem_chain.cu: two G1 accumulator stages (columns
// 0-255; the bf16 a2 tile over the first 64 columns of its stage, TMEM A operand of G2) and two G2 stages (columns
// 256-511).  640 threads: warps 0-3 control (TMA, MMA issue, TMEM allocation), 4-11 two epilogue groups, 12-19
// producers (two per scheduler: one sits in its MUFU phase while the other issues); setmaxnreg moves registers from
// the control warps to the producers.  No __noinline__ call anywhere in the kernel: one ABI call makes ptxas hold
// EVERY role to the smallest setmaxnreg value.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "epilogue_math.cuh"
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kProdWarps = 8;
constexpr int kThreads = 128 + 256 + kProdWarps * 32;
constexpr int kRegsLaunch = 96, kRegsCtrl = 56, kRegsProd = 112;  // 65536 / 640 threads -> 96 at launch
constexpr int kC = 128;                        // channels of a2 and of g
constexpr int kTW = 8, kTH = 16;               // output tile
// parity planes of the a1 patch in production order p = 0..3: (row parity, column parity) = (1,1) (1,0) (0,1) (0,0)
constexpr int kOffP00 = 0;                                   // 16 x 8 pixels of 128 bytes
constexpr int kOffP01 = kOffP00 + 16 * 8 * 128;              // 16 x 9
constexpr int kOffP10 = kOffP01 + 16 * 9 * 128;              // 17 x 8
constexpr int kOffP11 = kOffP10 + 17 * 8 * 128;              // 17 x 9 (padded to 20 KiB)
constexpr int kPatchBytes = kOffP11 + 20 * 1024;
// 36 m-tiles of 16 plane pixels: 10 + 9 + 9 + 8
constexpr int kTapBytes = 64 * 128;            // this CTA's 64 weight rows of one tap / of one k-block of cv1
constexpr int kOutBytes = 128 * 64;            // one 32-channel chunk of an output tile (SWIZZLE_64B rows)
// input patch: x rows 4 h0 - 3 .., columns 4 w0 - 8 .. (TMA wants the innermost start on a 16-byte boundary; the
// first column a tap reads is 4 w0 - 3 = patch column 5, the last 4 w0 + 31 = patch column 39)
constexpr int kXCols = 40, kXRows = 67;
constexpr int kXLoadBytes = 3 * kXRows * kXCols * 2;
constexpr int kXBytes = 16128;
constexpr int kOffPatch = 0;
constexpr int kOffW1 = kOffPatch + kPatchBytes;
constexpr int kOffW2 = kOffW1 + 9 * kTapBytes;
constexpr int kOffOut = kOffW2 + 2 * kTapBytes;
constexpr int kOffX = kOffOut + 2 * kOutBytes;
constexpr int kOffBfrag = kOffX + 2 * kXBytes;   // conv1's B fragments: [3 k-steps][8 n-tiles][32 lanes] x 8 bytes
constexpr int kBfragBytes = 3 * 8 * 32 * 8;
constexpr int kOffAffine = kOffBfrag + kBfragBytes;  // scale1, shift1, scale2, shift2: 4 x 128 floats, pre-halved
constexpr int kOffBars = kOffAffine + 4 * kC * 4;
constexpr int kNumBars = 4 + 4 + 2 + 2 + 2 + 2 + 2 + 1;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kPatchBytes % 1024 == 0 && kOffP01 % 1024 == 0 && kOffP10 % 1024 == 0 && kOffP11 % 1024 == 0,
              "planes start on swizzle-atom boundaries");
static_assert(kOffW1 % 1024 == 0 && kOffW2 % 1024 == 0 && kOffOut % 1024 == 0 && kOffX % 128 == 0, "operand alignment");
static_assert(kXLoadBytes <= kXBytes && kXBytes % 128 == 0, "input patch buffer");
static_assert(kSmemBytes <= 227 * 1024, "stem_fused shared-memory plan exceeds one CTA");
static_assert(kThreads * (kRegsLaunch + 8) > 65536 && kThreads * kRegsLaunch <= 65536 &&
                  kThreads * kRegsLaunch >= 128 * kRegsCtrl + 256 * kRegsLaunch + kProdWarps * 32 * kRegsProd,
              "register pool");

__device__ __forceinline__ int plane_offset(int pr, int pc) {
  return pr == 0 ? (pc == 0 ? kOffP00 : kOffP01) : (pc == 0 ? kOffP10 : kOffP11);
}

// Waits local to this kernel: ptx.cuh's mbar_wait calls a __noinline__ time-out reporter, and ONE such ABI call in the
// kernel makes ptxas hold every role to the smallest setmaxnreg value (measured: the producers were capped at the
// control warps' 88 registers and spilled their weight fragments).  Here a time-out leaves its marks in the mapped
// debug buffer (if one is set) and traps.
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

// mbarrier wait that also acquires what OTHER CTAs of the cluster released before arriving
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > (1ll << 31)) wait_timeout(addr, parity | 0x100u);
  }
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct StemFusedParams {
  int tiles_w, tiles_h, total_tiles;  // per image: tiles_w x tiles_h tiles of 8 x 16; total over the batch
  int out_c_off;
  int reverse;
  int debug;  // HGR_STEM_DEBUG=1: a timed-out wait leaves its marks in mapped host memory (printed by the launcher)
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

__global__ void __launch_bounds__(kThreads, 1)
stem_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
                  const StemFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* plane_full = bars;               // [4] leader's: every producer warp of both CTAs has stored its share
  uint64_t* plane_empty = bars + 4;          // [4] the taps that read the plane have retired (both CTAs)
  uint64_t* x_full = bars + 8;               // [2] this CTA's input patch has landed
  uint64_t* x_empty = bars + 10;             // [2] every producer warp of this CTA is done with the buffer
  uint64_t* acc_full = bars + 12;            // [2] G1 of an item has retired (both CTAs)
  uint64_t* a2_ready = bars + 14;            // [2] leader's: both CTAs' a2 tiles are in tensor memory
  uint64_t* acc2_full = bars + 16;           // [2] G2 of an item has retired (both CTAs)
  uint64_t* w_bar = bars + 18;               // leader's: both halves of conv2's and cv1's weights are resident
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
    // plane_full / a2_ready: every warp arrives on ITS OWN CTA's barrier (a remote release-arrive costs the arriving
    // warp a GPU-scope membar, measured ~800 cycles); the peer's otherwise idle warp 1 waits on the peer's copy and
    // forwards ONE arrival to the leader's
    const uint32_t fwd = cluster_ctarank() == 0 ? 1u : 0u;
    for (int i = 0; i < 4; ++i) {
      mbar_init(&plane_full[i], kProdWarps + fwd);
      mbar_init(&plane_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], kProdWarps);
      mbar_init(&acc_full[i], 1);
      mbar_init(&a2_ready[i], 4 + fwd);  // one arrival per epilogue warp of the group (+ the peer's forwarded one)
      mbar_init(&acc2_full[i], 1);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  // SiLU is evaluated on h = x / 2, so the 1/2 is folded into both affines (as in gemm_tcgen05.cu's load_affine)
  for (int i = threadIdx.x; i < kC; i += kThreads) {
    s_aff[i] = 0.5f * (p.scale1 ? p.scale1[i] : 1.0f);
    s_aff[kC + i] = 0.5f * (p.shift1 ? p.shift1[i] : 0.0f);
    s_aff[2 * kC + i] = 0.5f * (p.scale2 ? p.scale2[i] : 1.0f);
    s_aff[3 * kC + i] = 0.5f * (p.shift2 ? p.shift2[i] : 0.0f);
  }
  // conv1's weights -> B fragments of the 48-deep layout (see the header): entry (k-step s, n-tile nt, lane (g, t)) =
  // {W[n][16 s + 2 t], W[n][16 s + 2 t + 1]}, {W[n][16 s + 2 t + 8], W[n][16 s + 2 t + 9]}, n = 8 nt + g
  for (int e = threadIdx.x; e < 3 * 8 * 32; e += kThreads) {
    const int ln = e & 31, nt = (e >> 5) & 7, ks = e >> 8;
    const int n = nt * 8 + (ln >> 2), tt = ln & 3;
    const float hs = 0.5f * __ldg(p.shift0 + n);
    const __nv_bfloat16 hi = __float2bfloat16_rn(hs);
    const __nv_bfloat16 lo = __float2bfloat16_rn(hs - __bfloat162float(hi));
    uint32_t v[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      unsigned short w[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kk = 2 * tt + j + 8 * half;  // k within the k-step = 4 kh + slot
        const int kh = kk >> 2, slot = kk & 3;
        __nv_bfloat16 val = __float2bfloat16_rn(0.0f);
        if (kh < 3 && slot >= 1)
          val = __float2bfloat16_rn(0.5f * __bfloat162float(p.w0[n * 32 + (kh * 3 + slot - 1) * 3 + ks]));  // exact
        else if (ks == 0 && kh == 3 && slot < 2)
          val = slot == 0 ? hi : lo;
        w[j] = __bfloat16_as_ushort(val);
      }
      v[half] = (uint32_t)w[0] | ((uint32_t)w[1] << 16);
    }
    *reinterpret_cast<uint2*>(smem + kOffBfrag + e * 8) = make_uint2(v[0], v[1]);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t cta_rank = cluster_ctarank();
  const int total_items = (p.total_tiles + 1) / 2;  // work items of the pair-wide walk
  const int first = blockIdx.x / 2, stride = gridDim.x / 2;
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

  if (warp < 4) {
    setmaxnreg_dec<kRegsCtrl>();
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
        }
      }
    } else if (warp == 1 && cta_rank == 0) {
      // ================= MMA issuer (leader): G1 of item i plane by plane, then G2 of item i - 1 =================
      constexpr uint32_t idesc = umma_idesc_bf16(256, kC);
      auto issue_g2 = [&](int it) {
        const int g = it & 1;
        wait_cluster(&a2_ready[g], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t tmem_a = tmem_base + g * kC;          // a2 over the first 64 columns of the G1 stage
        const uint32_t tmem_d = tmem_base + 2 * kC + g * kC;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kC / 16; ++k)
            umma_bf16_ts_2sm(tmem_d, tmem_a + 8 * k,
                             umma_desc_sw128(smem_u32(smem + kOffW2 + (k >> 2) * kTapBytes), 1024) + 2 * (k & 3), idesc,
                             k != 0 ? 1u : 0u);
          umma_commit_2sm(&acc2_full[g], 0b11);
        }
        __syncwarp();
      };
      wait_cluster(w_bar, 0);
      int iter = 0;
      for (int item = first; item < total_items; item += stride, ++iter) {
        const int g = iter & 1;
        // stage g was last used by item iter - 2: its accumulator was read by E1 before a2_ready (waited on in
        // issue_g2(iter - 2)) and its a2 by G2(iter - 2), issued before this point: the tensor pipe keeps the order
        const uint32_t tmem_d = tmem_base + g * kC;
        const uint32_t patch = smem_u32(smem + kOffPatch);
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
          const int pr = pl < 2 ? 1 : 0, pc = (pl & 1) ? 0 : 1;
          const int pw = 8 + pc;  // plane width in pixels
          wait_cluster(&plane_full[pl], iter & 1);
          tc_fence_after();
          if (elect_one_sync()) {
#pragma unroll
            for (int a = 0; a < (pr ? 2 : 1); ++a)
#pragma unroll
              for (int b = 0; b < (pc ? 2 : 1); ++b) {
                const int kh = pr ? 2 * a : 1, kw = pc ? 2 * b : 1;
                const uint32_t a_addr = patch + plane_offset(pr, pc) + ((kh == 2 ? pw : 0) + (kw == 2 ? 1 : 0)) * 128;
                const uint64_t a_base = umma_desc_sw128(a_addr, pw * 128);
                const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffW1 + (kh * 3 + kw) * kTapBytes), 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ss_2sm(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (pl | a | b | k) != 0 ? 1u : 0u);
              }
            umma_commit_2sm(&plane_empty[pl], 0b11);
            if (pl == 3) umma_commit_2sm(&acc_full[g], 0b11);
          }
          __syncwarp();
        }
        if (iter >= 1) issue_g2(iter - 1);
      }
      if (iter >= 1) issue_g2(iter - 1);
    } else if (warp == 1) {
      // ================= forwarder (peer CTA): local plane_full / a2_ready completions -> one arrival on the leader's =================
      if (lane == 0) {
        int n_items = 0;
        for (int item = first; item < total_items; item += stride) ++n_items;
        int pi = 0, ai = 0;  // forwarded plane events (4 per item, in production order) and a2 events (1 per item)
        long long t0 = clock64();
        while (pi < 4 * n_items || ai < n_items) {
          bool progress = false;
          if (pi < 4 * n_items && mbar_try_wait(&plane_full[pi & 3], (pi >> 2) & 1)) {
            mbar_arrive_cluster(&plane_full[pi & 3], 0);
            ++pi;
            progress = true;
          }
          if (ai < n_items && mbar_try_wait(&a2_ready[ai & 1], (ai >> 1) & 1)) {
            mbar_arrive_cluster(&a2_ready[ai & 1], 0);
            ++ai;
            progress = true;
          }
          if (progress) t0 = clock64();
          else if (clock64() - t0 > (1ll << 31)) wait_timeout(smem_u32(&plane_full[pi & 3]), 0x200u | (uint32_t)pi);
        }
      }
      __syncwarp();
    }
  } else if (warp < 12) {
    // ================= epilogue groups: group g owns the stages g of G1 and G2 and staging chunk g =================
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
      wait_cta(&acc_full[group], ph);
      tc_fence_after();
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
      if (lane == 0) {
        mbar_arrive(&a2_ready[group]);
      }

      // ---------------- E2: g tile, one 32-channel chunk at a time -> staging -> TMA store ----------------
      wait_cta(&acc2_full[group], ph);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t acc[32], packed[16];
        tmem_ld_32x32b_x32(t_row + 2 * kC + group * kC + j * 32, acc);
        tmem_ld_wait();
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
      tc_fence_before();
    }
    if (gtid == 0) tma_store_wait_all();
  } else {
    // ================= conv1 producers: the a1 patch of the tile, m-tile by m-tile, into the parity planes =================
    setmaxnreg_inc<kRegsProd>();
    const int pwarp = warp - 12;
    const int g = lane >> 2, t = lane & 3;
    // row kh = 3 (lanes t >= 2, registers a2 / a3) carries no input: zeros, except the two 1.0 slots of k-step 0
    const uint32_t a_keep = (t >> 1) ? 0u : 0xFFFFFFFFu;
    const uint32_t a_one = t == 2 ? 0x3F803F80u : 0u;
    // this lane's A registers: (input row kh = 2 h + (t >> 1), column slots 2 (t & 1), + 1), h = 0 for a0 / a1
    const uint32_t lane_off = (t >> 1) * (kXCols * 2) + (t & 1) * 4;
    const uint32_t bfrag_addr = smem_u32(smem + kOffBfrag) + lane * 8;
    auto signal_plane = [&](int pl) {
      fence_proxy_async_smem();  // the stores above are read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&plane_full[pl]);  // own CTA's barrier (the peer's is forwarded by its warp 1)
    };
    // This warp's m-tiles, fixed for the whole kernel: mtg = pwarp + 8 k (k = 0..3) and, for one half of the warps
    // per item (alternating, so that the 36 m-tiles split evenly), mtg = 32 + (pwarp & 3).  Per m-tile two packed
    // words, computed once: {byte offset of pixel row g in the input patch | the same for row g + 8 << 16} and
    // {byte offset of pixel row g in the a1 patch with its swizzle phase applied | flags << 20}; row g + 8 lies 1024
    // bytes further on, in the same swizzle phase.
    uint32_t tab_x[5], tab_p[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int mtg = k < 4 ? pwarp + 8 * k : 32 + (pwarp & 3);
      const int pl = mtg < 10 ? 0 : (mtg < 19 ? 1 : (mtg < 28 ? 2 : 3));
      const int pr = pl < 2 ? 1 : 0, pc = (pl & 1) ? 0 : 1;
      const int pw = 8 + pc;
      const int npix = (16 + pr) * pw;
      const int mt = mtg - (pl == 0 ? 0 : (pl == 1 ? 10 : (pl == 2 ? 19 : 28)));
      const int i0 = mt * 16 + g, i1 = i0 + 8;
      const int j0 = min(i0, npix - 1), j1 = min(i1, npix - 1);
      const int y0 = pc ? (j0 * 57) >> 9 : j0 >> 3, y1 = pc ? (j1 * 57) >> 9 : j1 >> 3;
      const int x0 = j0 - y0 * pw, x1 = j1 - y1 * pw;
      // byte offset of (channel 0, input row kh = 0, column slot 0 = kw - 1) of the two pixels in the input patch
      const uint32_t pb0 = lane_off + 2 * ((4 * y0 - 2 * pr + 2) * kXCols + 4 * x0 - 2 * pc + 6);
      const uint32_t pb1 = lane_off + 2 * ((4 * y1 - 2 * pr + 2) * kXCols + 4 * x1 - 2 * pc + 6);
      tab_x[k] = pb0 | (pb1 << 16);
      // SWIZZLE_128B: 16-byte chunk nt of pixel row i sits at chunk nt ^ (i & 7) (planes start on 1024-byte lines)
      const uint32_t r0 = (uint32_t)(kOffPatch + plane_offset(pr, pc) + t * 4 + i0 * 128) ^ ((i0 & 7) << 4);
      const uint32_t flags = (i0 < npix ? 1u : 0u) | (i1 < npix ? 2u : 0u) |             // rows inside the plane
                             ((pr && y0 == 0) ? 4u : 0u) | ((pr && y1 == 0) ? 8u : 0u) |  // a1 row -1 when h0 == 0
                             ((pc && x0 == 0) ? 16u : 0u) | ((pc && x1 == 0) ? 32u : 0u) |  // a1 column -1 when w0 == 0
                             ((uint32_t)pl << 6);
      tab_p[k] = r0 | (flags << 20);
    }
    static_assert(kPatchBytes < (1 << 20) && kXBytes < (1 << 16), "table packing");
    const uint32_t patch_base = smem_u32(smem);
    int iter = 0;
    for (int item = first; item < total_items; item += stride, ++iter) {
      int w0, h0, n0;
      coords(item, w0, h0, n0);
      const int xb = iter & 1;
      wait_cta(&x_full[xb], (iter >> 1) & 1);
      const uint32_t xs = smem_u32(smem + kOffX + xb * kXBytes);
      const uint32_t eph = (iter & 1) ^ 1;
      const uint32_t edge = (h0 == 0 ? (4u | 8u) : 0u) | (w0 == 0 ? (16u | 32u) : 0u);
      const int nk = (((pwarp >> 2) ^ iter) & 1) ? 4 : 5;
      int done = 0;  // planes [0, done) have been signalled by this warp
      // a rolled loop (five copies of the body would not fit the instruction cache) over a table that lives in
      // registers: the body reads entry 0 and the table rotates by one entry per iteration, five times per item
#pragma unroll 1
      for (int k = 0; k < 5; ++k) {
        const uint32_t tx = tab_x[0], tp = tab_p[0];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tab_x[q] = tab_x[q + 1];
          tab_p[q] = tab_p[q + 1];
        }
        tab_x[4] = tx;
        tab_p[4] = tp;
        if (k >= nk) break;
        const uint32_t pb0 = xs + (tx & 0xFFFFu), pb1 = xs + (tx >> 16);
        uint32_t a[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          a[c][0] = lds32(pb0 + c * (kXRows * kXCols * 2));
          a[c][1] = lds32(pb1 + c * (kXRows * kXCols * 2));
          a[c][2] = lds32(pb0 + c * (kXRows * kXCols * 2) + 2 * kXCols * 2) & a_keep;
          a[c][3] = lds32(pb1 + c * (kXRows * kXCols * 2) + 2 * kXCols * 2) & a_keep;
        }
        a[0][2] |= a_one;  // (c = 0, kh = 3) slots 0, 1: 1.0 against shift / 2 (hi, lo)
        a[0][3] |= a_one;
        // all MMAs, then all thirty-two MUFU ops back to back
        float d[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const uint2 b = lds64(bfrag_addr + (c * 8 + nt) * 256);
            mma_bf16_16816(d[nt], a[c], b.x, b.y);
          }
        float th[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) th[nt][i] = tanh_approx(d[nt][i]);  // d is (conv + shift) / 2 already
        const uint32_t flags = tp >> 20;
        const int pl = (int)(flags >> 6);
        while (done < pl) signal_plane(done++);
        wait_cta(&plane_empty[pl], eph);  // the previous item's taps on this plane have retired
        const uint32_t r0 = patch_base + (tp & 0xFFFFFu), r1 = r0 + 1024;
        const bool st0 = flags & 1u, st1 = flags & 2u;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint32_t v0 = pack_bf16x2(fmaf(d[nt][0], th[nt][0], d[nt][0]), fmaf(d[nt][1], th[nt][1], d[nt][1]));
          const uint32_t v1 = pack_bf16x2(fmaf(d[nt][2], th[nt][2], d[nt][2]), fmaf(d[nt][3], th[nt][3], d[nt][3]));
          if (st0) sts32(r0 ^ (nt << 4), v0);
          if (st1) sts32(r1 ^ (nt << 4), v1);
        }
        // a1 pixels outside the map (row -1 / column -1) are conv2's zero padding: only tiles on the top / left edge
        const uint32_t z = flags & edge;
        if (__any_sync(0xffffffffu, z != 0)) {
          const bool z0 = st0 && (z & (4u | 16u)), z1 = st1 && (z & (8u | 32u));
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            if (z0) sts32(r0 ^ (nt << 4), 0u);
            if (z1) sts32(r1 ^ (nt << 4), 0u);
          }
        }
      }
      while (done < 4) signal_plane(done++);
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[xb]);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while its peer may still read its shared memory or arrive on its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace

bool stem_fused_supported(int S) { return S % 64 == 0 && S >= 64 && stem_chain_supported(S / 2, S / 2); }

// x: (B, 3, S, S) bf16 NCHW; w0 [64][32] bf16 + shift0 (conv1, packed as for conv1.cu); w1 [128][3][3][64],
// w2 [128][128] bf16; out: channel slice [out_coff, +128) of a (B, S / 4, S / 4, out_ctot) buffer.
int run_stem_fused(const void* x, int B, int S, const void* w0, const float* shift0, const void* w1,
                   const float* scale1, const float* shift1, const void* w2, const float* scale2, const float* shift2,
                   void* out, int out_ctot, int out_coff, int reverse, int num_sms, cudaStream_t stream) {
  if (!stem_fused_supported(S)) {
    set_error("stem_fused: image side %d must be a multiple of 64", S);
    return -1;
  }
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) {
    set_error("stem_fused: the input batch must be 16-byte aligned");
    return -1;
  }
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
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
  StemFusedParams p;
  p.tiles_w = Wo / kTW;
  p.tiles_h = Ho / kTH;
  p.total_tiles = p.tiles_w * p.tiles_h * B;
  p.out_c_off = out_coff;
  p.reverse = reverse;
  p.debug = getenv("HGR_STEM_DEBUG") ? atoi(getenv("HGR_STEM_DEBUG")) : 0;
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
  HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, stem_fused_kernel, tx, tw, tw2, to, p));
  if (p.debug) {
    cudaError_t e = cudaStreamSynchronize(stream);
    fprintf(stderr, "hgr stem_fused debug %d: sync -> %s; marks %x %u %u %u %x | %u %u %u %u\n", p.debug, cudaGetErrorName(e),
            dbg_host[0], dbg_host[1], dbg_host[2], dbg_host[3], dbg_host[4], dbg_host[8], dbg_host[9], dbg_host[10],
            dbg_host[11]);
  }
  return 0;
}

}  // namespace hgr
