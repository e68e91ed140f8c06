// encoder.conv2 -> encoder.cspelan1.cv1 as ONE kernel, second version: the input patch of the stride-2 convolution is
// staged ONCE per tile and the tensor between the two layers lives in tensor memory
// (reference model/gelan.py:156 `conv2 = Conv(64, 128, 3, 2)`, :127 GELANBlock.cv1 = Conv(128, 128, 1, 1);
// Conv.forward :56 = SiLU(BN(conv(x)))):
//
//     a2 = SiLU(BN1(conv3x3_s2(a1)))        G1 (K = 9 taps x 64 ch), E1
//     g  = SiLU(BN2(conv1x1(a2)))           G2 (K = 128),            E2
//
// conv_chain.cu loads the 128-pixel A tile of every tap through TMA (9 x 16 KiB) next to the tap's weights (9 x 8 KiB
// per CTA of the pair): 216 KiB per tile and CTA, and the kernel is bound by exactly that - the L2 can deliver ~42
// B/clk to each SM when all of them pull at once (919 TFLOP/s, half of what the MMAs could do).  Here a tile is 8 x 16
// output pixels and its (17 x 33)-pixel input patch arrives as the FOUR PARITY PLANES of the space-to-depth view
// (row parity x column parity: 17x9, 17x8, 16x9 and 16x8 pixels, 72 KiB together, out-of-image pixels zero-filled by
// TMA = the convolution's padding).  Tap (kh, kw) is plane (kh != 1, kw != 1) read through a shifted UMMA descriptor:
//   start = plane + ((kh == 2) * plane_width + (kw == 2)) * 128 B, 8-row groups plane_width * 128 B apart,
// the same absolute-address swizzle argument as conv3x3_halo_kernel.  Operand traffic per tile and CTA: 72 + 72 KiB.
// The a2 tile goes from the accumulator through SiLU to bf16 back INTO TENSOR MEMORY (over the accumulator columns
// its thread has consumed) and is the A operand of G2 (cta_group::2 MMA with A in TMEM), so no shared memory is spent
// on it; what is left holds two patch buffers, a four-deep weight ring, cv1's weights and one 16 KiB output staging
// chunk per epilogue group.
//
// CTA pairs (cta_group::2, M = 256) as in conv_chain.cu: each CTA stages its own tile's patch and half of every
// weight tile, the leader's MMA warp issues G1 of item i and then G2 of item i - 1.  TMEM: two G1 accumulator stages
// (columns 0-255; a2 over the first 64 columns of its stage) and two G2 stages (columns 256-511).  The G1 stage of
// item i is reused by item i + 2: its a2 is read by G2 of item i, which the same thread issues earlier, and the
// tensor pipe executes in issue order.
#include <cstdio>
#include <cstring>

#include "epilogue_math.cuh"
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 384;
constexpr int kC = 128;                        // channels of a2 and of g
constexpr int kTW = 8, kTH = 16;               // output tile
// parity planes [pr][pc] of the input patch: rows x columns of 128-byte pixels
constexpr int kPlaneW[2][2] = {{8, 9}, {8, 9}};
constexpr int kPlaneH[2][2] = {{16, 16}, {17, 17}};
constexpr int kOffP00 = 0;                                   // 16 x 8
constexpr int kOffP01 = kOffP00 + 16 * 8 * 128;              // 16 x 9
constexpr int kOffP10 = kOffP01 + 16 * 9 * 128;              // 17 x 8
constexpr int kOffP11 = kOffP10 + 17 * 8 * 128;              // 17 x 9 (padded to 1024)
constexpr int kPatchLoadBytes = (16 * 8 + 16 * 9 + 17 * 8 + 17 * 9) * 128;  // what the four boxes deliver
constexpr int kPatchBytes = kOffP11 + 20 * 1024;
constexpr int kBBytes = 64 * 128;              // this CTA's 64 weight rows of one tap
constexpr int kBStages = 4;
constexpr int kW2Bytes = 2 * kBBytes;          // this CTA's 64 rows of cv1's weights, two k-blocks
constexpr int kOutBytes = 128 * 128;           // one 64-channel chunk of an output tile
constexpr int kOffPatch = 0;
constexpr int kOffB = 2 * kPatchBytes;
constexpr int kOffW2 = kOffB + kBStages * kBBytes;
constexpr int kOffOut = kOffW2 + kW2Bytes;
constexpr int kOffAffine = kOffOut + 2 * kOutBytes;  // scale1, shift1, scale2, shift2: 4 x 128 floats, pre-halved
constexpr int kOffBars = kOffAffine + 4 * kC * 4;
constexpr int kNumBars = 4 + 2 * kBStages + 7;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kPatchBytes % 1024 == 0 && kOffP01 % 1024 == 0 && kOffP10 % 1024 == 0 && kOffP11 % 1024 == 0,
              "planes start on swizzle-atom boundaries");
static_assert(kSmemBytes <= 227 * 1024, "stem_chain shared-memory plan exceeds one CTA");

__device__ __forceinline__ int plane_offset(int pr, int pc) {
  return pr == 0 ? (pc == 0 ? kOffP00 : kOffP01) : (pc == 0 ? kOffP10 : kOffP11);
}

// mbarrier wait that also acquires what OTHER CTAs of the cluster released before arriving
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
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
    if (clock64() - t0 > (1ll << 31)) {
      printf("hgr: stem_chain mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
      __trap();
    }
  }
}

struct StemParams {
  int tiles_w, tiles_h, total_tiles;  // per image: tiles_w x tiles_h tiles of 8 x 16; total over the batch
  int out_c_off;
  int reverse;
  const float* scale1;
  const float* shift1;
  const float* scale2;
  const float* shift2;
};

// 64 values of one pixel row -> SiLU(affine) -> 32 packed bf16 pairs
__device__ __forceinline__ void activate_pack(const uint32_t (&acc)[64], const float* s_scale, const float* s_shift,
                                              uint32_t (&packed)[32]) {
#pragma unroll
  for (int e = 0; e < 64; e += 4) {
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
stem_chain_kernel(const __grid_constant__ CUtensorMap tmP00, const __grid_constant__ CUtensorMap tmP01,
                  const __grid_constant__ CUtensorMap tmP10, const __grid_constant__ CUtensorMap tmP11,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmO, const StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* patch_full = bars;                      // [2] leader's: both CTAs' patches have landed
  uint64_t* patch_empty = bars + 2;                 // [2] G1 of the item has retired (both CTAs)
  uint64_t* b_full = bars + 4;                      // [kBStages] leader's
  uint64_t* b_empty = bars + 4 + kBStages;          // [kBStages] both CTAs
  uint64_t* acc_full = bars + 4 + 2 * kBStages;     // [2] G1 of an item has retired (both CTAs)
  uint64_t* a2_ready = acc_full + 2;                // [2] leader's: both CTAs' a2 tiles are in tensor memory
  uint64_t* acc2_full = a2_ready + 2;               // [2] G2 of an item has retired (both CTAs)
  uint64_t* w2_bar = acc2_full + 2;                 // leader's: both halves of cv1's weights are resident
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  float* s_aff = reinterpret_cast<float*>(smem + kOffAffine);

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmP00);
    prefetch_tensormap(&tmP01);
    prefetch_tensormap(&tmP10);
    prefetch_tensormap(&tmP11);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmW2);
    prefetch_tensormap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&patch_full[i], 1);
      mbar_init(&patch_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&a2_ready[i], 4 * 2);  // one arrival per epilogue warp of the group, both CTAs
      mbar_init(&acc2_full[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(w2_bar, 1);
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

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own patch planes, own half of the weights =================
    if (elect_one_sync()) {
      if (cta_rank == 0) mbar_expect_tx(w2_bar, 2 * kW2Bytes);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
        tma_load_2d_2sm(smem + kOffW2 + kb * kBBytes, &tmW2, w2_bar, kb * 64, (int)cta_rank * (kC / 2));
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int item = first; item < total_items; item += stride, ++iter) {
        int w0, h0, n0;
        coords(item, w0, h0, n0);
        const int pb = iter & 1;
        uint8_t* patch = smem + kOffPatch + pb * kPatchBytes;
        mbar_wait(&patch_empty[pb], ((iter >> 1) & 1) ^ 1);
        if (cta_rank == 0) mbar_expect_tx(&patch_full[pb], 2 * kPatchLoadBytes);
        // space-to-depth view (c2 = pc * 64 + c, x, pr, y, n): input row 2 y + pr, column 2 x + pc.  Row parity 1 /
        // column parity 1 planes start one block earlier (taps kh = 0 / kw = 0 reach back to row 2 h0 - 1 / column 2 w0 - 1)
        tma_load_5d_2sm(patch + kOffP00, &tmP00, &patch_full[pb], 0, w0, 0, h0, n0);
        tma_load_5d_2sm(patch + kOffP01, &tmP01, &patch_full[pb], 64, w0 - 1, 0, h0, n0);
        tma_load_5d_2sm(patch + kOffP10, &tmP10, &patch_full[pb], 0, w0, 1, h0 - 1, n0);
        tma_load_5d_2sm(patch + kOffP11, &tmP11, &patch_full[pb], 64, w0 - 1, 1, h0 - 1, n0);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&b_empty[stage], phase ^ 1);
          if (cta_rank == 0) mbar_expect_tx(&b_full[stage], 2 * kBBytes);
          tma_load_2d_2sm(smem + kOffB + stage * kBBytes, &tmW, &b_full[stage], tap * 64, (int)cta_rank * (kC / 2));
          if (++stage == kBStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ================= MMA issuer (leader): G1 of item i, then G2 of item i - 1 =================
    constexpr uint32_t idesc = umma_idesc_bf16(256, kC);
    auto issue_g2 = [&](int it) {
      const int g = it & 1;
      mbar_wait_cluster(&a2_ready[g], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_a = tmem_base + g * kC;          // a2 over the first 64 columns of the G1 stage
      const uint32_t tmem_d = tmem_base + 2 * kC + g * kC;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < kC / 16; ++k)
          umma_bf16_ts_2sm(tmem_d, tmem_a + 8 * k,
                           umma_desc_sw128(smem_u32(smem + kOffW2 + (k >> 2) * kBBytes), 1024) + 2 * (k & 3), idesc,
                           k != 0 ? 1u : 0u);
        umma_commit_2sm(&acc2_full[g], 0b11);
      }
      __syncwarp();
    };
    mbar_wait_cluster(w2_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int item = first; item < total_items; item += stride, ++iter) {
      const int g = iter & 1;
      // stage g was last used by item iter - 2: its accumulator was read by E1 before a2_ready (waited on in
      // issue_g2(iter - 2)) and its a2 by G2(iter - 2), issued before this point: the tensor pipe keeps the order
      mbar_wait(&patch_full[g], (iter >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + g * kC;
      const uint32_t patch = smem_u32(smem + kOffPatch + g * kPatchBytes);
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap - kh * 3;
        const int pr = kh != 1, pc = kw != 1;
        const int pw = 8 + pc;  // plane width in pixels
        const uint32_t a_addr = patch + plane_offset(pr, pc) + ((kh == 2 ? pw : 0) + (kw == 2 ? 1 : 0)) * 128;
        mbar_wait(&b_full[stage], phase);
        tc_fence_after();
        const uint64_t a_base = umma_desc_sw128(a_addr, pw * 128);
        const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffB + stage * kBBytes), 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_2sm(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&b_empty[stage], 0b11);
          if (tap == 8) {
            umma_commit_2sm(&patch_empty[g], 0b11);
            umma_commit_2sm(&acc_full[g], 0b11);
          }
        }
        __syncwarp();
        if (++stage == kBStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (iter >= 1) issue_g2(iter - 1);
    }
    if (iter >= 1) issue_g2(iter - 1);
  } else if (warp >= 4) {
    // ================= epilogue groups: group g owns the stages g of G1 and G2 and staging chunk g =================
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - 128 - group * 128;
    const uint32_t bar_id = 1 + group;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint8_t* stage_out = smem + kOffOut + group * kOutBytes;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int iter = 0;
    for (int item = first; item < total_items; item += stride, ++iter) {
      if ((iter & 1) != group) continue;
      const uint32_t ph = (iter >> 1) & 1;
      int w0, h0, n0;
      coords(item, w0, h0, n0);

      // ---------------- E1: a2 tile -> tensor memory (A operand of G2) ----------------
      mbar_wait(&acc_full[group], ph);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint32_t acc[64], packed[32];
        tmem_ld_32x32b_x32(t_row + group * kC + j * 64, acc);
        tmem_ld_32x32b_x32(t_row + group * kC + j * 64 + 32, acc + 32);
        tmem_ld_wait();
        activate_pack(acc, s_aff + j * 64, s_aff + kC + j * 64, packed);
        // channels 64 j .. 64 j + 63 -> columns 32 j .. 32 j + 31 of the stage: inside what this thread has consumed
        tmem_st_32x32b_x32(t_row + group * kC + j * 32, packed);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank == 0) mbar_arrive(&a2_ready[group]);
        else mbar_arrive_cluster(&a2_ready[group], 0);
      }

      // ---------------- E2: g tile, one 64-channel chunk at a time -> staging -> TMA store ----------------
      mbar_wait(&acc2_full[group], ph);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint32_t acc[64], packed[32];
        tmem_ld_32x32b_x32(t_row + 2 * kC + group * kC + j * 64, acc);
        tmem_ld_32x32b_x32(t_row + 2 * kC + group * kC + j * 64 + 32, acc + 32);
        tmem_ld_wait();
        activate_pack(acc, s_aff + 2 * kC + j * 64, s_aff + 3 * kC + j * 64, packed);
        if (gtid == 0) tma_store_wait_read<0>();  // the previous chunk has left the staging buffer
        bar_sync(bar_id, 128);
#pragma unroll
        for (int v = 0; v < 8; ++v)
          *reinterpret_cast<uint4*>(stage_out + row * 128 + ((static_cast<uint32_t>(v) ^ sw) << 4)) =
              make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
        fence_proxy_async_smem();
        bar_sync(bar_id, 128);
        if (gtid == 0) {
          tma_store_4d(&tmO, stage_out, p.out_c_off + j * 64, w0, h0, n0);
          tma_store_commit();
        }
      }
      tc_fence_before();
    }
    if (gtid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while its peer may still read its shared memory or arrive on its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace

struct StemChainOp {
  CUtensorMap p00, p01, p10, p11, w, w2, o;
  StemParams p;
};

bool stem_chain_supported(int H, int W) { return H % 32 == 0 && W % 16 == 0 && H >= 32 && W >= 16; }

// in: a1 (B, H, W, 64) bf16; w1 [128][3][3][64], w2 [128][128] bf16; out: channel slice [out_coff, +128) of a
// (B, H / 2, W / 2, out_ctot) buffer.
int run_stem_chain(const void* in, int B, int H, int W, const void* w1, const float* scale1, const float* shift1,
                   const void* w2, const float* scale2, const float* shift2, void* out, int out_ctot, int out_coff,
                   int reverse, int num_sms, cudaStream_t stream) {
  if (!stem_chain_supported(H, W)) {
    set_error("stem_chain: input map %d x %d must tile into 32 x 16 blocks", H, W);
    return -1;
  }
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(stem_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  StemChainOp op;
  memset(&op, 0, sizeof(op));
  const int Ho = H / 2, Wo = W / 2;
  {
    // space-to-depth view of a1: (c2 = pc * 64 + c, x = W / 2, pr = 2, y = H / 2, n)
    const uint64_t dims[5] = {128, (uint64_t)Wo, 2, (uint64_t)Ho, (uint64_t)B};
    const uint64_t pix = 64 * 2;
    const uint64_t strides[4] = {2 * pix, pix * W, 2 * pix * W, pix * W * H};
    CUtensorMap* maps[2][2] = {{&op.p00, &op.p01}, {&op.p10, &op.p11}};
    for (int pr = 0; pr < 2; ++pr)
      for (int pc = 0; pc < 2; ++pc) {
        const uint32_t box[5] = {64, (uint32_t)kPlaneW[pr][pc], 1, (uint32_t)kPlaneH[pr][pc], 1};
        if (int r = make_tensor_map_bf16(maps[pr][pc], in, 5, dims, strides, box)) return r;
      }
  }
  {
    const uint64_t dims[2] = {576, (uint64_t)kC};
    const uint64_t strides[1] = {576 * 2};
    const uint32_t box[2] = {64, (uint32_t)(kC / 2)};
    if (int r = make_tensor_map_bf16(&op.w, w1, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)kC, (uint64_t)kC};
    const uint64_t strides[1] = {(uint64_t)kC * 2};
    const uint32_t box[2] = {64, (uint32_t)(kC / 2)};
    if (int r = make_tensor_map_bf16(&op.w2, w2, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)out_ctot, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t row = (uint64_t)out_ctot * 2;
    const uint64_t strides[3] = {row, row * Wo, row * Wo * Ho};
    const uint32_t box[4] = {64, (uint32_t)kTW, (uint32_t)kTH, 1};
    if (int r = make_tensor_map_bf16(&op.o, out, 4, dims, strides, box)) return r;
  }
  StemParams& p = op.p;
  p.tiles_w = Wo / kTW;
  p.tiles_h = Ho / kTH;
  p.total_tiles = p.tiles_w * p.tiles_h * B;
  p.out_c_off = out_coff;
  p.reverse = reverse;
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
  HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, stem_chain_kernel, op.p00, op.p01, op.p10, op.p11, op.w, op.w2, op.o, op.p));
  return 0;
}

}  // namespace hgr
