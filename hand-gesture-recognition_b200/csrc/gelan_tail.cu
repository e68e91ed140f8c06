// The tail of the first GELAN block as ONE kernel: cspelan1.cv3.0.cv2 (second conv of the ResBasicBlock, + residual,
// SiLU) -> cspelan1.cv4 (1x1 over the concatenation [y0 | y1 | y2 | y3])
// (reference model/gelan.py:73-87 ResBasicBlock.forward: out = act(x + cv2(cv1(x))); :137-142 GELANBlock.forward:
// cv4(cat(y, 1)); Conv.forward :56 = SiLU(BN(conv(x)))):
//
//     y3 = SiLU(BN_h(conv3x3(t)) + y2)        G1 (K = 9 taps x 64 ch, N = 64), E1 -> tensor memory
//     o  = SiLU(BN_4(conv1x1([y0 y1 y2 y3])))  G2 (K = 256: three 64-channel tiles from shared memory + y3 from
//                                               tensor memory, N = 128),       E2 -> TMA store
//
// As separate launches the 1x1 layer is the most HBM-bound kernel of the forward (0.27 ms at 6.7 TB/s: it reads the
// 256-channel concatenation, 1.2 GB at batch 1024) and the 64-channel 3x3 layer in front of it is shared-memory
// bound (0.22 ms); y3 is written by one and read back by the other and by nothing else.  Here y3 (0.3 GB written +
// 0.3 GB read) never leaves the SM: a tile is 8 x 16 pixels; its (10 x 18)-pixel input halo arrives as one TMA box
// and the nine taps read it through shifted UMMA descriptors with all weights resident (conv3x3_halo_kernel's
// scheme); the activated tile goes back INTO TENSOR MEMORY as bf16 over the accumulator columns its thread has
// consumed and is the TMEM A operand of the last K block of G2, whose first three K blocks (y0, y1, y2: the tile's
// pixels of the concatenation buffer, three TMA boxes) are issued while E1 is still at work.  The residual y2 is read
// by the row owners from global memory as in the stand-alone kernel (its lines were just pulled into L2 by the y2
// box).  Same rounding points and the same K order as the two launches.
//
// CTA pairs (cta_group::2, M = 256: each CTA stages its own tile's operands and half of every weight tile), warp 0
// TMA, warp 1 MMA issue (leader), two epilogue groups of four warps that own G1 stage / G2 stage g of the items
// with iteration parity g.  Tensor memory: G1 stages at columns 0 / 64 (y3 over the first 32 columns of its stage),
// G2 stages at 128 / 256.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "epilogue_math.cuh"
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 384;
constexpr int kCh = 64;                        // channels of t, y2, y3
constexpr int kCo = 128;                       // channels of the block's output
constexpr int kTW = 8, kTH = 16;               // tile
constexpr int kPatchBytes = 10 * 18 * 128;     // what the halo box delivers
constexpr int kPatchStride = 23 * 1024;
constexpr int kPatchStages = 2;
constexpr int kTapBytes = 32 * 128;            // this CTA's 32 weight rows of one tap of the 3x3 layer
constexpr int kW4Bytes = 64 * 128;             // this CTA's 64 rows of one 64-wide K block of cv4's weights
constexpr int kYBytes = 128 * 128;             // one 64-channel tile of the concatenation
constexpr int kYStages = 4;
constexpr int kOutBytes = 128 * 128;           // one 64-channel chunk of an output tile
constexpr int kOffWh = 0;
constexpr int kOffW4 = kOffWh + 9 * kTapBytes;
constexpr int kOffPatch = kOffW4 + 4 * kW4Bytes;
constexpr int kOffY = kOffPatch + kPatchStages * kPatchStride;
constexpr int kOffOut = kOffY + kYStages * kYBytes;
constexpr int kOffAffine = kOffOut + 2 * kOutBytes;  // scale_h, shift_h (64 each), scale_4, shift_4 (128 each), pre-halved
constexpr int kOffBars = kOffAffine + (2 * kCh + 2 * kCo) * 4;
constexpr int kNumBars = 2 * kPatchStages + 2 * kYStages + 2 + 2 + 2 + 2 + 1;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kOffW4 % 1024 == 0 && kOffPatch % 1024 == 0 && kOffY % 1024 == 0 && kOffOut % 1024 == 0, "operand alignment");
static_assert(kSmemBytes <= 227 * 1024, "gelan_tail shared-memory plan exceeds one CTA");

struct TailParams {
  int tiles_w, tiles_h, total_tiles;  // per image: tiles_w x tiles_h tiles of 8 x 16; total over the batch
  int reverse;
  const __nv_bfloat16* res;           // y2: channels [res_c_off, + 64) of the concatenation buffer
  long long res_sn, res_sh, res_sw;   // its strides in elements
  int res_c_off;
  const float* scale_h;
  const float* shift_h;
  const float* scale_4;
  const float* shift_4;
};

__global__ void __launch_bounds__(kThreads, 1)
gelan_tail_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmW4,
                  const __grid_constant__ CUtensorMap tmO, const TailParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* patch_full = bars;                                  // [kPatchStages] leader's: both CTAs' halos have landed
  uint64_t* patch_empty = patch_full + kPatchStages;            // G1 of the item has retired (both CTAs)
  uint64_t* y_full = patch_empty + kPatchStages;                // [kYStages] leader's: both CTAs' tiles have landed
  uint64_t* y_empty = y_full + kYStages;                        // the K block that read the slot has retired
  uint64_t* acc_full = y_empty + kYStages;                      // [2] G1 of an item has retired (both CTAs)
  uint64_t* y3_ready = acc_full + 2;                            // [2] leader's: both CTAs' y3 tiles are in tensor memory
  uint64_t* acc2_full = y3_ready + 2;                           // [2] G2 of an item has retired (both CTAs)
  uint64_t* acc2_empty = acc2_full + 2;                         // [2] leader's: both CTAs' E2 have read the stage
  uint64_t* w_bar = acc2_empty + 2;                             // leader's: both halves of the weights are resident
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  float* s_aff = reinterpret_cast<float*>(smem + kOffAffine);
  float* s_sc_h = s_aff;
  float* s_sh_h = s_aff + kCh;
  float* s_sc_4 = s_aff + 2 * kCh;
  float* s_sh_4 = s_aff + 2 * kCh + kCo;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();  // dynamic smem base not 1024-byte aligned
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmT);
    prefetch_tensormap(&tmY);
    prefetch_tensormap(&tmWh);
    prefetch_tensormap(&tmW4);
    prefetch_tensormap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPatchStages; ++i) {
      mbar_init(&patch_full[i], 1);
      mbar_init(&patch_empty[i], 1);
    }
    for (int i = 0; i < kYStages; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&y3_ready[i], 4 * 2);    // one arrival per epilogue warp of the group, both CTAs
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], 4 * 2);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, 512);
    tmem_relinquish_2sm();
  }
  // SiLU is evaluated on h = x / 2, so the 1/2 is folded into both affines (as in gemm_tcgen05.cu's load_affine)
  for (int i = threadIdx.x; i < kCo; i += kThreads) {
    if (i < kCh) {
      s_sc_h[i] = 0.5f * (p.scale_h ? p.scale_h[i] : 1.0f);
      s_sh_h[i] = 0.5f * (p.shift_h ? p.shift_h[i] : 0.0f);
    }
    s_sc_4[i] = 0.5f * (p.scale_4 ? p.scale_4[i] : 1.0f);
    s_sh_4[i] = 0.5f * (p.shift_4 ? p.shift_4[i] : 0.0f);
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
    // ================= TMA producer (both CTAs): own half of the weights once, then halos and concatenation tiles =================
    if (elect_one_sync()) {
      if (cta_rank == 0) mbar_expect_tx(w_bar, 2 * (9 * kTapBytes + 4 * kW4Bytes));
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d_2sm(smem + kOffWh + tap * kTapBytes, &tmWh, w_bar, tap * 64, (int)cta_rank * (kCh / 2));
      for (int kb = 0; kb < 4; ++kb)
        tma_load_2d_2sm(smem + kOffW4 + kb * kW4Bytes, &tmW4, w_bar, kb * 64, (int)cta_rank * (kCo / 2));
      int iter = 0, ys = 0;
      uint32_t yph = 0;
      for (int item = first; item < total_items; item += stride, ++iter) {
        int w0, h0, n0;
        coords(item, w0, h0, n0);
        const int ps = iter % kPatchStages;
        const uint32_t pph = (iter / kPatchStages) & 1;
        mbar_wait(&patch_empty[ps], pph ^ 1);
        if (cta_rank == 0) mbar_expect_tx(&patch_full[ps], 2 * kPatchBytes);
        tma_load_4d_2sm(smem + kOffPatch + ps * kPatchStride, &tmT, &patch_full[ps], 0, w0 - 1, h0 - 1, n0);
        for (int kb = 0; kb < 3; ++kb) {
          mbar_wait(&y_empty[ys], yph ^ 1);
          if (cta_rank == 0) mbar_expect_tx(&y_full[ys], 2 * kYBytes);
          tma_load_4d_2sm(smem + kOffY + ys * kYBytes, &tmY, &y_full[ys], kb * 64, w0, h0, n0);
          if (++ys == kYStages) {
            ys = 0;
            yph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ================= MMA issuer (leader): G1 and the first three K blocks of G2 of item i, then the last of item i - 1 =================
    constexpr uint32_t idesc1 = umma_idesc_bf16(256, kCh);
    constexpr uint32_t idesc2 = umma_idesc_bf16(256, kCo);
    auto issue_g2_tail = [&](int it) {
      const int g = it & 1;
      mbar_wait(&y3_ready[g], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_a = tmem_base + g * kCh;  // y3 over the first 32 columns of the G1 stage
      const uint32_t tmem_d = tmem_base + 2 * kCh + g * kCo;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts_2sm(tmem_d, tmem_a + 8 * k, umma_desc_sw128(smem_u32(smem + kOffW4 + 3 * kW4Bytes), 1024) + 2 * k,
                           idesc2, 1u);
        umma_commit_2sm(&acc2_full[g], 0b11);
      }
      __syncwarp();
    };
    mbar_wait(w_bar, 0);
    int iter = 0, ys = 0;
    uint32_t yph = 0;
    for (int item = first; item < total_items; item += stride, ++iter) {
      const int g = iter & 1;
      const int ps = iter % kPatchStages;
      const uint32_t pph = (iter / kPatchStages) & 1;
      // G1 stage g was last used by item iter - 2: its accumulator was read by E1 before y3_ready (waited on in
      // issue_g2_tail(iter - 2)) and its y3 by that MMA, issued before this point: the tensor pipe keeps the order
      mbar_wait(&patch_full[ps], pph);
      tc_fence_after();
      {
        const uint32_t tmem_d = tmem_base + g * kCh;
        const uint64_t a_base = umma_desc_sw128(smem_u32(smem + kOffPatch + ps * kPatchStride), 10 * 128);
        const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffWh), 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = a_base + static_cast<uint64_t>((((tap / 3) * 10 + (tap % 3)) * 128 + k * 32) >> 4);
              const uint64_t bd = b_base + static_cast<uint64_t>((tap * kTapBytes + k * 32) >> 4);
              umma_bf16_ss_2sm(tmem_d, ad, bd, idesc1, (tap | k) != 0 ? 1u : 0u);
            }
          umma_commit_2sm(&patch_empty[ps], 0b11);
          umma_commit_2sm(&acc_full[g], 0b11);
        }
        __syncwarp();
      }
      // G2, K blocks y0, y1, y2 into stage g, which E2 of item iter - 2 must have read
      mbar_wait(&acc2_empty[g], ((iter >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < 3; ++kb) {
        mbar_wait(&y_full[ys], yph);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + 2 * kCh + g * kCo;
        const uint64_t a_base = umma_desc_sw128(smem_u32(smem + kOffY + ys * kYBytes), 1024);
        const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffW4 + kb * kW4Bytes), 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_2sm(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc2, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&y_empty[ys], 0b11);
        }
        __syncwarp();
        if (++ys == kYStages) {
          ys = 0;
          yph ^= 1;
        }
      }
      if (iter >= 1) issue_g2_tail(iter - 1);
    }
    if (iter >= 1) issue_g2_tail(iter - 1);
  } else if (warp >= 4) {
    // ================= epilogue groups: group g owns the stages g of G1 and G2 and staging buffer g =================
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wi = row & 7, hi = row >> 3;  // this thread's pixel of the tile
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
      const bool valid = n0 * p.tiles_w * p.tiles_h < p.total_tiles;  // the pair's second tile may lie beyond the batch

      // ---------------- E1: y3 tile = SiLU(BN_h(acc) + y2) -> tensor memory (A operand of G2's last K block) ----------------
      // the residual row is requested before the accumulator is ready
      uint4 res[8];
      {
        const uint4* res_row = reinterpret_cast<const uint4*>(p.res + (long long)n0 * p.res_sn + (long long)(h0 + hi) * p.res_sh +
                                                              (long long)(w0 + wi) * p.res_sw + p.res_c_off);
#pragma unroll
        for (int v = 0; v < 8; ++v) res[v] = valid ? __ldg(res_row + v) : make_uint4(0, 0, 0, 0);
      }
      mbar_wait(&acc_full[group], ph);
      tc_fence_after();
      {
        uint32_t acc[64], packed[32];
        tmem_ld_32x32b_x32(t_row + group * kCh, acc);
        tmem_ld_32x32b_x32(t_row + group * kCh + 32, acc + 32);
        tmem_ld_wait();
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(res);
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
          const float4 sc = *reinterpret_cast<const float4*>(s_sc_h + e);
          const float4 sh = *reinterpret_cast<const float4*>(s_sh_h + e);
          // the stand-alone epilogue's order of operations (gemm_tcgen05.cu): affine, + residual / 2, SiLU on x / 2
          float v0 = fmaf(__uint_as_float(acc[e]), sc.x, sh.x);
          float v1 = fmaf(__uint_as_float(acc[e + 1]), sc.y, sh.y);
          float v2 = fmaf(__uint_as_float(acc[e + 2]), sc.z, sh.z);
          float v3 = fmaf(__uint_as_float(acc[e + 3]), sc.w, sh.w);
          v0 = fmaf(bf16_lo(rw[e >> 1]), 0.5f, v0);
          v1 = fmaf(bf16_hi(rw[e >> 1]), 0.5f, v1);
          v2 = fmaf(bf16_lo(rw[(e >> 1) + 1]), 0.5f, v2);
          v3 = fmaf(bf16_hi(rw[(e >> 1) + 1]), 0.5f, v3);
          packed[e >> 1] = pack_bf16x2(apply_act<ACT_SILU>(v0), apply_act<ACT_SILU>(v1));
          packed[(e >> 1) + 1] = pack_bf16x2(apply_act<ACT_SILU>(v2), apply_act<ACT_SILU>(v3));
        }
        // 64 channels -> columns 0 .. 31 of the stage: inside what this thread has consumed
        tmem_st_32x32b_x32(t_row + group * kCh, packed);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank == 0) mbar_arrive(&y3_ready[group]);
        else mbar_arrive_remote(&y3_ready[group], 0);
      }

      // ---------------- E2: output tile, one 64-channel chunk at a time -> staging -> TMA store ----------------
      mbar_wait(&acc2_full[group], ph);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint32_t acc[64], packed[32];
        tmem_ld_32x32b_x32(t_row + 2 * kCh + group * kCo + j * 64, acc);
        tmem_ld_32x32b_x32(t_row + 2 * kCh + group * kCo + j * 64 + 32, acc + 32);
        tmem_ld_wait();
        if (j == 1) {  // the stage has been read: G2 of the item after next may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (cta_rank == 0) mbar_arrive(&acc2_empty[group]);
            else mbar_arrive_remote(&acc2_empty[group], 0);
          }
        }
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
          const float4 sc = *reinterpret_cast<const float4*>(s_sc_4 + j * 64 + e);
          const float4 sh = *reinterpret_cast<const float4*>(s_sh_4 + j * 64 + e);
          const float v0 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e]), sc.x, sh.x));
          const float v1 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e + 1]), sc.y, sh.y));
          const float v2 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e + 2]), sc.z, sh.z));
          const float v3 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[e + 3]), sc.w, sh.w));
          packed[e >> 1] = pack_bf16x2(v0, v1);
          packed[(e >> 1) + 1] = pack_bf16x2(v2, v3);
        }
        if (gtid == 0) tma_store_wait_read<0>();  // the previous chunk has left the staging buffer
        bar_sync(bar_id, 128);
#pragma unroll
        for (int v = 0; v < 8; ++v)
          *reinterpret_cast<uint4*>(stage_out + row * 128 + ((static_cast<uint32_t>(v) ^ sw) << 4)) =
              make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
        fence_proxy_async_smem();
        bar_sync(bar_id, 128);
        if (gtid == 0) {
          tma_store_4d(&tmO, stage_out, j * 64, w0, h0, n0);
          tma_store_commit();
        }
      }
    }
    if (gtid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while its peer may still read its shared memory or arrive on its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace

bool gelan_tail_supported(int H, int W) { return H % kTH == 0 && W % kTW == 0 && H >= kTH && W >= kTW; }

// t: (B, H, W, 64) bf16, the first conv's output; g: (B, H, W, 256) bf16, the concatenation buffer with y0 | y1 | y2
// in channels 0 .. 191 (channels 192 .. 255 are neither read nor written); w_h [64][3][3][64], w4 [128][256] bf16;
// out: (B, H, W, 128) bf16.
int run_gelan_tail(const void* t, const void* g, int B, int H, int W, const void* w_h, const float* scale_h,
                   const float* shift_h, const void* w4, const float* scale_4, const float* shift_4, void* out,
                   int reverse, int num_sms, cudaStream_t stream) {
  if (!gelan_tail_supported(H, W)) {
    set_error("gelan_tail: the map %d x %d must tile into 16 x 8 blocks", H, W);
    return -1;
  }
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(gelan_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  CUtensorMap tt, ty, twh, tw4, to;
  {
    const uint64_t dims[4] = {(uint64_t)kCh, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t row = (uint64_t)kCh * 2;
    const uint64_t strides[3] = {row, row * W, row * W * H};
    const uint32_t box[4] = {64, 10, 18, 1};
    if (int r = make_tensor_map_bf16(&tt, t, 4, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {256, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t row = 256 * 2;
    const uint64_t strides[3] = {row, row * W, row * W * H};
    const uint32_t box[4] = {64, (uint32_t)kTW, (uint32_t)kTH, 1};
    if (int r = make_tensor_map_bf16(&ty, g, 4, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[2] = {576, (uint64_t)kCh};
    const uint64_t strides[1] = {576 * 2};
    const uint32_t box[2] = {64, (uint32_t)(kCh / 2)};
    if (int r = make_tensor_map_bf16(&twh, w_h, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[2] = {256, (uint64_t)kCo};
    const uint64_t strides[1] = {256 * 2};
    const uint32_t box[2] = {64, (uint32_t)(kCo / 2)};
    if (int r = make_tensor_map_bf16(&tw4, w4, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)kCo, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t row = (uint64_t)kCo * 2;
    const uint64_t strides[3] = {row, row * W, row * W * H};
    const uint32_t box[4] = {64, (uint32_t)kTW, (uint32_t)kTH, 1};
    if (int r = make_tensor_map_bf16(&to, out, 4, dims, strides, box)) return r;
  }
  TailParams p;
  p.tiles_w = W / kTW;
  p.tiles_h = H / kTH;
  p.total_tiles = p.tiles_w * p.tiles_h * B;
  p.reverse = reverse;
  p.res = static_cast<const __nv_bfloat16*>(g);
  p.res_sw = 256;
  p.res_sh = 256ll * W;
  p.res_sn = 256ll * W * H;
  p.res_c_off = 128;
  p.scale_h = scale_h;
  p.shift_h = shift_h;
  p.scale_4 = scale_4;
  p.shift_4 = shift_4;
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
  HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gelan_tail_kernel, tt, ty, twh, tw4, to, p));
  return 0;
}

}  // namespace hgr
