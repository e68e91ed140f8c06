// encoder.conv2 -> encoder.cspelan1.cv1 as ONE kernel
// (reference model/gelan.py:156 `conv2 = Conv(64, 128, 3, 2)`, :127 GELANBlock.cv1 = Conv(128, 128, 1, 1);
// Conv.forward :56 = SiLU(BN(conv(x)))):
//
//     a2 = SiLU(BN1(conv3x3_s2(a1)))        G1 (K = 9 taps x 64 ch), E1
//     g  = SiLU(BN2(conv1x1(a2)))           G2 (K = 128),            E2
//
// As two launches a2 (B x 48 x 48 x 128 bf16, 0.6 GB at batch 1024) is written to HBM and read straight back by
// the HBM-bound 1x1 layer.  Here the tile of a2 that an epilogue group has just activated goes to shared memory in
// the K-major SWIZZLE_128B layout, the tensor core multiplies it with the resident 128 x 128 weight block of cv1, and
// only g leaves the SM.  a2 is rounded to bf16 exactly where the separate launch stored it.
//
// The kernel is the CTA-pair (cta_group::2, M = 256) variant of the implicit-GEMM kernel of gemm_tcgen05.cu with a
// second, short MMA per tile: the leader's MMA warp issues G1 of tile i, then G2 of tile i - 1 (whose a2 tile was
// produced by the epilogue while G1 of tile i ran).  TMEM: two G1 accumulator stages (2 x 128 columns) and two G2
// accumulator stages (2 x 128 columns).  Each epilogue group owns one stage of each and one 32 KiB buffer that holds
// first its a2 tile (A operand of G2) and then, in place, the bf16 output tile for the TMA store.
#include <cstdio>
#include <cstring>

#include "epilogue_math.cuh"
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 384;
constexpr int kTileM = 128;
constexpr int kTileK = 64;
constexpr int kC = 128;                          // channels of a2 and of g
constexpr int kABytes = kTileM * kTileK * 2;     // 16 KiB: this CTA's 128 pixel rows of one k-step
constexpr int kBBytes = kC * kTileK * 2 / 2;     // 8 KiB: this CTA's 64 weight rows of one k-step
constexpr int kStages = 6;
constexpr int kChunkBytes = kTileM * 64 * 2;     // one 64-channel chunk of a 128-row bf16 tile
constexpr int kInterBytes = 2 * kChunkBytes;     // a2 tile / output staging of one epilogue group
constexpr int kW2Bytes = 2 * kBBytes;            // this CTA's 64 rows of cv1's weights, two k-blocks
constexpr int kOffA = 0;
constexpr int kOffB = kStages * kABytes;
constexpr int kOffInter = kOffB + kStages * kBBytes;
constexpr int kOffW2 = kOffInter + 2 * kInterBytes;
constexpr int kOffAffine = kOffW2 + kW2Bytes;    // scale1, shift1, scale2, shift2: 4 x 128 floats
constexpr int kOffBars = kOffAffine + 4 * kC * 4;
constexpr int kNumBars = 2 * kStages + 9;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kOffB % 1024 == 0 && kOffInter % 1024 == 0 && kOffW2 % 1024 == 0, "operand tiles need 1024-byte alignment");
static_assert(kSmemBytes <= 227 * 1024, "conv_chain shared-memory plan exceeds one CTA");

// tile index -> pixel-box origin (the same walk as gemm_tcgen05.cu's TileMap with one N tile)
struct PairTileMap {
  int tiles_w, tiles_h, bw, bh, bimg;
  int last;  // >= 0: walk the grid back to front
  int rank;  // this CTA takes M tile 2 * item + rank (it may lie beyond the map: TMA clips loads and stores)
  __device__ __forceinline__ void coords(int item, int& w0, int& h0, int& n0) const {
    if (last >= 0) item = last - item;
    int mt = item * 2 + rank;
    const int tw = mt % tiles_w;
    mt /= tiles_w;
    const int th = mt % tiles_h;
    const int tn = mt / tiles_h;
    w0 = tw * bw;
    h0 = th * bh;
    n0 = tn * bimg;
  }
};

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
      printf("hgr: conv_chain mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
      __trap();
    }
  }
}

// 64 accumulator columns of one pixel row -> SiLU(affine) -> bf16 -> one SWIZZLE_128B row of a chunk buffer
__device__ __forceinline__ void activate_store_chunk(const uint32_t (&acc)[64], const float* s_scale, const float* s_shift,
                                                     uint8_t* buf, int row, uint32_t sw) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float4* sc4 = reinterpret_cast<const float4*>(s_scale + half * 32);
    const float4* sh4 = reinterpret_cast<const float4*>(s_shift + half * 32);
    uint32_t packed[16];
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      const float4 sc = sc4[e >> 2], sh = sh4[e >> 2];
      const float v0 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[half * 32 + e]), sc.x, sh.x));
      const float v1 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[half * 32 + e + 1]), sc.y, sh.y));
      const float v2 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[half * 32 + e + 2]), sc.z, sh.z));
      const float v3 = apply_act<ACT_SILU>(fmaf(__uint_as_float(acc[half * 32 + e + 3]), sc.w, sh.w));
      packed[e >> 1] = pack_bf16x2(v0, v1);
      packed[(e >> 1) + 1] = pack_bf16x2(v2, v3);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint32_t chunk = static_cast<uint32_t>(half * 4 + v) ^ sw;
      *reinterpret_cast<uint4*>(buf + row * 128 + chunk * 16) =
          make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_chain_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO, const GemmParams p,
                  const float* __restrict__ scale2, const float* __restrict__ shift2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* full_bar = bars;                            // [kStages]
  uint64_t* empty_bar = bars + kStages;                 // [kStages]
  uint64_t* acc_full_bar = bars + 2 * kStages;          // [2] G1 of a tile has retired (both CTAs)
  uint64_t* acc_empty_bar = bars + 2 * kStages + 2;     // [2] leader's: both CTAs' readers are done with the stage
  uint64_t* inter_ready_bar = bars + 2 * kStages + 4;   // [2] leader's: both CTAs' a2 tiles are in shared memory
  uint64_t* acc2_full_bar = bars + 2 * kStages + 6;     // [2] G2 of a tile has retired (both CTAs)
  uint64_t* w2_bar = bars + 2 * kStages + 8;            // leader's: both halves of cv1's weights are resident
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  float* s_aff = reinterpret_cast<float*>(smem + kOffAffine);  // [scale1 | shift1 | scale2 | shift2], pre-halved

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmW2);
    prefetch_tensormap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4 * 2);    // one arrival per epilogue warp of the group, both CTAs
      mbar_init(&inter_ready_bar[i], 4 * 2);
      mbar_init(&acc2_full_bar[i], 1);
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
    s_aff[i] = 0.5f * (p.scale ? p.scale[i] : 1.0f);
    s_aff[kC + i] = 0.5f * (p.shift ? p.shift[i] : 0.0f);
    s_aff[2 * kC + i] = 0.5f * (scale2 ? scale2[i] : 1.0f);
    s_aff[3 * kC + i] = 0.5f * (shift2 ? shift2[i] : 0.0f);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t cta_rank = cluster_ctarank();
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_items = (tiles_m + 1) / 2;  // work items of the pair-wide walk
  const int ksteps = p.num_taps * p.chunks_per_tap;
  const int bw = 1 << p.bw_log2, bh = 1 << p.bh_log2;
  const int bimg = kTileM >> (p.bw_log2 + p.bh_log2);
  const int first = blockIdx.x / 2, stride = gridDim.x / 2;
  const PairTileMap tm{p.tiles_w, p.tiles_h, bw, bh, bimg, p.reverse ? total_items - 1 : -1, (int)cta_rank};

  if (warp == 0) {
    // ================= TMA producer (both CTAs): own A rows, own half of the weights =================
    if (elect_one_sync()) {
      if (cta_rank == 0) mbar_expect_tx(w2_bar, 2 * kW2Bytes);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
        tma_load_2d_2sm(smem + kOffW2 + kb * kBBytes, &tmW2, w2_bar, kb * kTileK, (int)cta_rank * (kC / 2));
      int stage = 0;
      uint32_t phase = 0;
      for (int item = first; item < total_items; item += stride) {
        int w0, h0, n0;
        tm.coords(item, w0, h0, n0);
        if (p.prefetch_dist > 0) {
          // The loop is bound by bytes in flight (stages x 24 KiB per CTA against ~2.5 us of loaded-HBM latency);
          // an L2 prefetch of a later item's input patch - the four space-to-depth parity planes, taps (1,1),
          // (1,2), (2,1), (2,2) - extends the window beyond what shared memory can hold.
          const int ahead = item + p.prefetch_dist * stride;
          if (ahead < total_items && p.num_taps == 9) {
            int pw0, ph0, pn0;
            tm.coords(ahead, pw0, ph0, pn0);
            const int taps[4] = {4, 5, 7, 8};
#pragma unroll
            for (int t = 0; t < 4; ++t)
              for (int chunk = 0; chunk < p.chunks_per_tap; ++chunk)
                tma_prefetch_5d(&tmA, p.a_c_off + p.tap_dc[taps[t]] + chunk * kTileK, pw0 + p.tap_dw[taps[t]],
                                p.tap_p[taps[t]], ph0 + p.tap_dh[taps[t]], pn0);
          }
        }
        int ks = 0;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int cw = w0 + p.tap_dw[tap];
          const int ch = h0 + p.tap_dh[tap];
          const int cp = p.tap_p[tap];
          const int cc = p.a_c_off + p.tap_dc[tap];
          for (int chunk = 0; chunk < p.chunks_per_tap; ++chunk, ++ks) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (kABytes + kBBytes));
            tma_load_5d_2sm(smem + kOffA + stage * kABytes, &tmA, &full_bar[stage], cc + chunk * kTileK, cw, cp, ch, n0);
            tma_load_2d_2sm(smem + kOffB + stage * kBBytes, &tmW, &full_bar[stage], ks * kTileK,
                            (int)cta_rank * (kC / 2));
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ================= MMA issuer (leader): G1 of item i, then G2 of item i - 1 =================
    constexpr uint32_t idesc = umma_idesc_bf16(2 * kTileM, kC);
    auto issue_g2 = [&](int it) {
      const int g = it & 1;
      mbar_wait_cluster(&inter_ready_bar[g], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + 2 * kC + g * kC;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t a_base = umma_desc_sw128(smem_u32(smem + kOffInter + g * kInterBytes + kb * kChunkBytes), 1024);
        const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffW2 + kb * kBBytes), 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k)
            umma_bf16_ss_2sm(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          if (kb == 1) umma_commit_2sm(&acc2_full_bar[g], 0b11);
        }
        __syncwarp();
      }
    };
    mbar_wait_cluster(w2_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int item = first; item < total_items; item += stride, ++iter) {
      const int acc = iter & 1;
      mbar_wait(&acc_empty_bar[acc], ((iter >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * kC;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_base = umma_desc_sw128(smem_u32(smem + kOffA + stage * kABytes), 1024);
        const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffB + stage * kBBytes), 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k)
            umma_bf16_ss_2sm(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 0b11);
          if (ks == ksteps - 1) umma_commit_2sm(&acc_full_bar[acc], 0b11);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (iter >= 1) issue_g2(iter - 1);
    }
    if (iter >= 1) issue_g2(iter - 1);
  } else if (warp >= 4) {
    // ================= epilogue groups: group g owns accumulator stages g of G1 and G2 and buffer g =================
    const int group = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = threadIdx.x - 128 - group * 128;
    const uint32_t bar_id = 1 + group;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint8_t* inter = smem + kOffInter + group * kInterBytes;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int iter = 0;
    for (int item = first; item < total_items; item += stride, ++iter) {
      if ((iter & 1) != group) continue;
      const uint32_t ph = (iter >> 1) & 1;
      int w0, h0, n0;
      tm.coords(item, w0, h0, n0);

      // ---------------- E1: a2 tile -> shared memory (A operand of G2) ----------------
      mbar_wait(&acc_full_bar[group], ph);
      tc_fence_after();
      if (gtid == 0) tma_store_wait_read<0>();  // the previous output tile of this group has left the buffer
      bar_sync(bar_id, 128);
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint32_t acc[64];
        tmem_ld_32x32b_x32(t_row + group * kC + j * 64, acc);
        tmem_ld_32x32b_x32(t_row + group * kC + j * 64 + 32, acc + 32);
        tmem_ld_wait();
        if (j == 1) {
          // the G1 stage goes back to the leader's MMA warp before the arithmetic
          // (one release-arrive per WARP: 128 cluster-scope arrives per tile showed up as membar stalls in ncu)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (cta_rank == 0) mbar_arrive(&acc_empty_bar[group]);
            else mbar_arrive_cluster(&acc_empty_bar[group], 0);
          }
        }
        activate_store_chunk(acc, s_aff + j * 64, s_aff + kC + j * 64, inter + j * kChunkBytes, row, sw);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta_rank == 0) mbar_arrive(&inter_ready_bar[group]);
        else mbar_arrive_cluster(&inter_ready_bar[group], 0);
      }

      // ---------------- E2: g tile -> the same buffer, in place -> TMA store ----------------
      mbar_wait(&acc2_full_bar[group], ph);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        uint32_t acc[64];
        tmem_ld_32x32b_x32(t_row + 2 * kC + group * kC + j * 64, acc);
        tmem_ld_32x32b_x32(t_row + 2 * kC + group * kC + j * 64 + 32, acc + 32);
        tmem_ld_wait();
        activate_store_chunk(acc, s_aff + 2 * kC + j * 64, s_aff + 3 * kC + j * 64, inter + j * kChunkBytes, row, sw);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      bar_sync(bar_id, 128);
      if (gtid == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_store_4d(&tmO, inter + j * kChunkBytes, p.out_c_off + j * 64, w0 + p.out_w_off, h0, n0);
        tma_store_commit();
      }
    }
    if (gtid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA leaves while its peer may still read its shared memory or arrive on its barriers
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace

int launch_conv_chain(const ConvChainOp& op, int num_sms, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const GemmParams& p = op.p;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int items = (tiles_m + 1) / 2;
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
  HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel, op.a, op.w, op.w2, op.o, op.p, op.scale2, op.shift2));
  return 0;
}

// first: a 3x3 conv producing 128 channels, built for the pair mode (its W map box holds 64 rows);
// second: the 1x1 128 -> 128 conv that consumes exactly that tensor (its output map and channel offset are used).
int build_conv_chain_op(ConvChainOp& op, const GemmOp& first, const GemmOp& second, const void* w2) {
  if (first.bn != kC || first.p.cout != kC || first.p.cluster != 2 || first.halo || first.p.res != nullptr ||
      first.p.act != ACT_SILU || second.p.cout != kC || second.p.num_taps != 1 || second.p.chunks_per_tap != kC / 64 ||
      second.p.act != ACT_SILU || second.p.res != nullptr || first.p.tiles_w != second.p.tiles_w ||
      first.p.tiles_h != second.p.tiles_h || first.p.tiles_n != second.p.tiles_n ||
      first.p.bw_log2 != second.p.bw_log2 || first.p.bh_log2 != second.p.bh_log2) {
    set_error("conv_chain: needs a pair-mode 128-channel SiLU conv followed by a 1x1 128 -> 128 SiLU conv on the same grid");
    return -1;
  }
  memset(&op, 0, sizeof(op));
  op.a = first.a;
  op.w = first.w;
  op.o = second.o;
  op.p = first.p;
  op.p.out_c_off = second.p.out_c_off;
  op.p.out_w_off = second.p.out_w_off;
  op.p.prefetch_dist = conv_chain_prefetch();
  op.scale2 = second.p.scale;
  op.shift2 = second.p.shift;
  {
    const uint64_t dims[2] = {(uint64_t)kC, (uint64_t)kC};
    const uint64_t strides[1] = {(uint64_t)kC * 2};
    const uint32_t box[2] = {64, (uint32_t)(kC / 2)};
    if (int r = make_tensor_map_bf16(&op.w2, w2, 2, dims, strides, box)) return r;
  }
  op.flops = first.flops + second.flops;
  // a2 is neither written nor read: A1 + W1 + W2 + OUT
  op.bytes = first.bytes + second.bytes - 2.0 * 2.0 * (double)first.p.W * first.p.H * first.p.NIMG * kC;
  return 0;
}

}  // namespace hgr
