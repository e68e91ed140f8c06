// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores.
//
// Replaces, for every Conv(+BN+SiLU) of the GELAN backbone (reference
// model/gelan.py:18-56, 59-87, 124-142), the 1x1 `proj` (model/multitasknet.py:26)
// and every nn.Linear of the ViT (model/transformer.py:33-37, 65, 75), the
// cuDNN/cuBLAS launch plus the separate BatchNorm / activation / residual
// kernels of the reference.
//
// Shape of the kernel (one persistent CTA per SM, 384 threads):
//   warp 0      TMA producer   - one box load of A (128 pixels x 64 channels,
//                                tap-shifted NHWC coordinates, OOB = zero
//                                padding) and one of W (BN x 64) per k-step
//   warp 1      MMA issuer     - tcgen05.mma cta_group::1 kind::f16,
//                                M=128, N=BN, K=16, accumulators in TMEM
//   warp 2      TMEM allocator
//   warps 4-7   epilogue group 0  (TMEM accumulator stage 0)
//   warps 8-11  epilogue group 1  (TMEM accumulator stage 1)
// Tiles alternate between the two accumulator stages, so one group drains
// tile i (tcgen05.ld -> scale/shift -> +residual -> SiLU/GELU -> bf16 ->
// swizzled smem -> TMA store) while the tensor core already works on tile
// i+1 and the other group is still finishing tile i-1.
#include <cstdio>

#include "hgr_internal.h"
#include "ptx.cuh"
#include "epilogue_math.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 384;
constexpr int kTileM = 128;
constexpr int kTileK = 64;                       // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kTileM * kTileK * 2;     // 16 KiB
constexpr int kStageBufBytes = kTileM * 64 * 2;  // one 64-channel output chunk, 16 KiB
constexpr int kMaxCout = 1024;

// CL = 1: one CTA per tile (tcgen05.mma cta_group::1).  CL = 2: a CTA PAIR works on two neighbouring M tiles of
// the same N tile with cta_group::2 MMAs (M = 256): every CTA stages its own 128 rows of A and only HALF of the
// weight tile, so a pipeline stage is smaller and - what decides - shared-memory traffic per MMA cycle drops:
// with cta_group::1 a BN = 256 k-step reads 12 KiB and TMA writes 48 KiB per 512 MMA cycles (192 B/clk against
// the SM's 128 B/clk, measured as 70-77 % tensor-pipe utilisation at best, 50 % for BN = 128); the pair needs
// 128 B/clk (BN = 256).
template <int BN, int CL = 1>
struct Cfg {
  static constexpr int kBBytes = BN * kTileK * 2 / CL;  // weight bytes THIS CTA stages per k-step
  static constexpr int kStages = CL == 2 ? (BN == 256 ? 5 : 7) : (BN == 256 ? 3 : (BN == 128 ? 5 : 6));
  static constexpr int kOutBufs = (BN == 128 || CL == 2) ? 1 : 2;  // staging buffers per epilogue group
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kStages * kABytes;
  static constexpr int kOffOut = kOffB + kStages * kBBytes;     // [2 groups][kOutBufs][16 KiB]
  static constexpr int kOffScale = kOffOut + 2 * kOutBufs * kStageBufBytes;  // scale[kMaxCout], shift[kMaxCout]
  static constexpr int kOffBars = kOffScale + 2 * kMaxCout * 4;
  // accumulator stages in tensor memory: with two, a tile's MMAs cannot start before the epilogue of the tile two
  // back has pulled its last column; with four (BN <= 128: all 512 columns) every epilogue group alternates between
  // two stages of its own and the MMAs run a tile ahead (+0.4 % on the step, three same-box A/B pairs).  Moving
  // to_qkv from 256- to 128-wide tiles to get the four stages was slower (0.071 -> 0.079 ms): the extra operand
  // traffic of the narrower tile costs more than the overlap returns.
  static constexpr int kAccStages = BN <= 128 ? 4 : 2;
  static constexpr int kNumBars = 2 * kStages + 2 * kAccStages;
  static constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
  static constexpr int kSmemBytes = kOffTmemPtr + 16;
  static constexpr int kTmemCols = kAccStages * BN;  // 256 / 512 / 512 columns
  static_assert(kSmemBytes <= 227 * 1024, "shared-memory plan exceeds one CTA");
};

// tile index -> pixel-box origin and output-channel offset
struct TileMap {
  int tiles_w, tiles_h, tiles_nout, bw, bh, bimg, bn;
  int last;  // >= 0: walk the grid back to front (tile -> last - tile)
  int cl, rank;  // cluster size and this CTA's rank: a cluster walks (M-tile group, N-tile) work items together,
                 // rank r takes M tile cl * group + r (it may lie beyond the map: TMA clips loads and stores)
  __device__ __forceinline__ void coords(int tile, int& w0, int& h0, int& n0, int& noff) const {
    if (last >= 0) tile = last - tile;
    const int nt = tile % tiles_nout;
    int mt = (tile / tiles_nout) * cl + rank;
    const int tw = mt % tiles_w;
    mt /= tiles_w;
    const int th = mt % tiles_h;
    const int tn = mt / tiles_h;
    w0 = tw * bw;
    h0 = th * bh;
    n0 = tn * bimg;
    noff = nt * bn;
  }
};

// One epilogue group (4 warps, 128 threads) drains accumulator stage `group` of every other tile:
// tcgen05.ld -> scale/shift (+ residual) -> activation -> bf16 -> swizzled smem -> TMA store.
// NBUF = staging buffers per group (each one 64-column chunk, 16 KiB).
enum : int { ROW_NONE = 0, ROW_NORM_IN = 1, ROW_STATS_OUT = 2 };

template <int BN, int ACT, bool RES, int NBUF, int ROW = ROW_NONE, int NACC = 2>
__device__ __forceinline__ void epilogue_group(const GemmParams& p, const CUtensorMap* tmO, const TileMap& tm,
                                               uint8_t* out_bufs, const float* s_scale, const float* s_shift,
                                               uint64_t* acc_full_bar, uint64_t* acc_empty_bar, uint32_t tmem_base,
                                               int group, int total_tiles, int first, int stride,
                                               int pair_rank = -1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;         // TMEM lane quarter this warp may touch
  const int row = q * 32 + lane;  // pixel row inside the tile == TMEM lane
  const int gtid = threadIdx.x - 128 - group * 128;
  const uint32_t bar_id = 1 + group;
  const int wi = row & (tm.bw - 1);
  const int hi = (row >> p.bw_log2) & (tm.bh - 1);
  const int ni = row >> (p.bw_log2 + p.bh_log2);
  const uint32_t sw = static_cast<uint32_t>(row & 7);
  uint32_t store_count = 0;
  int iter = 0;
  for (int tile = first; tile < total_tiles; tile += stride, ++iter) {
    if ((iter & 1) != group) continue;
    const int acc = iter % NACC;  // the group's tiles alternate between its NACC / 2 accumulator stages
    const uint32_t acc_phase = (iter / NACC) & 1;
    int w0, h0, n0, noff;
    tm.coords(tile, w0, h0, n0, noff);
    const bool valid = (w0 + wi < p.W) && (h0 + hi < p.H) && (n0 + ni < p.NIMG);
    const uint4* res_row = nullptr;
    uint4 rnext[8];
    if constexpr (RES) {
      if (valid)
        res_row = reinterpret_cast<const uint4*>(p.res + (long long)(n0 + ni) * p.res_sn +
                                                 (long long)(h0 + hi) * p.res_sh + (long long)(w0 + wi) * p.res_sw + noff);
      // the residual of the first 64-column chunk is requested before the accumulator is even ready
#pragma unroll
      for (int v = 0; v < 8; ++v) rnext[v] = res_row ? __ldg(res_row + v) : make_uint4(0, 0, 0, 0);
    }

    // folded LayerNorm: per-row scalars of the A operand (consumer) / running sums of the output row (producer)
    float rstd = 1.0f, rmean = 0.0f, s1 = 0.0f, s2 = 0.0f;
    const long long stats_row = (long long)(n0 + ni) * p.st_sn + (w0 + wi) + p.out_w_off;
    if constexpr (ROW == ROW_NORM_IN) {
      if (valid) {
        const float2 st = __ldg(p.stats_in + stats_row);
        rstd = st.y;
        rmean = st.x * st.y;
      }
    }

    mbar_wait(&acc_full_bar[acc], acc_phase);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

#pragma unroll 1
    for (int j = 0; j < BN / 64; ++j, ++store_count) {
      uint8_t* buf = out_bufs + (NBUF == 1 ? 0 : (store_count & 1)) * kStageBufBytes;
      uint4 rcur[8];
      if constexpr (RES) {
#pragma unroll
        for (int v = 0; v < 8; ++v) rcur[v] = rnext[v];
        if (j + 1 < BN / 64) {  // next chunk's residual flies while this chunk is computed
#pragma unroll
          for (int v = 0; v < 8; ++v) rnext[v] = res_row ? __ldg(res_row + (j + 1) * 8 + v) : make_uint4(0, 0, 0, 0);
        }
      }
      // the TMA store that last read this buffer must be done
      if (gtid == 0) tma_store_wait_read<NBUF - 1>();
      bar_sync(bar_id, 128);
      // Pull the whole 64-column chunk out of TMEM first; on the last chunk the accumulator stage goes back to
      // the MMA warp BEFORE the arithmetic, so the tensor core restarts on this stage while the values are
      // still being activated and stored (for BN = 64 that is right after one TMEM round trip).
      uint32_t acc_all[64];
      tmem_ld_32x32b_x32(t_row + j * 64, acc_all);
      tmem_ld_32x32b_x32(t_row + j * 64 + 32, acc_all + 32);
      tmem_ld_wait();
      if (j == BN / 64 - 1) {
        tc_fence_before();
        // pair mode: the accumulator stage belongs to the leader's MMA thread, which waits for BOTH CTAs' readers.
        // One arrival per warp, and from the peer a plain (CTA-scope release) remote arrive: cluster-scope
        // release-arrives show up as membar stalls in the epilogue (~800 cycles each).
        if (p.warp_arrive) __syncwarp();
        if (!p.warp_arrive || lane == 0) {
          if (pair_rank <= 0) mbar_arrive(&acc_empty_bar[acc]);
          else mbar_arrive_remote(&acc_empty_bar[acc], 0);
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c0 = j * 64 + half * 32;  // column inside the tile
        const uint32_t* acc = acc_all + half * 32;
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale + noff + c0);
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift + noff + c0);
        uint32_t packed[16];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float4 sc = sc4[e >> 2], sh = sh4[e >> 2];
          float v0, v1, v2, v3;
          if constexpr (ROW == ROW_NORM_IN) {
            v0 = fmaf(__uint_as_float(acc[e]), rstd, fmaf(-rmean, sc.x, sh.x));
            v1 = fmaf(__uint_as_float(acc[e + 1]), rstd, fmaf(-rmean, sc.y, sh.y));
            v2 = fmaf(__uint_as_float(acc[e + 2]), rstd, fmaf(-rmean, sc.z, sh.z));
            v3 = fmaf(__uint_as_float(acc[e + 3]), rstd, fmaf(-rmean, sc.w, sh.w));
          } else {
            v0 = fmaf(__uint_as_float(acc[e]), sc.x, sh.x);
            v1 = fmaf(__uint_as_float(acc[e + 1]), sc.y, sh.y);
            v2 = fmaf(__uint_as_float(acc[e + 2]), sc.z, sh.z);
            v3 = fmaf(__uint_as_float(acc[e + 3]), sc.w, sh.w);
          }
          if constexpr (RES) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(rcur) + half * 16 + (e >> 1);
            constexpr float rs = ACT == ACT_SILU ? 0.5f : 1.0f;
            v0 = fmaf(bf16_lo(rw[0]), rs, v0);
            v1 = fmaf(bf16_hi(rw[0]), rs, v1);
            v2 = fmaf(bf16_lo(rw[1]), rs, v2);
            v3 = fmaf(bf16_hi(rw[1]), rs, v3);
          }
          v0 = apply_act<ACT>(v0);
          v1 = apply_act<ACT>(v1);
          v2 = apply_act<ACT>(v2);
          v3 = apply_act<ACT>(v3);
          if constexpr (ROW == ROW_STATS_OUT) {
            s1 += (v0 + v1) + (v2 + v3);
            s2 = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, s2))));
          }
          packed[e >> 1] = pack_bf16x2(v0, v1);
          packed[(e >> 1) + 1] = pack_bf16x2(v2, v3);
        }
        // row-major 128-byte rows, 16-byte chunks XOR-swizzled by (row % 8):
        // the layout CU_TENSOR_MAP_SWIZZLE_128B expects on the store side.
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint32_t chunk = static_cast<uint32_t>(half * 4 + v) ^ sw;
          *reinterpret_cast<uint4*>(buf + row * 128 + chunk * 16) =
              make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
        }
      }
      fence_proxy_async_smem();
      bar_sync(bar_id, 128);
      if (gtid == 0) {
        tma_store_4d(tmO, buf, p.out_c_off + noff + j * 64, w0 + p.out_w_off, h0, n0);
        tma_store_commit();
      }
    }
    if constexpr (ROW == ROW_STATS_OUT) {
      static_assert(ROW != ROW_STATS_OUT || BN == 256, "row statistics need the whole 256-wide row in one tile");
      if (valid) {
        const float mean = s1 * (1.0f / BN);
        const float var = fmaxf(fmaf(s2, 1.0f / BN, -mean * mean), 0.0f);
        p.stats_out[stats_row] = make_float2(mean, rsqrtf(var + 1e-5f));
      }
    }
  }
  if (gtid == 0) tma_store_wait_all();
}

// SiLU is evaluated on h = x/2, so the 1/2 is folded into the affine that the epilogue applies.
template <int ACT>
__device__ __forceinline__ void load_affine(const GemmParams& p, float* s_scale, float* s_shift) {
  const float pre = ACT == ACT_SILU ? 0.5f : 1.0f;
  for (int i = threadIdx.x; i < p.cout; i += kThreads) {
    s_scale[i] = pre * (p.scale ? p.scale[i] : 1.0f);
    s_shift[i] = pre * (p.shift ? p.shift[i] : 0.0f);
  }
}

// CL = 2: the kernel runs as clusters of two CTAs that work on the same N tile and k-steps of two neighbouring M
// tiles.  Each CTA fetches HALF of every weight tile and multicasts it into both shared memories, which halves the
// weight traffic from L2 - the bound of these layers is operand delivery (about 74 B/clk/SM), not the tensor pipe.
// A pipeline slot is released by BOTH MMA warps (multicast tcgen05.commit), because either producer writes both.
template <int BN, int ACT, bool RES, int ROW = ROW_NONE, int CL = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  using C = Cfg<BN, CL>;
  extern __shared__ __align__(1024) uint8_t smem[];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* full_bar = bars;                              // [kStages]
  uint64_t* empty_bar = bars + C::kStages;                // [kStages]
  uint64_t* acc_full_bar = bars + 2 * C::kStages;                       // [kAccStages]
  uint64_t* acc_empty_bar = bars + 2 * C::kStages + C::kAccStages;      // [kAccStages]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + C::kOffTmemPtr);
  float* s_scale = reinterpret_cast<float*>(smem + C::kOffScale);
  float* s_shift = s_scale + kMaxCout;

  // ---- one-time setup --------------------------------------------------
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < C::kAccStages; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      // pair mode: the readers of both CTAs release the leader's stage; one arrival per warp or per thread
      mbar_init(&acc_empty_bar[i], (p.warp_arrive ? 4 : 128) * CL);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CL == 2) {
      tmem_alloc_2sm(tmem_ptr_smem, C::kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_smem, C::kTmemCols);
      tmem_relinquish();
    }
  }
  load_affine<ACT>(p, s_scale, s_shift);
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();  // the peer's barriers exist before anything is sent to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above touched only parameters; from here on the previous kernel's output is read
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = ((tiles_m + CL - 1) / CL) * p.tiles_nout;  // work items of one cluster-wide walk
  const int ksteps = p.num_taps * p.chunks_per_tap;
  const int bw = 1 << p.bw_log2, bh = 1 << p.bh_log2;
  const int bimg = kTileM >> (p.bw_log2 + p.bh_log2);
  const int first = blockIdx.x / CL, stride = gridDim.x / CL;

  const TileMap tm{p.tiles_w, p.tiles_h, p.tiles_nout, bw, bh, bimg, BN, p.reverse ? total_tiles - 1 : -1, CL,
                   (int)cta_rank};

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first; tile < total_tiles; tile += stride) {
        int w0, h0, n0, noff;
        tm.coords(tile, w0, h0, n0, noff);
        if (p.prefetch_dist > 0) {
          // L2 prefetch of the A tile this CTA will need `prefetch_dist` tiles from now (centre tap, every channel
          // chunk): the activations were written by the previous kernel and mostly come from HBM on first touch
          const int ahead = tile + p.prefetch_dist * stride;
          if (ahead < total_tiles) {
            int pw0, ph0, pn0, pnoff;
            tm.coords(ahead, pw0, ph0, pn0, pnoff);
            const int ct = p.num_taps >> 1;
            for (int chunk = 0; chunk < p.chunks_per_tap; ++chunk)
              tma_prefetch_5d(&tmA, p.a_c_off + p.tap_dc[ct] + chunk * kTileK, pw0 + p.tap_dw[ct], p.tap_p[ct],
                              ph0 + p.tap_dh[ct], pn0);
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
            if constexpr (CL == 1) {
              mbar_expect_tx(&full_bar[stage], kABytes + C::kBBytes);
              tma_load_5d(smem + C::kOffA + stage * kABytes, &tmA, &full_bar[stage], cc + chunk * kTileK, cw, cp, ch,
                          n0);
              tma_load_2d(smem + C::kOffB + stage * C::kBBytes, &tmW, &full_bar[stage], ks * kTileK, noff);
            } else {
              // the leader's barrier counts the bytes both CTAs of the pair bring in
              if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], CL * (kABytes + C::kBBytes));
              tma_load_5d_2sm(smem + C::kOffA + stage * kABytes, &tmA, &full_bar[stage], cc + chunk * kTileK, cw, cp,
                              ch, n0);
              tma_load_2d_2sm(smem + C::kOffB + stage * C::kBBytes, &tmW, &full_bar[stage], ks * kTileK,
                              noff + (int)cta_rank * (BN / CL));
            }
            if (++stage == C::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && (CL == 1 || cta_rank == 0)) {
    // ================= MMA issuer (pair mode: the leader only) =================
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM * CL, BN);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = first; tile < total_tiles; tile += stride, ++iter) {
      const int acc = iter % C::kAccStages;
      const uint32_t acc_phase = (iter / C::kAccStages) & 1;
      mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        // the issue block is guarded by elect.sync (not `lane == 0`) so that ptxas keeps the
        // descriptors in uniform registers: ~4 instructions per MMA instead of a 15-instruction waterfall
        const uint64_t a_base = umma_desc_sw128(smem_u32(smem + C::kOffA + stage * kABytes), 1024);
        const uint64_t b_base = umma_desc_sw128(smem_u32(smem + C::kOffB + stage * C::kBBytes), 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k) {
            if constexpr (CL == 1) umma_bf16_ss(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
            else umma_bf16_ss_2sm(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          if constexpr (CL == 1) {
            umma_commit(&empty_bar[stage]);
            if (ks == ksteps - 1) umma_commit(&acc_full_bar[acc]);
          } else {
            umma_commit_2sm(&empty_bar[stage], 0b11);
            if (ks == ksteps - 1) umma_commit_2sm(&acc_full_bar[acc], 0b11);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue groups =================
    const int group = (warp - 4) >> 2;  // accumulator stage this group drains
    epilogue_group<BN, ACT, RES, C::kOutBufs, ROW, C::kAccStages>(p, &tmO, tm, smem + C::kOffOut + group * C::kOutBufs * kStageBufBytes,
                                                   s_scale, s_shift, acc_full_bar, acc_empty_bar, tmem_base, group,
                                                   total_tiles, first, stride, CL == 2 ? (int)cta_rank : -1);
  }

  // ---- teardown ---------------------------------------------------------
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  else __syncthreads();
  if (warp == 2) {
    if constexpr (CL == 2) tmem_dealloc_2sm(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// 3x3 stride-1 convolution, 64 -> 64 channels, with the input halo staged ONCE per tile.
//
// The generic kernel above re-loads the 128-pixel A tile for each of the 9 taps, which makes the
// 64-channel 3x3 convs of cspelan1 (gelan.py:73-74 at 48x48, 15.5 % of the FLOPs) L2-bandwidth
// bound.  Here a tile is 8 x 16 output pixels; its (8+2) x (16+2) input halo for 64 channels is
// ONE TMA box of 180 pixel rows (128 B each, SWIZZLE_128B, out-of-image rows zero-filled), and
// every tap is the same smem patch addressed through a shifted UMMA descriptor:
//   start = patch + ((kh * 10 + kw) * 128 B),  8-row groups 10 * 128 B apart (one image row),
// which works because the tensor core applies the 128B swizzle to absolute smem address bits
// (tools/umma_probe.cu: any 128 B-aligned start and any stride byte offset with base_offset = 0).
// All nine 64x64 weight tiles (72 KiB) stay resident in smem for the whole kernel, so the
// per-tile operand traffic drops from 216 KiB to 22.5 KiB.
template <int CL>
struct HaloCfg {
  static constexpr int kPatchRows = 10 * 18;
  static constexpr int kPatchBytes = kPatchRows * 128;     // 23040, what one TMA box delivers
  static constexpr int kPatchStride = 23 * 1024;           // stage pitch, keeps 1024-byte alignment
  static constexpr int kStages = CL == 2 ? 5 : 4;
  static constexpr int kTapBytes = 64 * 128 / CL;          // this CTA's rows of one tap's [64 x 64] weight tile
  static constexpr int kWBytes = 9 * kTapBytes;
  static constexpr int kOffW = 0;
  static constexpr int kOffA = kWBytes;
  static constexpr int kOffOut = kOffA + kStages * kPatchStride;  // [2 groups][1 buf][16 KiB]
  static constexpr int kOffScale = kOffOut + 2 * kStageBufBytes;
  static constexpr int kOffBars = kOffScale + 2 * kMaxCout * 4;
  static constexpr int kNumBars = 2 * kStages + 5;
  static constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
  static constexpr int kSmemBytes = kOffTmemPtr + 16;
  static constexpr int kTmemCols = 128;
  static_assert(kOffA % 1024 == 0 && kSmemBytes <= 227 * 1024, "halo shared-memory plan");
};

// CL = 2: CTA-pair mode, same protocol as gemm_kernel<..., 2>: two neighbouring 8 x 16 tiles form one M = 256 MMA,
// every CTA stages its own halo patch and keeps only HALF of each tap's weight rows resident.  A 64-channel MMA is
// bound by shared-memory reads (4 KiB of A + 2 KiB of B per 32 MMA cycles = 192 B/clk); the pair reads 4 + 1 KiB.
template <int ACT, bool RES, int CL>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  using C = HaloCfg<CL>;
  constexpr int BN = 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* acc_full_bar = bars + 2 * C::kStages;
  uint64_t* acc_empty_bar = bars + 2 * C::kStages + 2;
  uint64_t* w_bar = bars + 2 * C::kStages + 4;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + C::kOffTmemPtr);
  float* s_scale = reinterpret_cast<float*>(smem + C::kOffScale);
  float* s_shift = s_scale + kMaxCout;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], (p.warp_arrive ? 4 : 128) * CL);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CL == 2) {
      tmem_alloc_2sm(tmem_ptr_smem, C::kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_smem, C::kTmemCols);
      tmem_relinquish();
    }
  }
  load_affine<ACT>(p, s_scale, s_shift);
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // everything above touched only parameters; from here on the previous kernel's output is read
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = (tiles_m + CL - 1) / CL;  // work items of one cluster-wide walk
  const int first = blockIdx.x / CL, stride = gridDim.x / CL;
  const TileMap tm{p.tiles_w, p.tiles_h, 1, 8, 16, 1, BN, p.reverse ? total_tiles - 1 : -1, CL, (int)cta_rank};

  if (warp == 0) {
    if (elect_one_sync()) {
      // resident weights: tap t is rows [0, 64) x K columns [64 t, 64 t + 64); a pair CTA keeps rows [32 rank, +32)
      if constexpr (CL == 1) {
        mbar_expect_tx(w_bar, C::kWBytes);
        for (int t = 0; t < 9; ++t) tma_load_2d(smem + C::kOffW + t * C::kTapBytes, &tmW, w_bar, t * 64, 0);
      } else {
        if (cta_rank == 0) mbar_expect_tx(w_bar, CL * C::kWBytes);
        for (int t = 0; t < 9; ++t)
          tma_load_2d_2sm(smem + C::kOffW + t * C::kTapBytes, &tmW, w_bar, t * 64, (int)cta_rank * (BN / CL));
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first; tile < total_tiles; tile += stride) {
        int w0, h0, n0, noff;
        tm.coords(tile, w0, h0, n0, noff);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if constexpr (CL == 1) {
          mbar_expect_tx(&full_bar[stage], C::kPatchBytes);
          tma_load_5d(smem + C::kOffA + stage * C::kPatchStride, &tmA, &full_bar[stage], p.a_c_off, w0 - 1, 0, h0 - 1,
                      n0);
        } else {
          if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], CL * C::kPatchBytes);
          tma_load_5d_2sm(smem + C::kOffA + stage * C::kPatchStride, &tmA, &full_bar[stage], p.a_c_off, w0 - 1, 0,
                          h0 - 1, n0);
        }
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && (CL == 1 || cta_rank == 0)) {
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM * CL, BN);
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = first; tile < total_tiles; tile += stride, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      {
        const uint32_t tmem_d = tmem_base + acc * BN;
        const uint32_t patch = smem_u32(smem + C::kOffA + stage * C::kPatchStride);
        const uint32_t wres = smem_u32(smem + C::kOffW);
        // descriptors differ only in the 14-bit start-address field: build the two constant parts once
        const uint64_t a_base = umma_desc_sw128(patch, 10 * 128);
        const uint64_t b_base = umma_desc_sw128(wres, 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = a_base + static_cast<uint64_t>((((tap / 3) * 10 + (tap % 3)) * 128 + k * 32) >> 4);
              const uint64_t bd = b_base + static_cast<uint64_t>((tap * C::kTapBytes + k * 32) >> 4);
              if constexpr (CL == 1) umma_bf16_ss(tmem_d, ad, bd, idesc, (tap | k) != 0 ? 1u : 0u);
              else umma_bf16_ss_2sm(tmem_d, ad, bd, idesc, (tap | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (CL == 1) {
            umma_commit(&empty_bar[stage]);
            umma_commit(&acc_full_bar[acc]);
          } else {
            umma_commit_2sm(&empty_bar[stage], 0b11);
            umma_commit_2sm(&acc_full_bar[acc], 0b11);
          }
        }
      }
      __syncwarp();
      if (++stage == C::kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int group = (warp - 4) >> 2;
    epilogue_group<BN, ACT, RES, 1>(p, &tmO, tm, smem + C::kOffOut + group * kStageBufBytes, s_scale, s_shift,
                                    acc_full_bar, acc_empty_bar, tmem_base, group, total_tiles, first, stride,
                                    CL == 2 ? (int)cta_rank : -1);
  }

  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    if constexpr (CL == 2) tmem_dealloc_2sm(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int ACT, bool RES, int CL>
int launch_halo_cl(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                   int num_sms, cudaStream_t stream) {
  using C = HaloCfg<CL>;
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<ACT, RES, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::kSmemBytes));
    configured = true;
  }
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int items = (tiles_m + CL - 1) / CL;
  int grid = items * CL < num_sms ? items * CL : num_sms;
  grid -= grid % CL;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<ACT, RES, CL>, tmA, tmW, tmO, p));
  return 0;
}

template <int ACT, bool RES>
int launch_halo_impl(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                     int num_sms, cudaStream_t stream) {
  if (p.cluster == 2) return launch_halo_cl<ACT, RES, 2>(tmA, tmW, tmO, p, num_sms, stream);
  return launch_halo_cl<ACT, RES, 1>(tmA, tmW, tmO, p, num_sms, stream);
}

template <int BN, int ACT, bool RES, int ROW, int CL>
int launch_impl_cl(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                   int num_sms, cudaStream_t stream) {
  using C = Cfg<BN, CL>;
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, ACT, RES, ROW, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        C::kSmemBytes));
    configured = true;
  }
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int items = ((tiles_m + CL - 1) / CL) * p.tiles_nout;
  int grid = items * CL < num_sms ? items * CL : num_sms;
  grid -= grid % CL;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  HGR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, ACT, RES, ROW, CL>, tmA, tmW, tmO, p));
  return 0;
}

template <int BN, int ACT, bool RES, int ROW = ROW_NONE>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                int num_sms, cudaStream_t stream) {
  if constexpr (BN >= 128) {
    if (p.cluster == 2) return launch_impl_cl<BN, ACT, RES, ROW, 2>(tmA, tmW, tmO, p, num_sms, stream);
  }
  return launch_impl_cl<BN, ACT, RES, ROW, 1>(tmA, tmW, tmO, p, num_sms, stream);
}

template <int BN>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p, int num_sms,
              cudaStream_t stream) {
  const bool res = p.res != nullptr;
  switch (p.act) {
    case ACT_NONE:
      return res ? launch_impl<BN, ACT_NONE, true>(tmA, tmW, tmO, p, num_sms, stream)
                 : launch_impl<BN, ACT_NONE, false>(tmA, tmW, tmO, p, num_sms, stream);
    case ACT_SILU:
      return res ? launch_impl<BN, ACT_SILU, true>(tmA, tmW, tmO, p, num_sms, stream)
                 : launch_impl<BN, ACT_SILU, false>(tmA, tmW, tmO, p, num_sms, stream);
    case ACT_GELU:
      return res ? launch_impl<BN, ACT_GELU, true>(tmA, tmW, tmO, p, num_sms, stream)
                 : launch_impl<BN, ACT_GELU, false>(tmA, tmW, tmO, p, num_sms, stream);
    default: set_error("launch_gemm: bad activation %d", p.act); return -1;
  }
}

}  // namespace

int gemm_smem_bytes(int bn) {
  return bn == 256 ? Cfg<256>::kSmemBytes : (bn == 128 ? Cfg<128>::kSmemBytes : Cfg<64>::kSmemBytes);
}

int launch_conv3x3_halo(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                        int num_sms, cudaStream_t stream) {
  if (p.cout != 64 || p.num_taps != 9 || p.chunks_per_tap != 1 || p.bw_log2 != 3 || p.bh_log2 != 4) {
    set_error("launch_conv3x3_halo: needs a 64->64 3x3 layer tiled 8x16");
    return -1;
  }
  const bool res = p.res != nullptr;
  switch (p.act) {
    case ACT_NONE:
      return res ? launch_halo_impl<ACT_NONE, true>(tmA, tmW, tmO, p, num_sms, stream)
                 : launch_halo_impl<ACT_NONE, false>(tmA, tmW, tmO, p, num_sms, stream);
    case ACT_SILU:
      return res ? launch_halo_impl<ACT_SILU, true>(tmA, tmW, tmO, p, num_sms, stream)
                 : launch_halo_impl<ACT_SILU, false>(tmA, tmW, tmO, p, num_sms, stream);
    default: set_error("launch_conv3x3_halo: unsupported activation %d", p.act); return -1;
  }
}

int launch_gemm(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                int num_sms, cudaStream_t stream) {
  if (p.cout > kMaxCout || p.cout % bn != 0) {
    set_error("launch_gemm: cout=%d not supported with bn=%d", p.cout, bn);
    return -1;
  }
  if (p.tiles_nout * bn != p.cout) {
    set_error("launch_gemm: tiles_nout=%d * bn=%d != cout=%d", p.tiles_nout, bn, p.cout);
    return -1;
  }
  if (p.stats_in != nullptr || p.stats_out != nullptr) {
    // LayerNorm-folded variants exist for the ViT's 256-wide token rows only
    if (bn != 256 || p.cout % 256 != 0 || (p.stats_in && p.stats_out)) {
      set_error("launch_gemm: row statistics need bn == 256 (cout %d) and only one of stats_in / stats_out", p.cout);
      return -1;
    }
    if (p.stats_out != nullptr) {
      if (p.cout != 256 || p.act != ACT_NONE || p.res == nullptr) {
        set_error("launch_gemm: stats_out is built for the residual-stream producers (cout 256, no activation, residual)");
        return -1;
      }
      return launch_impl<256, ACT_NONE, true, ROW_STATS_OUT>(tmA, tmW, tmO, p, num_sms, stream);
    }
    if (p.res != nullptr || (p.act != ACT_NONE && p.act != ACT_GELU)) {
      set_error("launch_gemm: stats_in is built for to_qkv (no activation) and net.1 (GELU), without residual");
      return -1;
    }
    return p.act == ACT_GELU ? launch_impl<256, ACT_GELU, false, ROW_NORM_IN>(tmA, tmW, tmO, p, num_sms, stream)
                             : launch_impl<256, ACT_NONE, false, ROW_NORM_IN>(tmA, tmW, tmO, p, num_sms, stream);
  }
  switch (bn) {
    case 64: return launch_bn<64>(tmA, tmW, tmO, p, num_sms, stream);
    case 128: return launch_bn<128>(tmA, tmW, tmO, p, num_sms, stream);
    case 256: return launch_bn<256>(tmA, tmW, tmO, p, num_sms, stream);
    default: set_error("launch_gemm: bad bn %d", bn); return -1;
  }
}

}  // namespace hgr
