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

namespace hgr {

namespace {

constexpr int kThreads = 384;
constexpr int kTileM = 128;
constexpr int kTileK = 64;                       // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kTileM * kTileK * 2;     // 16 KiB
constexpr int kStageBufBytes = kTileM * 64 * 2;  // one 64-channel output chunk, 16 KiB
constexpr int kMaxCout = 1024;

template <int BN>
struct Cfg {
  static constexpr int kBBytes = BN * kTileK * 2;
  static constexpr int kStages = BN == 256 ? 3 : (BN == 128 ? 4 : 6);
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kStages * kABytes;
  static constexpr int kOffOut = kOffB + kStages * kBBytes;     // [2 groups][2 bufs][16 KiB]
  static constexpr int kOffScale = kOffOut + 4 * kStageBufBytes;  // scale[kMaxCout], shift[kMaxCout]
  static constexpr int kOffBars = kOffScale + 2 * kMaxCout * 4;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
  static constexpr int kSmemBytes = kOffTmemPtr + 16;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages: 128 / 256 / 512 columns
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
            const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* full_bar = bars;                              // [kStages]
  uint64_t* empty_bar = bars + C::kStages;                // [kStages]
  uint64_t* acc_full_bar = bars + 2 * C::kStages;         // [2]
  uint64_t* acc_empty_bar = bars + 2 * C::kStages + 2;    // [2]
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, C::kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.cout; i += kThreads) {
    s_scale[i] = p.scale ? p.scale[i] : 1.0f;
    s_shift[i] = p.shift ? p.shift[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = tiles_m * p.tiles_nout;
  const int ksteps = p.num_taps * p.chunks_per_tap;
  const int bw = 1 << p.bw_log2, bh = 1 << p.bh_log2;
  const int bimg = kTileM >> (p.bw_log2 + p.bh_log2);

  auto tile_coords = [&](int tile, int& w0, int& h0, int& n0, int& noff) {
    const int nt = tile % p.tiles_nout;
    int mt = tile / p.tiles_nout;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    const int tn = mt / p.tiles_h;
    w0 = tw * bw;
    h0 = th * bh;
    n0 = tn * bimg;
    noff = nt * BN;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int w0, h0, n0, noff;
        tile_coords(tile, w0, h0, n0, noff);
        int ks = 0;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int cw = w0 + p.tap_dw[tap];
          const int ch = h0 + p.tap_dh[tap];
          const int cp = p.tap_p[tap];
          const int cc = p.a_c_off + p.tap_dc[tap];
          for (int chunk = 0; chunk < p.chunks_per_tap; ++chunk, ++ks) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], kABytes + C::kBBytes);
            tma_load_5d(smem + C::kOffA + stage * kABytes, &tmA, &full_bar[stage], cc + chunk * kTileK, cw, cp, ch,
                        n0);
            tma_load_2d(smem + C::kOffB + stage * C::kBBytes, &tmW, &full_bar[stage], ks * kTileK, noff);
            if (++stage == C::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + C::kOffA + stage * kABytes);
          const uint32_t b_addr = smem_u32(smem + C::kOffB + stage * C::kBBytes);
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k) {
            const uint64_t adesc = umma_desc_sw128(a_addr + k * 32, 1024);
            const uint64_t bdesc = umma_desc_sw128(b_addr + k * 32, 1024);
            umma_bf16_ss(tmem_d, adesc, bdesc, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (ks == ksteps - 1) umma_commit(&acc_full_bar[acc]);
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
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;      // pixel row inside the tile == TMEM lane
    const int gtid = threadIdx.x - 128 - group * 128;
    const uint32_t bar_id = 1 + group;
    uint8_t* out_bufs = smem + C::kOffOut + group * 2 * kStageBufBytes;
    const int wi = row & (bw - 1);
    const int hi = (row >> p.bw_log2) & (bh - 1);
    const int ni = row >> (p.bw_log2 + p.bh_log2);
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint32_t store_count = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      if ((iter & 1) != group) continue;
      const uint32_t acc_phase = (iter >> 1) & 1;
      int w0, h0, n0, noff;
      tile_coords(tile, w0, h0, n0, noff);
      const bool valid = (w0 + wi < p.W) && (h0 + hi < p.H) && (n0 + ni < p.NIMG);
      const __nv_bfloat16* res_row = nullptr;
      if (p.res != nullptr && valid)
        res_row = p.res + (long long)(n0 + ni) * p.res_sn + (long long)(h0 + hi) * p.res_sh +
                  (long long)(w0 + wi) * p.res_sw + noff;

      mbar_wait(&acc_full_bar[group], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + group * BN;

#pragma unroll 1
      for (int j = 0; j < BN / 64; ++j, ++store_count) {
        uint8_t* buf = out_bufs + (store_count & 1) * kStageBufBytes;
        // the TMA store that last read this buffer (two stores ago) must be done
        if (gtid == 0) tma_store_wait_read<1>();
        bar_sync(bar_id, 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = j * 64 + half * 32;  // column inside the tile
          uint4 rv[4];
          if (res_row != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(res_row + c0);
#pragma unroll
            for (int v = 0; v < 4; ++v) rv[v] = __ldg(rp + v);
          }
          uint32_t acc[32];
          tmem_ld_32x32b_x32(t_row + c0, acc);
          tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float v0 = fmaf(__uint_as_float(acc[e]), s_scale[noff + c0 + e], s_shift[noff + c0 + e]);
            float v1 = fmaf(__uint_as_float(acc[e + 1]), s_scale[noff + c0 + e + 1], s_shift[noff + c0 + e + 1]);
            if (res_row != nullptr) {
              const uint32_t r = reinterpret_cast<const uint32_t*>(rv)[e >> 1];
              v0 += bf16_lo(r);
              v1 += bf16_hi(r);
            }
            if (p.act == ACT_SILU) {
              v0 = silu_f(v0);
              v1 = silu_f(v1);
            } else if (p.act == ACT_GELU) {
              v0 = gelu_erf_f(v0);
              v1 = gelu_erf_f(v1);
            }
            packed[e >> 1] = pack_bf16x2(v0, v1);
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
        if (j == BN / 64 - 1) {
          // every TMEM read of this accumulator stage is complete: hand it back
          tc_fence_before();
          mbar_arrive(&acc_empty_bar[group]);
        }
        fence_proxy_async_smem();
        bar_sync(bar_id, 128);
        if (gtid == 0) {
          tma_store_4d(&tmO, buf, p.out_c_off + noff + j * 64, w0 + p.out_w_off, h0, n0);
          tma_store_commit();
        }
      }
    }
    if (gtid == 0) tma_store_wait_all();
  }

  // ---- teardown ---------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BN>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                int num_sms, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const int total = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_nout;
  const int grid = total < num_sms ? total : num_sms;
  gemm_kernel<BN><<<grid, kThreads, C::kSmemBytes, stream>>>(tmA, tmW, tmO, p);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int gemm_smem_bytes(int bn) {
  return bn == 256 ? Cfg<256>::kSmemBytes : (bn == 128 ? Cfg<128>::kSmemBytes : Cfg<64>::kSmemBytes);
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
  switch (bn) {
    case 64: return launch_impl<64>(tmA, tmW, tmO, p, num_sms, stream);
    case 128: return launch_impl<128>(tmA, tmW, tmO, p, num_sms, stream);
    case 256: return launch_impl<256>(tmA, tmW, tmO, p, num_sms, stream);
    default: set_error("launch_gemm: bad bn %d", bn); return -1;
  }
}

}  // namespace hgr
