// Weight gradients on the 5th-gen tensor cores (the autograd backward of every Conv2d / Linear
// of the path except conv1, reference model/gelan.py:18-56, model/transformer.py:29-77):
//     dW[co][tap][ci] = sum over output pixels p of  G[p][co] * X[shift_tap(p)][ci]
// The reduction runs over PIXELS while both operands are stored channel-contiguous (NHWC), so for tcgen05.mma both
// are MN-major: A = G^T (M = 128 output channels), B = X^T (N = BN input channels), K = 64 pixels per pipeline
// stage.  TMA brings a 64-pixel box as [64 rows x 64 channels] SWIZZLE_128B tiles - two for A, BN / 64 for B, at the
// tap-shifted coordinates of the forward implicit GEMM (gemm_tcgen05.cu; out-of-bounds rows are the zero padding) -
// and a K = 16 MMA reads 16 pixel rows of every tile: 8-row groups 1024 B apart (stride byte offset), 64-channel
// tiles 8 KiB apart (leading byte offset), instruction-descriptor bits 15 / 16 = MN-major A / B.
//
// grid = (co tile x ci tile, tap, pixel chunk); a CTA accumulates its chunk's boxes in tensor memory (BN fp32
// columns) and writes the 128 x BN partial tile [chunk][tap][co][ci]; the fixed-order reduce of train_wgrad.cu sums
// the chunks and writes PyTorch's (Cout, Cin, kh, kw).  One CTA per SM (4 x 48 KiB stages at BN = 256), one wave.
//
// Why: the mma.sync kernel it replaces is at that path's ceiling on B200 (an HMMA.16816 holds a scheduler's tensor
// sub-pipe ~22 cycles: ~420 TFLOP/s per GPU; 0.84 ms of the 4.0 ms training step at batch 32).  A layer with 64
// output channels runs the same M = 128 MMA: the second A tile lies beyond the G map's channel extent, TMA fills it
// with zeros and its 64 accumulator rows are not written.  conv1 (3 input channels) stays on train_wgrad.cu.
#include <cstring>

#include "hgr_internal.h"
#include "ptx.cuh"
#include "train.h"

namespace hgr {

namespace {

constexpr int kThreads = 192;   // warp 0 TMA producer, warp 1 MMA issuer + tensor-memory owner, warps 2-5 epilogue
constexpr int kBoxPix = 64;     // pixels (K) per pipeline stage
constexpr int kTileBytes = kBoxPix * 128;  // [64 pixels][64 channels] bf16
constexpr int kStages = 4;

struct WgradTcParams {
  float* partial;  // [chunks][taps][Cout][Cin]
  int Cout, Cin, taps;
  int tiles_ci;               // Cin / BN
  int tiles_w, tiles_h, tiles_n, bw, bh, bi;  // 64-pixel boxes of the OUTPUT map
  int boxes_per_chunk, nbox;
  int tap_dc[9], tap_dw[9], tap_p[9], tap_dh[9];  // tap -> coordinate offsets of the X map (as gemm_tcgen05.cu)
};

// MN-major SWIZZLE_128B operand (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>, LayoutType::B128:
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 MN elements per 128-byte row, the next 64 MN elements
// `lbo_bytes` further, 8-row K groups `sbo_bytes` apart.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                const WgradTcParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int kNB = BN / 64;                        // B tiles per stage
  constexpr int kStageBytes = (2 + kNB) * kTileBytes;  // A: 2 tiles (128 output channels), B: kNB tiles
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* acc_full = empty + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int co0 = (blockIdx.x / p.tiles_ci) * 128;
  const int ci0 = (blockIdx.x % p.tiles_ci) * BN;
  const int tap = blockIdx.y;
  const int b0 = blockIdx.z * p.boxes_per_chunk;
  const int b1 = b0 + p.boxes_per_chunk < p.nbox ? b0 + p.boxes_per_chunk : p.nbox;
  const int nk = b1 > b0 ? b1 - b0 : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    prefetch_tensormap(&tmG);
    prefetch_tensormap(&tmX);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % kStages;
        if (i >= kStages) mbar_wait(&empty[s], ((i / kStages) - 1) & 1);
        int box = b0 + i;
        const int w0 = (box % p.tiles_w) * p.bw;
        box /= p.tiles_w;
        const int h0 = (box % p.tiles_h) * p.bh;
        const int n0 = (box / p.tiles_h) * p.bi;
        uint8_t* st = smem + s * kStageBytes;
        mbar_expect_tx(&full[s], kStageBytes);
        tma_load_4d(st, &tmG, &full[s], co0, w0, h0, n0);
        tma_load_4d(st + kTileBytes, &tmG, &full[s], co0 + 64, w0, h0, n0);
#pragma unroll
        for (int q = 0; q < kNB; ++q)
          tma_load_5d(st + (2 + q) * kTileBytes, &tmX, &full[s], ci0 + 64 * q + p.tap_dc[tap], w0 + p.tap_dw[tap],
                      p.tap_p[tap], h0 + p.tap_dh[tap], n0);
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = BN, A and B MN-major
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN) | (1u << 15) | (1u << 16);
    for (int i = 0; i < nk; ++i) {
      const int s = i % kStages;
      mbar_wait(&full[s], (i / kStages) & 1);
      tc_fence_after();
      const uint32_t a = smem_u32(smem + s * kStageBytes);
      const uint32_t b = a + 2 * kTileBytes;
#pragma unroll
      for (int k = 0; k < kBoxPix / 16; ++k)  // 16 pixel rows = 2048 B per MMA
        umma_bf16_ss_elect(tmem, desc_mn_sw128(a + k * 2048, kTileBytes, 1024),
                           desc_mn_sw128(b + k * 2048, kTileBytes, 1024), idesc, (i | k) != 0 ? 1u : 0u);
      umma_commit_elect(&empty[s]);
    }
    if (nk > 0) umma_commit_elect(acc_full);
  } else {
    // epilogue: warp w may touch tensor-memory lanes 32 (w % 4) ..; lane = output channel row of the tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* out = p.partial + (((size_t)blockIdx.z * p.taps + tap) * p.Cout + co0 + row) * p.Cin + ci0;
    if (nk > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      if (nk > 0) {
        tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0u;  // an empty chunk still owes its (zero) partial to the reduce
      }
      if (co0 + row < p.Cout) {
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<uint4*>(out + c + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, BN);
  }
}

// 64-pixel box (bw x bh x bi) of a W x H map
void pick_box64(int W, int H, int& bw, int& bh, int& bi) {
  bw = 16;
  while (bw > 1 && W % bw != 0) bw >>= 1;
  bh = kBoxPix / bw;
  while (bh > 1 && H % bh != 0) bh >>= 1;
  bi = kBoxPix / (bw * bh);
}

int pick_bn(int Cin) { return Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64); }

template <int BN>
int launch_impl(const CUtensorMap& tmG, const CUtensorMap& tmX, const WgradTcParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = kStages * (2 + BN / 64) * kTileBytes + (2 * kStages + 1) * 8 + 16;
  static bool configured = false;
  if (!configured) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  HGR_CHECK_CUDA(launch_pdl(wgrad_tc_kernel<BN>, grid, dim3(kThreads), smem, st, tmG, tmX, p));
  return 0;
}

}  // namespace

bool wgrad_tc_supported(int g_ctot, int x_ctot, int Cin, int Cout, int k, int s, int H, int W) {
  if (!wgrad_tc_enabled()) return false;
  if (Cout % 64 != 0 || Cin % 64 != 0 || g_ctot % 8 != 0 || x_ctot % 8 != 0) return false;
  if (!((k == 1 && s == 1) || (k == 3 && (s == 1 || s == 2)))) return false;
  if (s == 2 && (x_ctot != Cin || (H & 1) || (W & 1))) return false;  // the stride-2 view reads a whole buffer
  return true;
}

int wgrad_tc_chunks(int Cout, int Cin, int k, long long P) {
  const int taps = k * k;
  const long long tiles = (long long)((Cout + 127) / 128) * (Cin / pick_bn(Cin)) * taps;
  long long chunks = 148 / tiles;                   // one CTA per SM, one wave
  const long long cap = (P + 255) / 256;            // the mma.sync path's cap: the partial buffer is sized for it
  if (chunks > cap) chunks = cap;
  return chunks < 1 ? 1 : (int)chunks;
}

// g: [P][g_ctot] and x: [B][H][W][x_ctot] with their channel offsets applied (both 16-byte aligned)
int launch_wgrad_tc(const __nv_bfloat16* g, int g_ctot, const __nv_bfloat16* x, int x_ctot, int B, int H, int W, int Cin,
                    int Cout, int k, int s, float* partial, int* chunks_out, cudaStream_t st) {
  const int Ho = H / s, Wo = W / s;
  WgradTcParams p;
  memset(&p, 0, sizeof(p));
  pick_box64(Wo, Ho, p.bw, p.bh, p.bi);
  p.partial = partial;
  p.Cout = Cout;
  p.Cin = Cin;
  p.taps = k * k;
  const int BN = pick_bn(Cin);
  p.tiles_ci = Cin / BN;
  p.tiles_w = Wo / p.bw;
  p.tiles_h = Ho / p.bh;
  p.tiles_n = (B + p.bi - 1) / p.bi;
  p.nbox = p.tiles_w * p.tiles_h * p.tiles_n;
  const int chunks = wgrad_tc_chunks(Cout, Cin, k, (long long)B * Ho * Wo);
  p.boxes_per_chunk = (p.nbox + chunks - 1) / chunks;
  if (chunks_out) *chunks_out = chunks;
  CUtensorMap tmG, tmX;
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t row = (uint64_t)g_ctot * 2;
    const uint64_t strides[3] = {row, row * Wo, row * Wo * Ho};
    const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bi};
    if (int r = make_tensor_map_bf16(&tmG, g, 4, dims, strides, box)) return r;
  }
  if (s == 1) {
    const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, 1, (uint64_t)H, (uint64_t)B};
    const uint64_t row = (uint64_t)x_ctot * 2;
    const uint64_t strides[4] = {row, row * W, row * W, row * W * H};
    const uint32_t box[5] = {64, (uint32_t)p.bw, 1, (uint32_t)p.bh, (uint32_t)p.bi};
    if (int r = make_tensor_map_bf16(&tmX, x, 5, dims, strides, box)) return r;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const int t = kh * k + kw;
        p.tap_dw[t] = kw - k / 2;
        p.tap_dh[t] = kh - k / 2;
      }
  } else {
    // stride 2 as the space-to-depth view of the NHWC buffer used by the forward kernel (plan.cu build_conv_op):
    // (c2 = pw * C + c, W / 2, ph = 2, H / 2, n); input column 2 ow + kw - 1 is block ow + (kw == 0 ? -1 : 0) with
    // parity (kw == 1 ? 0 : 1)
    const uint64_t dims[5] = {(uint64_t)(2 * Cin), (uint64_t)(W / 2), 2, (uint64_t)(H / 2), (uint64_t)B};
    const uint64_t pix = (uint64_t)Cin * 2;
    const uint64_t strides[4] = {2 * pix, pix * W, 2 * pix * W, pix * W * H};
    const uint32_t box[5] = {64, (uint32_t)p.bw, 1, (uint32_t)p.bh, (uint32_t)p.bi};
    if (int r = make_tensor_map_bf16(&tmX, x, 5, dims, strides, box)) return r;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int t = kh * 3 + kw;
        p.tap_dc[t] = (kw == 1 ? 0 : 1) * Cin;
        p.tap_dw[t] = kw == 0 ? -1 : 0;
        p.tap_p[t] = kh == 1 ? 0 : 1;
        p.tap_dh[t] = kh == 0 ? -1 : 0;
      }
  }
  const dim3 grid(((Cout + 127) / 128) * p.tiles_ci, p.taps, chunks);
  if (BN == 256) return launch_impl<256>(tmG, tmX, p, grid, st);
  if (BN == 128) return launch_impl<128>(tmG, tmX, p, grid, st);
  return launch_impl<64>(tmG, tmX, p, grid, st);
}

}  // namespace hgr
