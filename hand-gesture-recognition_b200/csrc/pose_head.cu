// Pose head: tokens 1.. -> (256, F, F) -> bilinear x4 (align_corners=True)
// -> ReLU -> 1x1 conv 256 -> J (+bias)   (reference model/transformer.py:118-127,
// 146-150; F.interpolate semantics = ATen upsample_bilinear2d: src = dst *
// (in-1)/(out-1), i0 = floor(src), i1 = i0 + (i0 < in-1), lambda1 = src - i0).
//
// The reference materialises the up-sampled (256, 4F, 4F) tensor in HBM
// (590 K elements per image at 192x192) and reads it back for the conv.  Here
// a CTA owns 4 output rows of one image: it pulls the (at most 3) token rows
// they depend on into shared memory with one burst of cp.async.  Phase A interpolates VERTICALLY in
// fp32 and leaves, per (row, token column, channel pair), the value a = v(x)
// and the horizontal slope d = v(x+1) - v(x) as bf16x2 in shared memory.
// Phase B builds the MMA A-fragments on the fly: one 128-bit shared load and
// two fma.rn.relu.bf16x2 (a + w*d, ReLU fused) give a pixel's four channels of
// a k-step, already in fragment layout - the up-sampled tensor never exists.
// The contraction (M = pixels, N = J padded to 24, K = 256) runs on mma.sync
// m16n8k16; at 0.6 % of the network's FLOPs and N = 21 it cannot fill a
// tcgen05 tile.  The MMA's k index is only a summation index, so inside each
// 16-channel block fragment slot k = 2t + e + 8*hf is bound to channel
// 4t + 2*hf + e (for A and W alike): a thread's four channels are then
// contiguous and every fragment is one conflict-free vector load.
// Heatmaps are written NCHW, the layout get_max_preds and the reference's
// callers expect.
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kDim = 256;
constexpr int kRows = 4;   // output rows per CTA
constexpr int kWarps = 6;
constexpr int kTokRows = 3;  // token rows a 4-row output band can touch: floor(s*oy0) .. floor(s*(oy0+3)) + 1
constexpr int kVPitch = kDim * 4 + 64;  // bytes per (row, token column): 128 channel pairs x (a, d) bf16x2; pitch = 64 (mod 128) so that
                                        // the two token columns a quarter-warp touches in phase B land in disjoint banks
constexpr int kWPitch = kDim + 16;      // bf16 elements per weight row (544 B: two conflict-free wavefronts per LDS.64)
constexpr int kJPad = 24;

template <typename TOut>
__device__ __forceinline__ void store_out(TOut* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// relu(a + w * d) on two bf16 lanes
__device__ __forceinline__ uint32_t fma_relu_bf16x2(uint32_t w, uint32_t d, uint32_t a) {
  uint32_t r;
  asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(d), "r"(a));
  return r;
}

template <typename TOut, int F>
__global__ void __launch_bounds__(kWarps * 32)
pose_head_kernel(const __nv_bfloat16* __restrict__ tokens, const __nv_bfloat16* __restrict__ w,
                 const float* __restrict__ bias, TOut* __restrict__ heat, int J) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* vbuf = smem_raw;                                                          // [kRows][F][kVPitch]
  __nv_bfloat16* sw = reinterpret_cast<__nv_bfloat16*>(vbuf + kRows * F * kVPitch);  // [kJPad][kWPitch], K permuted
  __nv_bfloat16* stok = sw + kJPad * kWPitch;                                        // [kTokRows][F][kDim]

  const int So = 4 * F;
  const int b = blockIdx.y;
  const int oy0 = blockIdx.x * kRows;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float scale = (float)(F - 1) / (float)(So - 1);
  pdl_launch_dependents();
  pdl_wait();

  // ---- the token rows this band interpolates between: one burst of cp.async, a single exposed latency ----
  const int ty0 = (int)(scale * (float)oy0);
  const __nv_bfloat16* tok = tokens + ((size_t)b * (F * F + 1) + 1) * kDim;  // skip the class token
  for (int i = tid; i < kTokRows * F * (kDim / 8); i += kWarps * 32) {
    const int row = i / (F * (kDim / 8));
    const int ty = ty0 + row < F ? ty0 + row : F - 1;
    const int rem = i % (F * (kDim / 8));
    cp_async_16(stok + (size_t)i * 8, tok + (size_t)ty * F * kDim + rem * 8, 16u);
  }
  cp_async_commit();

  // ---- weights -> smem in natural order (rows >= J are zero-filled) ----
  for (int i = tid; i < kJPad * (kDim / 8); i += kWarps * 32) {
    const int j = i / (kDim / 8), c8 = i % (kDim / 8);
    cp_async_16(sw + j * kWPitch + c8 * 8, j < J ? w + (size_t)j * kDim + c8 * 8 : w, j < J ? 16u : 0u);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // ---- phase A: vertical interpolation, value + horizontal slope per token column ----
  for (int i = tid; i < kRows * F * (kDim / 8); i += kWarps * 32) {
    const int c8 = i % (kDim / 8);
    const int x = (i / (kDim / 8)) % F;
    const int r = i / ((kDim / 8) * F);
    const int x1 = x + (x < F - 1 ? 1 : 0);
    const float sy = scale * (float)(oy0 + r);
    const int y0 = (int)sy;
    const int y1 = y0 + (y0 < F - 1 ? 1 : 0);
    const float l1 = sy - (float)y0, l0 = 1.0f - l1;
    const uint4 q00 = *reinterpret_cast<const uint4*>(stok + ((y0 - ty0) * F + x) * kDim + c8 * 8);
    const uint4 q10 = *reinterpret_cast<const uint4*>(stok + ((y1 - ty0) * F + x) * kDim + c8 * 8);
    const uint4 q01 = *reinterpret_cast<const uint4*>(stok + ((y0 - ty0) * F + x1) * kDim + c8 * 8);
    const uint4 q11 = *reinterpret_cast<const uint4*>(stok + ((y1 - ty0) * F + x1) * kDim + c8 * 8);
    const uint32_t* t00 = reinterpret_cast<const uint32_t*>(&q00);
    const uint32_t* t10 = reinterpret_cast<const uint32_t*>(&q10);
    const uint32_t* t01 = reinterpret_cast<const uint32_t*>(&q01);
    const uint32_t* t11 = reinterpret_cast<const uint32_t*>(&q11);
    uint32_t ad[8];  // (a, d) for the four channel pairs 4*c8 .. 4*c8+3
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a0 = l0 * bf16_lo(t00[k]) + l1 * bf16_lo(t10[k]);
      const float a1 = l0 * bf16_hi(t00[k]) + l1 * bf16_hi(t10[k]);
      const float n0 = l0 * bf16_lo(t01[k]) + l1 * bf16_lo(t11[k]);
      const float n1 = l0 * bf16_hi(t01[k]) + l1 * bf16_hi(t11[k]);
      ad[2 * k] = pack_bf16x2(a0, a1);
      ad[2 * k + 1] = pack_bf16x2(n0 - a0, n1 - a1);
    }
    // a lane owns 32 contiguous bytes; which 16-byte half goes first alternates so that the eight lanes of a
    // store phase cover all 32 banks exactly once
    uint4* dst = reinterpret_cast<uint4*>(vbuf + (r * F + x) * kVPitch + c8 * 32);
    const int hsel = (lane ^ (lane >> 2)) & 1;
    const uint4 lo = make_uint4(ad[0], ad[1], ad[2], ad[3]), hi = make_uint4(ad[4], ad[5], ad[6], ad[7]);
    dst[hsel] = hsel ? hi : lo;
    dst[hsel ^ 1] = hsel ? lo : hi;
  }
  __syncthreads();

  // ---- phase B: horizontal interpolation + ReLU straight into A fragments, MMA against the weights ----
  // Two 16-pixel m-tiles per warp and pass (tile u of the pass sits kWarps tiles further): they share every
  // weight-fragment load, which is 40 % of the shared-memory wavefronts of a single-tile loop.
  const int mt_per_row = So >> 4;
  const int mtiles = kRows * mt_per_row;
  for (int mt0 = warp; mt0 < mtiles; mt0 += 2 * kWarps) {
    const uint8_t* pa[2];
    const uint8_t* pb[2];
    uint32_t wa2[2], wb2[2];
    int rr[2], oxs[2];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int mt = mt0 + u * kWarps;
      on[u] = mt < mtiles;
      const int mtc = on[u] ? mt : mt0;
      rr[u] = mtc / mt_per_row;
      oxs[u] = (mtc % mt_per_row) << 4;
      const float sxa = scale * (float)(oxs[u] + g), sxb = scale * (float)(oxs[u] + g + 8);
      const int xa = (int)sxa, xb = (int)sxb;
      const float wa = sxa - (float)xa, wb = sxb - (float)xb;
      wa2[u] = pack_bf16x2(wa, wa);
      wb2[u] = pack_bf16x2(wb, wb);
      pa[u] = vbuf + (rr[u] * F + xa) * kVPitch + t * 16;
      pb[u] = vbuf + (rr[u] * F + xb) * kVPitch + t * 16;
    }

    float acc[2][3][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) acc[u][nt][0] = acc[u][nt][1] = acc[u][nt][2] = acc[u][nt][3] = 0.f;

#pragma unroll 4
    for (int ks = 0; ks < kDim / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        // (a, d) of channel pairs (2t, 2t+1) and (2t+8, 2t+9) for both pixels
        const uint4 va = *reinterpret_cast<const uint4*>(pa[u] + ks * 64);
        const uint4 vb = *reinterpret_cast<const uint4*>(pb[u] + ks * 64);
        a[u][0] = fma_relu_bf16x2(wa2[u], va.y, va.x);  // row g,   slots 2t, 2t+1   = channels 4t, 4t+1
        a[u][1] = fma_relu_bf16x2(wb2[u], vb.y, vb.x);  // row g+8
        a[u][2] = fma_relu_bf16x2(wa2[u], va.w, va.z);  // row g,   slots 2t+8, 2t+9 = channels 4t+2, 4t+3
        a[u][3] = fma_relu_bf16x2(wb2[u], vb.w, vb.z);  // row g+8
      }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const uint2 bw = *reinterpret_cast<const uint2*>(sw + (nt * 8 + g) * kWPitch + ks * 16 + 4 * t);
        mma_bf16_16816(acc[0][nt], a[0], bw.x, bw.y);
        mma_bf16_16816(acc[1][nt], a[1], bw.x, bw.y);
      }
    }
    // ---- NCHW store: for one joint, lanes g = 0..7 cover 8 consecutive pixels ----
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!on[u]) continue;
      const int oy = oy0 + rr[u];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          if (j < J) {
            const float bj = bias[j];
            TOut* dst = heat + (((size_t)b * J + j) * So + oy) * So + oxs[u] + g;
            store_out<TOut>(dst, acc[u][nt][e] + bj);
            store_out<TOut>(dst + 8, acc[u][nt][2 + e] + bj);
          }
        }
      }
    }
  }
}

template <typename TOut, int F>
int launch_pose_impl(const __nv_bfloat16* tokens, const __nv_bfloat16* w, const float* bias, TOut* heat, int B, int J,
                     cudaStream_t stream) {
  constexpr size_t smem = (size_t)kRows * F * kVPitch + (size_t)kJPad * kWPitch * 2 + (size_t)kTokRows * F * kDim * 2;
  static_assert(smem <= 227 * 1024, "feature side does not fit shared memory");
  HGR_CHECK_CUDA(cudaFuncSetAttribute(pose_head_kernel<TOut, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(4 * F / kRows, B);
  HGR_CHECK_CUDA(launch_pdl(pose_head_kernel<TOut, F>, grid, dim3(kWarps * 32), smem, stream, tokens, w, bias, heat, J));
  return 0;
}

template <typename TOut>
int launch_pose_f(const __nv_bfloat16* tokens, const __nv_bfloat16* w, const float* bias, TOut* heat, int B, int F,
                  int J, cudaStream_t stream) {
  // the feature side is a compile-time constant (no integer divisions in the staging loops): 64..512-pixel inputs
  switch (F) {
    case 4: return launch_pose_impl<TOut, 4>(tokens, w, bias, heat, B, J, stream);
    case 8: return launch_pose_impl<TOut, 8>(tokens, w, bias, heat, B, J, stream);
    case 12: return launch_pose_impl<TOut, 12>(tokens, w, bias, heat, B, J, stream);
    case 16: return launch_pose_impl<TOut, 16>(tokens, w, bias, heat, B, J, stream);
    case 20: return launch_pose_impl<TOut, 20>(tokens, w, bias, heat, B, J, stream);
    case 24: return launch_pose_impl<TOut, 24>(tokens, w, bias, heat, B, J, stream);
    case 28: return launch_pose_impl<TOut, 28>(tokens, w, bias, heat, B, J, stream);
    case 32: return launch_pose_impl<TOut, 32>(tokens, w, bias, heat, B, J, stream);
    default: set_error("pose_head: feature side %d not instantiated (multiples of 4 up to 32)", F); return -1;
  }
}

}  // namespace

int launch_pose_head(const __nv_bfloat16* tokens, const __nv_bfloat16* w, const float* bias, void* heatmaps,
                     int out_dtype, int B, int F, int J, cudaStream_t stream, float* preds, float* maxvals) {
  if (J > kJPad || J < 1) {
    set_error("pose_head: unsupported J=%d", J);
    return -1;
  }
  if (pose_head_tc_enabled() && pose_head_tc_supported(F, J))
    return launch_pose_head_tc(tokens, w, bias, heatmaps, out_dtype, preds, maxvals, B, F, J, device_sm_count(), stream);
  // mma.sync kernel: it always writes the heatmaps; the decode, if asked for, is a second launch over them
  if (heatmaps == nullptr) {
    set_error("pose_head: the keypoints-only mode needs the tcgen05 kernel (F <= 20, HGR_POSE_TC not 0)");
    return -1;
  }
  int rc;
  if (out_dtype == DT_F32)
    rc = launch_pose_f<float>(tokens, w, bias, static_cast<float*>(heatmaps), B, F, J, stream);
  else
    rc = launch_pose_f<__nv_bfloat16>(tokens, w, bias, static_cast<__nv_bfloat16*>(heatmaps), B, F, J, stream);
  if (rc == 0 && preds != nullptr)
    rc = launch_get_max_preds(heatmaps, out_dtype, (long long)B * J, 16 * F * F, 4 * F, preds, maxvals, stream);
  return rc;
}

}  // namespace hgr
