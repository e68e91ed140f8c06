// Pose head: tokens 1.. -> (256, F, F) -> bilinear x4 (align_corners=True)
// -> ReLU -> 1x1 conv 256 -> J (+bias)   (reference model/transformer.py:118-127,
// 146-150; F.interpolate semantics = ATen upsample_bilinear2d: src = dst *
// (in-1)/(out-1), i0 = floor(src), i1 = i0 + (i0 < in-1), lambda1 = src - i0).
//
// The reference materialises the up-sampled (256, 4F, 4F) tensor in HBM
// (590 K elements per image at 192x192) and reads it back for the conv.  Here
// a CTA owns 8 output rows of one image: it interpolates vertically once into
// shared memory (fp32), and each warp then builds its MMA A-fragments on the
// fly by interpolating horizontally, applying ReLU and rounding to bf16 - the
// up-sampled tensor never exists.  The contraction (M = pixels, N = J padded
// to 24, K = 256) runs on mma.sync m16n8k16; at 0.6 % of the network's FLOPs
// and N = 21 it cannot fill a tcgen05 tile.  Heatmaps are written NCHW, the
// layout get_max_preds and the reference's callers expect.
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kDim = 256;
constexpr int kRows = 8;        // output rows per CTA
constexpr int kWarps = 8;
constexpr int kVPitch = kDim + 8;   // fp32 words per (row, x) line of the vertical-interp buffer
constexpr int kWPitch = kDim + 8;   // bf16 elements per weight row
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

template <typename TOut>
__global__ void __launch_bounds__(kWarps * 32)
pose_head_kernel(const __nv_bfloat16* __restrict__ tokens, const __nv_bfloat16* __restrict__ w,
                 const float* __restrict__ bias, TOut* __restrict__ heat, int F, int J) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* vbuf = reinterpret_cast<float*>(smem_raw);                                   // [kRows][F][kVPitch]
  __nv_bfloat16* sw = reinterpret_cast<__nv_bfloat16*>(vbuf + kRows * F * kVPitch);   // [kJPad][kWPitch]

  const int So = 4 * F;
  const int b = blockIdx.y;
  const int oy0 = blockIdx.x * kRows;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float scale = (float)(F - 1) / (float)(So - 1);

  // ---- weights -> smem (rows >= J are zero) ------------------------------
  for (int i = tid; i < kJPad * (kDim / 8); i += kWarps * 32) {
    const int j = i / (kDim / 8), c8 = i % (kDim / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < J) v = __ldg(reinterpret_cast<const uint4*>(w + (size_t)j * kDim) + c8);
    *reinterpret_cast<uint4*>(sw + j * kWPitch + c8 * 8) = v;
  }

  // ---- vertical interpolation of the 8 output rows -----------------------
  const __nv_bfloat16* tok = tokens + ((size_t)b * (F * F + 1) + 1) * kDim;  // skip the class token
  const int nitems = kRows * F * (kDim / 8);
  for (int i0 = tid; i0 < nitems; i0 += 4 * kWarps * 32) {
    // eight 16-byte requests in flight per thread before the first one is consumed
    uint4 ua[4], ub[4];
    float l1v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kWarps * 32;
      const int c8 = i % (kDim / 8);
      const int x = (i / (kDim / 8)) % F;
      const int r = i / ((kDim / 8) * F);
      const float sy = scale * (float)(oy0 + r);
      const int y0 = (int)sy;
      const int y1 = y0 + (y0 < F - 1 ? 1 : 0);
      l1v[u] = sy - (float)y0;
      ua[u] = ub[u] = make_uint4(0, 0, 0, 0);
      if (i < nitems) {
        ua[u] = __ldg(reinterpret_cast<const uint4*>(tok + (size_t)(y0 * F + x) * kDim) + c8);
        ub[u] = __ldg(reinterpret_cast<const uint4*>(tok + (size_t)(y1 * F + x) * kDim) + c8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kWarps * 32;
      if (i >= nitems) break;
      const int c8 = i % (kDim / 8);
      const int rx = i / (kDim / 8);  // r * F + x
      const float l1 = l1v[u], l0 = 1.0f - l1;
      const uint32_t a[4] = {ua[u].x, ua[u].y, ua[u].z, ua[u].w};
      const uint32_t bb[4] = {ub[u].x, ub[u].y, ub[u].z, ub[u].w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[2 * k] = l0 * bf16_lo(a[k]) + l1 * bf16_lo(bb[k]);
        o[2 * k + 1] = l0 * bf16_hi(a[k]) + l1 * bf16_hi(bb[k]);
      }
      float4* dst = reinterpret_cast<float4*>(vbuf + rx * kVPitch + c8 * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  __syncthreads();

  const int mt_per_row = So >> 4;
  const int mtiles = kRows * mt_per_row;
  for (int mt = warp; mt < mtiles; mt += kWarps) {
    const int r = mt / mt_per_row;
    const int ox0 = (mt % mt_per_row) << 4;
    // horizontal source positions of this thread's two pixels
    const float sx0 = scale * (float)(ox0 + g), sx1 = scale * (float)(ox0 + g + 8);
    const int xa0 = (int)sx0, xb0 = (int)sx1;
    const int xa1 = xa0 + (xa0 < F - 1 ? 1 : 0), xb1 = xb0 + (xb0 < F - 1 ? 1 : 0);
    const float wa1 = sx0 - (float)xa0, wa0 = 1.0f - wa1;
    const float wb1 = sx1 - (float)xb0, wb0 = 1.0f - wb1;
    const float* va0 = vbuf + (r * F + xa0) * kVPitch + 2 * t;
    const float* va1 = vbuf + (r * F + xa1) * kVPitch + 2 * t;
    const float* vb0 = vbuf + (r * F + xb0) * kVPitch + 2 * t;
    const float* vb1 = vbuf + (r * F + xb1) * kVPitch + 2 * t;

    float acc[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;

#pragma unroll 4
    for (int ks = 0; ks < kDim / 16; ++ks) {
      uint32_t a[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // channel pairs (2t,2t+1) and (2t+8,2t+9) of this k-step
        const int c = ks * 16 + hf * 8;
        const float2 p00 = *reinterpret_cast<const float2*>(va0 + c);
        const float2 p01 = *reinterpret_cast<const float2*>(va1 + c);
        const float2 p10 = *reinterpret_cast<const float2*>(vb0 + c);
        const float2 p11 = *reinterpret_cast<const float2*>(vb1 + c);
        a[2 * hf] = pack_bf16x2(fmaxf(wa0 * p00.x + wa1 * p01.x, 0.f), fmaxf(wa0 * p00.y + wa1 * p01.y, 0.f));
        a[2 * hf + 1] = pack_bf16x2(fmaxf(wb0 * p10.x + wb1 * p11.x, 0.f), fmaxf(wb0 * p10.y + wb1 * p11.y, 0.f));
      }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const __nv_bfloat16* wr = sw + (nt * 8 + g) * kWPitch + ks * 16 + 2 * t;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
        mma_bf16_16816(acc[nt], a, b0, b1);
      }
    }
    // ---- NCHW store: for one joint, lanes g = 0..7 cover 8 consecutive pixels ----
    const int oy = oy0 + r;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * t + e;
        if (j < J) {
          const float bj = bias[j];
          TOut* dst = heat + (((size_t)b * J + j) * So + oy) * So + ox0 + g;
          store_out<TOut>(dst, acc[nt][e] + bj);
          store_out<TOut>(dst + 8, acc[nt][2 + e] + bj);
        }
      }
    }
  }
}

}  // namespace

int launch_pose_head(const __nv_bfloat16* tokens, const __nv_bfloat16* w, const float* bias, void* heatmaps,
                     int out_dtype, int B, int F, int J, cudaStream_t stream) {
  if (J > kJPad || (4 * F) % 16 != 0 || (4 * F) % kRows != 0) {
    set_error("pose_head: unsupported J=%d F=%d", J, F);
    return -1;
  }
  const size_t smem = (size_t)kRows * F * kVPitch * 4 + (size_t)kJPad * kWPitch * 2;
  if (smem > 227 * 1024) {
    set_error("pose_head: feature side %d does not fit shared memory", F);
    return -1;
  }
  dim3 grid(4 * F / kRows, B);
  if (out_dtype == DT_F32) {
    HGR_CHECK_CUDA(
        cudaFuncSetAttribute(pose_head_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pose_head_kernel<float>
        <<<grid, kWarps * 32, smem, stream>>>(tokens, w, bias, static_cast<float*>(heatmaps), F, J);
  } else {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(pose_head_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    pose_head_kernel<__nv_bfloat16>
        <<<grid, kWarps * 32, smem, stream>>>(tokens, w, bias, static_cast<__nv_bfloat16*>(heatmaps), F, J);
  }
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hgr
