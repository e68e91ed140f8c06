// Small ViT glue kernels: LayerNorm over 256-wide token rows
// (reference model/transformer.py:33,63,114 - nn.LayerNorm(256), eps 1e-5,
// biased variance), the class-token row (transformer.py:137-139) and the
// class head  Linear(256, C)(LayerNorm(x[:, 0]))  (transformer.py:113-116,142-144).
//
// All three are memory-bound: one warp per token row, 128-bit loads, fp32
// statistics through warp shuffles.
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kDim = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Loads the 8 bf16 this lane owns of a 256-wide row and returns them as fp32.
__device__ __forceinline__ void load_row8(const __nv_bfloat16* row, int lane, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(row) + lane);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = bf16_lo(w[i]);
    v[2 * i + 1] = bf16_hi(w[i]);
  }
}

// In-register LayerNorm of one row spread over a warp (8 values per lane).
__device__ __forceinline__ void layernorm8(float (&v)[8], const float* gamma, const float* beta, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / kDim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float d = v[i] - mean;
    sq += d * d;
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / kDim) + 1e-5f);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * lane);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * lane + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * lane);
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * lane + 1);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[8];
  load_row8(x + row * kDim, lane, v);
  layernorm8(v, gamma, beta, lane);
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(y + row * kDim)[lane] = o;
}

__global__ void fill_cls_kernel(__nv_bfloat16* __restrict__ tokens, const float* __restrict__ cls,
                                const float* __restrict__ cls_stats, float* __restrict__ row_stats, int B, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kDim) return;
  const int b = i / kDim, c = i % kDim;
  tokens[(size_t)b * T * kDim + c] = __float2bfloat16_rn(cls[c]);
  if (c < 2 && row_stats != nullptr) row_stats[(size_t)b * T * 2 + c] = cls_stats[c];
}

template <typename TOut>
__global__ void __launch_bounds__(256)
cls_head_kernel(const __nv_bfloat16* __restrict__ tokens, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ w, const float* __restrict__ bias,
                TOut* __restrict__ logits, int B, int T, int num_classes) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float v[8];
  load_row8(tokens + (size_t)b * T * kDim, lane, v);
  layernorm8(v, gamma, beta, lane);
  for (int c = 0; c < num_classes; ++c) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)c * kDim) + 2 * lane);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)c * kDim) + 2 * lane + 1);
    float s = v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z +
              v[7] * w1.w;
    s = warp_sum(s);
    if (lane == 0) {
      const float r = s + bias[c];
      if constexpr (sizeof(TOut) == 4)
        logits[(size_t)b * num_classes + c] = r;
      else
        logits[(size_t)b * num_classes + c] = __float2bfloat16_rn(r);
    }
  }
}

}  // namespace

int launch_layernorm(const __nv_bfloat16* x, __nv_bfloat16* y, const float* gamma, const float* beta, long long rows,
                     cudaStream_t stream) {
  if (rows <= 0) return 0;
  const long long blocks = (rows + 7) / 8;
  layernorm_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, y, gamma, beta, rows);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_fill_cls(__nv_bfloat16* tokens, const float* cls, const float* cls_stats, float* row_stats, int B, int T,
                    cudaStream_t stream) {
  const int n = B * kDim;
  HGR_CHECK_CUDA(launch_pdl(fill_cls_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, tokens, cls, cls_stats, row_stats,
                            B, T));
  return 0;
}

int launch_cls_head(const __nv_bfloat16* tokens, const float* gamma, const float* beta, const float* w,
                    const float* bias, void* logits, int out_dtype, int B, int T, int num_classes,
                    cudaStream_t stream) {
  const int blocks = (B + 7) / 8;
  if (out_dtype == DT_F32)
    HGR_CHECK_CUDA(launch_pdl(cls_head_kernel<float>, dim3(blocks), dim3(256), 0, stream, tokens, gamma, beta, w, bias,
                              static_cast<float*>(logits), B, T, num_classes));
  else
    HGR_CHECK_CUDA(launch_pdl(cls_head_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, tokens, gamma, beta, w, bias,
                              static_cast<__nv_bfloat16*>(logits), B, T, num_classes));
  return 0;
}

}  // namespace hgr
