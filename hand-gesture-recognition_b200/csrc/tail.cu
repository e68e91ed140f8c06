// Memory-bound tail of the hot path.
//
// get_max_preds  - reference libs/utils.py:4-32 (numpy on the host, after a
//                  device->host copy of every heatmap).  Bit-exact contract:
//                  first index on ties, NaN wins and masks the prediction to
//                  (0, 0), x = idx % W and y = floor(idx / W) evaluated in
//                  fp32 like the reference does, zeroed where max <= 0.
// crop_normalize - reference detect.py:106-112 (and its training twin
//                  libs/load.py:46-50): HWC uint8 -> CHW float,
//                  ((v / 255) - mean[c]) / std[c] in fp32 with the ImageNet
//                  constants applied by channel INDEX (the frames are BGR).
//
// Both use 128-bit loads; the arg-max reduces through warp shuffles.
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

// "a beats b" under numpy's argmax order: NaN is maximal, otherwise larger
// value; equal values keep the smaller index.
__device__ __forceinline__ bool beats(float av, int ai, float bv, int bi) {
  const bool an = av != av, bn = bv != bv;
  if (an || bn) {
    if (an && bn) return ai < bi;
    return an;
  }
  if (av > bv) return true;
  if (av < bv) return false;
  return ai < bi;
}

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int kN = 4;
  __device__ static void load(const float* p, float (&v)[8]) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
  __device__ static float one(const float* p) { return __ldg(p); }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int kN = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = bf16_lo(w[i]);
      v[2 * i + 1] = bf16_hi(w[i]);
    }
  }
  __device__ static float one(const __nv_bfloat16* p) { return __bfloat162float(*p); }
};

template <typename T>
__global__ void __launch_bounds__(256)
max_preds_kernel(const T* __restrict__ heat, long long rows, int hw, int width, float* __restrict__ preds,
                 float* __restrict__ maxvals) {
  constexpr int N = Vec<T>::kN;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* p = heat + row * hw;

  float best = 0.f;
  int besti = 0x7fffffff;  // "nothing yet": any real element beats it through the index rule below
  bool have = false;
  const bool vec_ok = (hw % N == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
  if (vec_ok) {
    for (int i = lane * N; i < hw; i += 32 * N) {
      float v[8];
      Vec<T>::load(p + i, v);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (!have || beats(v[k], i + k, best, besti)) {
          best = v[k];
          besti = i + k;
          have = true;
        }
      }
    }
  } else {
    for (int i = lane; i < hw; i += 32) {
      const float v = Vec<T>::one(p + i);
      if (!have || beats(v, i, best, besti)) {
        best = v;
        besti = i;
        have = true;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    const bool oh = __shfl_xor_sync(0xffffffffu, (int)have, o) != 0;
    if (oh && (!have || beats(ov, oi, best, besti))) {
      best = ov;
      besti = oi;
      have = true;
    }
  }
  if (lane == 0) {
    const float fi = (float)besti, fw = (float)width;
    float x = fmodf(fi, fw);
    float y = floorf(__fdiv_rn(fi, fw));
    const float mask = best > 0.0f ? 1.0f : 0.0f;  // NaN > 0 is false, as in numpy
    x = __fmul_rn(x, mask);
    y = __fmul_rn(y, mask);
    preds[row * 2] = x;
    preds[row * 2 + 1] = y;
    maxvals[row] = best;
  }
}

template <typename TOut>
__device__ __forceinline__ void store4(TOut* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}
template <typename TOut>
__device__ __forceinline__ void store1(TOut* p, float a);
template <>
__device__ __forceinline__ void store1<float>(float* p, float a) {
  *p = a;
}
template <>
__device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float a) {
  *p = __float2bfloat16_rn(a);
}

template <typename TOut>
__global__ void __launch_bounds__(256)
crop_normalize_kernel(const uint8_t* __restrict__ hwc, TOut* __restrict__ chw, long long npix_total, int hw) {
  // 256-entry table per channel, built with the reference's operation order in IEEE fp32.
  __shared__ float lut[3][256];
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, v = i & 255;
    lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean[c]), stdv[c]);
  }
  __syncthreads();
  const bool vec_ok = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(hwc) & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(chw) & 15) == 0);
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec_ok) {
    const long long nquads = npix_total >> 2;
    for (long long qd = (long long)blockIdx.x * blockDim.x + threadIdx.x; qd < nquads; qd += stride) {
      const long long pix = qd << 2;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(hwc + pix * 3);
      const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
      uint8_t by[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        by[k] = (w0 >> (8 * k)) & 0xff;
        by[4 + k] = (w1 >> (8 * k)) & 0xff;
        by[8 + k] = (w2 >> (8 * k)) & 0xff;
      }
      const long long img = pix / hw;
      const long long off = pix - img * hw;
      TOut* dst = chw + img * 3 * hw + off;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        store4<TOut>(dst + (long long)c * hw, lut[c][by[c]], lut[c][by[3 + c]], lut[c][by[6 + c]], lut[c][by[9 + c]]);
    }
  } else {
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix_total; pix += stride) {
      const long long img = pix / hw;
      const long long off = pix - img * hw;
#pragma unroll
      for (int c = 0; c < 3; ++c) store1<TOut>(chw + img * 3 * hw + (long long)c * hw + off, lut[c][hwc[pix * 3 + c]]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Crop front-end (reference detect.py:92-117): cv2.warpAffine(frame, trans, (S, S),
// INTER_LINEAR) of the detector box followed by the crop normalisation, fused - the uint8 crop never exists.
// OpenCV's warpAffine is integer arithmetic and is reproduced bit for bit: the inverse map is evaluated in
// double precision exactly as imgwarp.cpp does (no FMA contraction), source coordinates are rounded to 1/1024 px
// (AB_BITS = 10), reduced to 1/32 px (INTER_BITS = 5), the four taps are weighted with
// (32 - fx)(32 - fy) * 32, ... out of 2^15 and the result is (sum + 2^14) >> 15; taps outside the frame read 0
// (BORDER_CONSTANT).  One thread per output pixel, three planes written with coalesced stores.
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256)
crop_warp_normalize_kernel(const uint8_t* __restrict__ frames, int Hf, int Wf, const int* __restrict__ frame_index,
                           const double* __restrict__ inv_mats, int S, TOut* __restrict__ out) {
  __shared__ float lut[3][256];
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, v = i & 255;
    lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), mean[c]), stdv[c]);
  }
  __syncthreads();
  const int n = blockIdx.y;
  const double* m = inv_mats + (size_t)n * 6;
  const uint8_t* img = frames + (size_t)frame_index[n] * Hf * Wf * 3;
  const int hw = S * S;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += gridDim.x * blockDim.x) {
    const int y = pix / S, x = pix - y * S;
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0));
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    int sx = X >> 5, sy = Y >> 5;
    sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);  // saturate_cast<short>
    sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    int acc[3] = {0, 0, 0};
    const bool x0ok = sx >= 0 && sx < Wf, x1ok = sx + 1 >= 0 && sx + 1 < Wf;
    const bool y0ok = sy >= 0 && sy < Hf, y1ok = sy + 1 >= 0 && sy + 1 < Hf;
    if (y0ok) {
      const uint8_t* row = img + (size_t)sy * Wf * 3;
      if (x0ok) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w00 * row[sx * 3 + c];
      }
      if (x1ok) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w01 * row[(sx + 1) * 3 + c];
      }
    }
    if (y1ok) {
      const uint8_t* row = img + (size_t)(sy + 1) * Wf * 3;
      if (x0ok) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w10 * row[sx * 3 + c];
      }
      if (x1ok) {
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += w11 * row[(sx + 1) * 3 + c];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v = (acc[c] + (1 << 14)) >> 15;
      store1<TOut>(out + ((size_t)n * 3 + c) * hw + pix, lut[c][v > 255 ? 255 : v]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PCK-style pose accuracy (reference libs/metrics.py:6-62) on decoded keypoints:
//   dist[j][n] = || pred[n,j] / norm - target[n,j] / norm ||_2  with norm = (h / 10, w / 10) applied to (x, y)
//                (the reference's own pairing), -1 where the target keypoint is not > 1 in both coordinates;
//   acc[j + 1] = #(dist < thr) / #(dist != -1), or -1 when no sample is valid;  acc[0] = mean of the valid joints.
// Arithmetic in float64 like numpy's (float32 / float64 promotes), one CTA per joint, ordered integer counts.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pck_joint_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, int J, double nx, double ny,
                 double thr, int* __restrict__ counts /*[J][2] = valid, hit*/) {
  __shared__ int s_valid[256], s_hit[256];
  const int j = blockIdx.x;
  int valid = 0, hit = 0;
  for (int n = threadIdx.x; n < B; n += 256) {
    const float tx = target[((size_t)n * J + j) * 2], ty = target[((size_t)n * J + j) * 2 + 1];
    if (tx > 1.0f && ty > 1.0f) {
      const float px = pred[((size_t)n * J + j) * 2], py = pred[((size_t)n * J + j) * 2 + 1];
      const double dx = __dsub_rn(__ddiv_rn((double)px, nx), __ddiv_rn((double)tx, nx));
      const double dy = __dsub_rn(__ddiv_rn((double)py, ny), __ddiv_rn((double)ty, ny));
      const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      ++valid;
      hit += d < thr ? 1 : 0;
    }
  }
  s_valid[threadIdx.x] = valid;
  s_hit[threadIdx.x] = hit;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_valid[threadIdx.x] += s_valid[threadIdx.x + o];
      s_hit[threadIdx.x] += s_hit[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[2 * j] = s_valid[0];
    counts[2 * j + 1] = s_hit[0];
  }
}

// acc (J + 1 doubles), avg_cnt = {avg_acc, cnt}
__global__ void pck_final_kernel(const int* __restrict__ counts, int J, double* __restrict__ acc,
                                 double* __restrict__ avg_cnt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sum = 0.0;
  int cnt = 0;
  for (int j = 0; j < J; ++j) {
    const int v = counts[2 * j], h = counts[2 * j + 1];
    const double a = v > 0 ? (double)h * 1.0 / (double)v : -1.0;
    acc[j + 1] = a;
    if (a >= 0) {
      sum = sum + a;
      ++cnt;
    }
  }
  const double avg = cnt != 0 ? sum / (double)cnt : 0.0;
  acc[0] = cnt != 0 ? avg : 0.0;
  avg_cnt[0] = avg;
  avg_cnt[1] = (double)cnt;
}

}  // namespace

int launch_get_max_preds(const void* heatmaps, int dtype, long long rows, int hw, int width, float* preds,
                         float* maxvals, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (hw <= 0 || width <= 0) {
    set_error("get_max_preds: empty heatmap (hw=%d width=%d)", hw, width);
    return -1;
  }
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  if (dtype == DT_F32)
    max_preds_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(heatmaps), rows, hw, width, preds,
                                                         maxvals);
  else
    max_preds_kernel<__nv_bfloat16>
        <<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(heatmaps), rows, hw, width, preds, maxvals);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_crop_normalize(const uint8_t* hwc, void* chw, int out_dtype, int B, int H, int W, cudaStream_t stream) {
  const long long npix = (long long)B * H * W;
  if (npix <= 0) return 0;
  long long want = (npix / 4 + 255) / 256;
  const int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
  if (out_dtype == DT_F32)
    crop_normalize_kernel<float><<<blocks, 256, 0, stream>>>(hwc, static_cast<float*>(chw), npix, H * W);
  else
    crop_normalize_kernel<__nv_bfloat16>
        <<<blocks, 256, 0, stream>>>(hwc, static_cast<__nv_bfloat16*>(chw), npix, H * W);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pose_accuracy(const float* pred, const float* target, int B, int J, int H, int W, double thr, int* counts,
                         double* acc, double* avg_cnt, cudaStream_t stream) {
  if (B < 0 || J < 1 || H < 1 || W < 1) {
    set_error("pose_accuracy: bad shape (%d, %d, %d, %d)", B, J, H, W);
    return -1;
  }
  // norm = np.ones((B, 2)) * np.array([h, w]) / 10: x is divided by h / 10, y by w / 10 (libs/metrics.py:45)
  pck_joint_kernel<<<J, 256, 0, stream>>>(pred, target, B, J, (double)H / 10.0, (double)W / 10.0, thr, counts);
  pck_final_kernel<<<1, 32, 0, stream>>>(counts, J, acc, avg_cnt);
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_crop_warp_normalize(const uint8_t* frames, int F, int Hf, int Wf, const int* frame_index,
                               const double* inv_mats, int N, int S, void* out, int out_dtype, cudaStream_t stream) {
  if (N <= 0) return 0;
  if (F < 1 || Hf < 1 || Wf < 1 || S < 1 || Hf > 32767 || Wf > 32767) {
    set_error("crop_warp_normalize: bad shape (frames %d x %d x %d, crop %d)", F, Hf, Wf, S);
    return -1;
  }
  dim3 grid((S * S + 255) / 256, N);
  if (out_dtype == DT_F32)
    crop_warp_normalize_kernel<float><<<grid, 256, 0, stream>>>(frames, Hf, Wf, frame_index, inv_mats, S,
                                                                static_cast<float*>(out));
  else
    crop_warp_normalize_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(frames, Hf, Wf, frame_index, inv_mats, S,
                                                                        static_cast<__nv_bfloat16*>(out));
  HGR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hgr
