// Stem convolution: encoder.conv1 = Conv(3, 64, k=3, s=2) + BN + SiLU
// (reference model/gelan.py:155 via Conv.forward :55-56).
//
// K = 27 is too thin for the tcgen05 pipeline (one 64-wide k-step would be
// 58 % zero padding and the layer is bound by its 1.18 MB/image output
// anyway), so this kernel reads the NCHW fp32/bf16 crop directly, stages a
// 17-row input patch (8 output rows) in shared memory as bf16 with 128-bit loads, and runs the contraction with
// register-resident weights on mma.sync m16n8k16 (K padded 27 -> 32).  The
// output is written as NHWC bf16 in full 128-byte pixel rows, which is the
// layout every later TMA box load expects.
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kRowsOut = 8;                // output rows per CTA
constexpr int kRowsIn = 2 * kRowsOut + 1;  // 17 input rows
constexpr int kWarps = 8;
constexpr int kLeft = 8;                   // patch x index of input column 0 (index 7 is the zero padding column)

// 8 consecutive input elements as packed bf16 (one 16-byte smem store).
__device__ __forceinline__ uint4 load8(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

// Stages the 17-row input patch of row block `rb` (zero padded) into `patch`.
// bf16 input: asynchronous 16-byte copies (cp.async, zero-fill for padding); fp32 input: load, convert, store.
template <typename TIn>
__device__ __forceinline__ void stage_patch(__nv_bfloat16* patch, const TIn* xb, int rb, int S, int pitch, int tid) {
  const int ih0 = 2 * rb * kRowsOut - 1;
  const int cpr = pitch >> 3;  // chunks per patch row, including one padding chunk on each side
  const int nchunks = 3 * kRowsIn * cpr;
  if constexpr (sizeof(TIn) == 2) {
    for (int i = tid; i < nchunks; i += kWarps * 32) {
      const int ck = i % cpr;
      const int rr = i / cpr;  // c * kRowsIn + r
      const int r = rr % kRowsIn, c = rr / kRowsIn;
      const int ih = ih0 + r;
      const bool inside = ck >= 1 && ck <= (S >> 3) && ih >= 0 && ih < S;
      const TIn* src = inside ? xb + ((size_t)c * S + ih) * S + (ck - 1) * 8 : xb;
      cp_async_16(patch + rr * pitch + ck * 8, src, inside ? 16u : 0u);
    }
  } else {
    for (int i0 = tid; i0 < nchunks; i0 += 4 * kWarps * 32) {
      uint4 v[4];  // four requests in flight per thread before the first shared-memory store
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kWarps * 32;
        const int ck = i % cpr;
        const int rr = i / cpr;
        const int r = rr % kRowsIn, c = rr / kRowsIn;
        const int ih = ih0 + r;
        v[u] = make_uint4(0, 0, 0, 0);
        if (i < nchunks && ck >= 1 && ck <= (S >> 3) && ih >= 0 && ih < S)
          v[u] = load8(xb + ((size_t)c * S + ih) * S + (ck - 1) * 8);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kWarps * 32;
        if (i < nchunks) *reinterpret_cast<uint4*>(patch + (i / cpr) * pitch + (i % cpr) * 8) = v[u];
      }
    }
  }
}

// w: [64][32] bf16, k = (kh*3+kw)*3 + c, BN scale already folded in, k>=27 zero.
// grid = (splits, B): a CTA walks `nrb` consecutive 8-row blocks of one image with two patch buffers, so the
// copy of block i+1 overlaps the MMAs and the stores of block i.
// RAW = true (training): the un-normalised convolution output is written (batch-statistics BatchNorm follows
// as separate kernels), `shift` is ignored.
template <typename TIn, bool RAW = false>
__global__ void __launch_bounds__(kWarps * 32, 2)
conv1_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ w,
             const float* __restrict__ shift, int S, int nrb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int So = S >> 1;
  const int pitch = S + 16;  // [0,8): left padding (index 7 = column -1), [8, 8+S): the row, then right padding
  const int patch_elems = 3 * kRowsIn * pitch;
  __nv_bfloat16* patch0 = reinterpret_cast<__nv_bfloat16*>(smem_raw);  // 2 x [3][17][pitch]
  __nv_bfloat16* stage = patch0 + 2 * patch_elems;                     // [warps][16][64]

  const int b = blockIdx.y;
  const int rb0 = blockIdx.x * nrb;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const TIn* xb = x + (size_t)b * 3 * S * S;
  pdl_launch_dependents();
  pdl_wait();

  stage_patch<TIn>(patch0, xb, rb0, S, pitch, tid);
  cp_async_commit();

  // ---- weights -> B fragments (registers, loaded once) ------------------
  // SiLU is evaluated on h = x / 2 (x * sigmoid(x) = h + h * tanh(h)).  The halving and the BN shift ride INSIDE the
  // MMA: every weight is halved (exact in bf16), and two of the five zero-padding slots of K carry shift / 2 as a
  // bf16 hi + lo pair (k = 27, 28) against A slots forced to 1.0 - the accumulator then IS h and the epilogue is one
  // MUFU, one FMA and half a pack per output (32 FMAs and 16 registers per tile less than an explicit affine).
  uint32_t bfrag[8][2][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const __nv_bfloat16* wr = w + (nt * 8 + g) * 32 + s * 16 + 2 * t;
      bfrag[nt][s][0] = __ldg(reinterpret_cast<const uint32_t*>(wr));
      bfrag[nt][s][1] = __ldg(reinterpret_cast<const uint32_t*>(wr + 8));
      if constexpr (!RAW) {
        const uint32_t half2 = 0x3F003F00u;  // (0.5, 0.5) bf16
#pragma unroll
        for (int i = 0; i < 2; ++i)
          asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(bfrag[nt][s][i]) : "r"(bfrag[nt][s][i]), "r"(half2));
      }
    }
  if constexpr (!RAW) {
    // k = 16 s + 2 t + {0, 1, 8, 9}: k = 27 is the high half of bfrag[.][1][1] in lanes t == 1, k = 28 the low half
    // of bfrag[.][1][1] in lanes t == 2
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float hs = 0.5f * __ldg(shift + nt * 8 + g);
      const __nv_bfloat16 hi = __float2bfloat16_rn(hs);
      const __nv_bfloat16 lo = __float2bfloat16_rn(hs - __bfloat162float(hi));
      if (t == 1) bfrag[nt][1][1] = (bfrag[nt][1][1] & 0x0000FFFFu) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      if (t == 2) bfrag[nt][1][1] = (bfrag[nt][1][1] & 0xFFFF0000u) | (uint32_t)__bfloat16_as_ushort(lo);
    }
  }
  // the matching A slots hold 1.0: (value & keep) | one, applied to a[1][2] / a[1][3] (k = 16 + 2 t + 8, + 9)
  const uint32_t a_keep = RAW ? 0xFFFFFFFFu : (t == 1 ? 0x0000FFFFu : (t == 2 ? 0xFFFF0000u : 0xFFFFFFFFu));
  const uint32_t a_one = RAW ? 0u : (t == 1 ? 0x3F800000u : (t == 2 ? 0x00003F80u : 0u));
  // per-thread gather offsets of its 8 k values: k = 16*s + 2*t + {0,1,8,9}
  int koff[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = 16 * (i >> 2) + 2 * t + (i & 1) + ((i >> 1) & 1) * 8;
    // slots k >= 27 multiply zero weight columns, so they may read ANY staged (finite) element: point them at tap 0
    const int kk = k < 27 ? k : 0;
    const int c = kk % 3, kw = (kk / 3) % 3, kh = kk / 9;
    koff[i] = (c * kRowsIn + kh) * pitch + kw + (kLeft - 1);
  }
  // staging swizzle of this lane, per 8-channel group: ((nt ^ g) & 7) * 8 elements
  int swz[8];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) swz[nt] = (((nt ^ g) & 7) << 3) + 2 * t;

  __nv_bfloat16* st = stage + warp * 16 * 64;
  const int mtiles_per_row = So >> 4;
  const int mtiles = kRowsOut * mtiles_per_row;
  for (int it = 0; it < nrb; ++it) {
    const unsigned short* pu = reinterpret_cast<const unsigned short*>(patch0 + (it & 1) * patch_elems);
    if (it + 1 < nrb) {
      stage_patch<TIn>(patch0 + ((it + 1) & 1) * patch_elems, xb, rb0 + it + 1, S, pitch, tid);
      cp_async_commit();
      cp_async_wait<1>();  // block `it` has landed, block `it + 1` may still be in flight
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int oh0 = (rb0 + it) * kRowsOut;
    // (orow, tcol) walk the 16-pixel tiles of this warp without divisions
    int orow = warp / mtiles_per_row, tcol = warp % mtiles_per_row;
    for (int mt = warp; mt < mtiles; mt += kWarps, tcol += kWarps) {
      while (tcol >= mtiles_per_row) {
        tcol -= mtiles_per_row;
        ++orow;
      }
      const int ow0 = tcol << 4;
      // pixel (orow, ow0+g) and (orow, ow0+g+8): patch offset of tap (0,0), channel 0
      const int base0 = (2 * orow) * pitch + 2 * (ow0 + g);
      const int base1 = base0 + 16;
      uint32_t a[2][4];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        unsigned short e[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ko = koff[s * 4 + i];
          e[i] = pu[base0 + ko];
          e[4 + i] = pu[base1 + ko];
        }
        a[s][0] = (uint32_t)e[0] | ((uint32_t)e[1] << 16);  // row g,   k 2t,2t+1
        a[s][1] = (uint32_t)e[4] | ((uint32_t)e[5] << 16);  // row g+8, k 2t,2t+1
        a[s][2] = (uint32_t)e[2] | ((uint32_t)e[3] << 16);  // row g,   k 2t+8,2t+9
        a[s][3] = (uint32_t)e[6] | ((uint32_t)e[7] << 16);  // row g+8, k 2t+8,2t+9
      }
      a[1][2] = (a[1][2] & a_keep) | a_one;  // slots 27 / 28: 1.0 against shift / 2 (hi, lo)
      a[1][3] = (a[1][3] & a_keep) | a_one;
      __syncwarp();  // previous tile's staging reads are done
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16_16816(d, a[0], bfrag[nt][0][0], bfrag[nt][0][1]);
        mma_bf16_16816(d, a[1], bfrag[nt][1][0], bfrag[nt][1][1]);
        float h[4];
        if constexpr (RAW) {
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = d[i];
        } else {
          // d is (conv + shift) / 2 already
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = fmaf(d[i], tanh_approx(d[i]), d[i]);
        }
        // staging is [16 pixels][64 ch]; XOR the 16-byte chunk with the pixel to spread banks
        *reinterpret_cast<uint32_t*>(st + g * 64 + swz[nt]) = pack_bf16x2(h[0], h[1]);
        *reinterpret_cast<uint32_t*>(st + (g + 8) * 64 + swz[nt]) = pack_bf16x2(h[2], h[3]);
      }
      __syncwarp();
      // 16 pixels x 128 B are contiguous in NHWC: 4 fully coalesced 512-byte stores
      __nv_bfloat16* orow_ptr = out + (((size_t)b * So + (oh0 + orow)) * So + ow0) * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = i * 32 + lane;  // 16-byte chunk index in the 2 KB tile
        const int px = idx >> 3, chunk = idx & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(st + px * 64 + (((chunk ^ px) & 7) << 3));
        *reinterpret_cast<uint4*>(orow_ptr + px * 64 + chunk * 8) = v;
      }
    }
    __syncthreads();  // everyone is done with this patch buffer before block it+2 is copied into it
  }
}

}  // namespace

template <bool RAW>
static int launch_conv1_impl(const void* x, int x_dtype, __nv_bfloat16* out, const __nv_bfloat16* w, const float* shift,
                             int B, int S, cudaStream_t stream) {
  if (S % 32 != 0 || S < 32 || S > 1024) {
    set_error("conv1: image side %d must be a multiple of 32 in [32, 1024]", S);
    return -1;
  }
  const int pitch = S + 16;
  const size_t smem = (size_t)2 * 3 * kRowsIn * pitch * 2 + (size_t)kWarps * 16 * 64 * 2;
  const int blocks = (S / 2) / kRowsOut;          // 8-row blocks per image
  const int splits = blocks % 2 == 0 ? 2 : 1;     // CTAs per image
  dim3 grid(splits, B);
  if (x_dtype == DT_F32) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(conv1_kernel<float, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HGR_CHECK_CUDA(launch_pdl(conv1_kernel<float, RAW>, grid, dim3(kWarps * 32), smem, stream, static_cast<const float*>(x), out, w,
                              shift, S, blocks / splits));
  } else {
    HGR_CHECK_CUDA(
        cudaFuncSetAttribute(conv1_kernel<__nv_bfloat16, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HGR_CHECK_CUDA(launch_pdl(conv1_kernel<__nv_bfloat16, RAW>, grid, dim3(kWarps * 32), smem, stream,
                              static_cast<const __nv_bfloat16*>(x), out, w, shift, S, blocks / splits));
  }
  return 0;
}

// The tcgen05 kernel (conv1_tc.cu) serves every image side the plans accept (multiples of 64) unless HGR_CONV1_TC=0;
// this mma.sync kernel remains for other sides (hgr_conv1 accepts multiples of 32) and as the A/B reference.
int launch_conv1(const void* x, int x_dtype, __nv_bfloat16* out, const __nv_bfloat16* w, const float* shift, int B,
                 int S, cudaStream_t stream) {
  if (conv1_tc_enabled() && conv1_tc_supported(S)) return launch_conv1_tc(x, x_dtype, out, w, shift, B, S, false, stream);
  return launch_conv1_impl<false>(x, x_dtype, out, w, shift, B, S, stream);
}

int launch_conv1_raw(const void* x, int x_dtype, __nv_bfloat16* out, const __nv_bfloat16* w, int B, int S,
                     cudaStream_t stream) {
  if (conv1_tc_enabled() && conv1_tc_supported(S)) return launch_conv1_tc(x, x_dtype, out, w, nullptr, B, S, true, stream);
  return launch_conv1_impl<true>(x, x_dtype, out, w, nullptr, B, S, stream);
}

}  // namespace hgr
