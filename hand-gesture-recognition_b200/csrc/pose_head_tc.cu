// Pose head on the 5th-gen tensor cores (reference model/transformer.py:118-127, 146-150): tokens 1.. -> (256, F, F)
// -> bilinear x4 (align_corners=True) -> ReLU -> 1x1 conv 256 -> J (+bias), optionally followed by the keypoint
// decode of libs/utils.py:4-32 while the heatmap tile is still on chip.
//
// pose_head.cu builds mma.sync fragments on the fly (0.28-0.32 ms at batch 1024: HMMA pipe 21 %, issue 53 %, DRAM 6 %).
// Here BOTH contractions of the head are tcgen05 MMAs and the up-sampled (256, 4F, 4F) tensor exists only in TMEM:
//
//   MMA 1 (interpolation as a GEMM):  D1[128 pixels][256 ch] = U[128 pixels][K] * TOK[K][256 ch]
//       K = 3 F token positions (the three token rows a 128-pixel tile can touch), padded to a multiple of 16;
//       U holds the four bilinear weights of each pixel (bf16, K-major SWIZZLE_128B, rebuilt per tile by the aux
//       warps - it depends on the tile only), TOK is the MN-major B operand exactly as TMA delivers the token rows
//       ([positions][64 channels] boxes), fp32 accumulation in TMEM;
//   E1:  ReLU + bf16 (one cvt.rn.relu.bf16x2 per two values) from TMEM back INTO TENSOR MEMORY over the columns
//       just consumed, as the A operand of
//   MMA 2 (the 1x1 conv):  D2[128 pixels][32] = A2[128][256] * W[32][256]^T  (W resident in shared memory, J padded);
//   E2:  + bias -> NCHW heatmaps (fp32 or bf16), and / or the running arg-max per (image, joint) with numpy's
//       tie / NaN order, reduced over the CTA at the end of each image.
// A CTA owns whole images (18 tiles at 192 x 192), so the decode needs no second kernel and the 198 MB of fp32
// heatmaps that HandPipeline wrote only for max_preds_kernel to read back are never written.
//
// Warp roles (512 threads): 0 TMA producer (token boxes, 3 stages), 1 MMA issuer, 2 TMEM allocator, 4-7 aux
// (U tiles, E2), 8-15 E1 (two warps per TMEM lane quarter, 128 channels each).  Two 256-column TMEM buffers:
// D1 [0,256) -> A2 [0,64) + [128,192) (each E1 warp overwrites columns it has consumed itself) -> D2 [64,96).
//
// Numerics: the interpolation weights are bf16 (products of two fp32 lambdas rounded once; a mixed fp16 x bf16
// kind::f16 MMA raises an illegal-instruction fault on sm_100a, measured), the tokens bf16, the interpolated value
// fp32 until the single bf16 rounding in front of the second MMA - the same number of bf16 roundings as
// pose_head.cu's value + weight * slope form.  Supported: F = 4..20 in steps of 4 (K = 3 F <= 64 fits one
// swizzle row); larger maps use pose_head.cu.
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kDim = 256;
constexpr int kThreads = 512;
constexpr int kJPad = 32;
constexpr int kTokStages = 3;
constexpr int kUBytes = 128 * 128;       // U tile: [128 pixels][64 k] bf16, one swizzle row per pixel
constexpr int kWBytes = 4 * kJPad * 128; // W: 4 channel chunks of [32 joints][64 ch] bf16
constexpr int kBufCols = 256;
constexpr int kA2Hi = 128;               // A2 columns of channels 128-255
constexpr int kD2Col = 64;

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// relu(a), relu(b) -> bf16x2 (a in the low half)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// "a beats b" under numpy's argmax order (tail.cu): NaN is maximal, otherwise the larger value; ties keep the
// smaller index.
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
// the value the heatmap WOULD hold: the decode must see exactly what a stored heatmap would give back
template <typename TOut>
__device__ __forceinline__ float as_stored(float v);
template <>
__device__ __forceinline__ float as_stored<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float as_stored<__nv_bfloat16>(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

struct PoseTcParams {
  const __nv_bfloat16* w;  // [J][256]
  const float* bias;       // [J]
  void* heat;              // (B, J, 4F, 4F) or nullptr (keypoints only)
  float* preds;            // (B, J, 2) or nullptr
  float* maxvals;          // (B, J, 1)
  int B, J;
};

template <typename TOut, int F>
__global__ void __launch_bounds__(kThreads, 1)
pose_head_tc_kernel(const __grid_constant__ CUtensorMap tmTok, const PoseTcParams p) {
  constexpr int So = 4 * F;
  constexpr int kTiles = So * So / 128;        // tiles per image
  constexpr int kK = (3 * F + 15) / 16 * 16;   // contraction length of the interpolation GEMM
  constexpr int kKSteps = kK / 16;
  constexpr int kChunkBytes = kK * 128;        // one [kK positions][64 ch] token box
  constexpr int kTokBytes = 4 * kChunkBytes;
  static_assert(3 * F <= 64 && (So * So) % 128 == 0, "feature side not supported by the tcgen05 pose head");
  constexpr int kOffU = kTokStages * kTokBytes;
  constexpr int kOffW = kOffU + 2 * kUBytes;
  constexpr int kOffRed = kOffW + kWBytes;     // arg-max exchange: [4 warps][32 joints] (value, index)
  constexpr int kOffBars = kOffRed + 4 * kJPad * 8;
  static_assert(kChunkBytes % 1024 == 0 || kK % 8 == 0, "token boxes must keep the swizzle phase");

  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* tok_full = bars;                     // [3]
  uint64_t* tok_empty = bars + 3;                // [3]
  uint64_t* u_full = bars + 6;                   // [2] 128 arrivals
  uint64_t* u_empty = bars + 8;                  // [2]
  uint64_t* d1_full = bars + 10;                 // [2]
  uint64_t* a2_ready = bars + 12;                // [2] 256 arrivals
  uint64_t* d2_full = bars + 14;                 // [2]
  uint64_t* buf_free = bars + 16;                // [2] 128 arrivals
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmTok);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&tok_full[i], 1);
      mbar_init(&tok_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&u_full[i], 128);
      mbar_init(&u_empty[i], 1);
      mbar_init(&d1_full[i], 1);
      mbar_init(&a2_ready[i], 256);
      mbar_init(&d2_full[i], 1);
      mbar_init(&buf_free[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  pdl_wait();
  // ---- conv weights -> K-major SWIZZLE_128B rows, joints >= J zero ----
  for (int i = threadIdx.x; i < kJPad * (kDim / 8); i += kThreads) {
    const int j = i / (kDim / 8), c8 = i % (kDim / 8);  // 8-channel piece c8 of joint j
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < p.J) v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)j * kDim + c8 * 8));
    const int chunk = c8 >> 3, piece = c8 & 7;
    *reinterpret_cast<uint4*>(smem + kOffW + chunk * (kJPad * 128) + j * 128 + ((piece ^ (j & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int my_images = (int)blockIdx.x < p.B ? (p.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int num_units = my_images * kTiles;
  const float scale = (float)(F - 1) / (float)(So - 1);

  if (warp == 0) {
    // ================= TMA producer: the three token rows of a tile as four [kK positions][64 ch] boxes ==========
    if (elect_one_sync()) {
      int s = 0, k = 0;
      for (int im = 0; im < my_images; ++im) {
        const int b = (int)blockIdx.x + im * (int)gridDim.x;
        for (int t = 0; t < kTiles; ++t) {
          const int ty0 = (int)(scale * (float)((t * 128) / So));
          uint8_t* stage = smem + s * kTokBytes;
          mbar_wait(&tok_empty[s], (k & 1) ^ 1);
          mbar_expect_tx(&tok_full[s], kTokBytes);
          // token 0 is the class token; positions beyond the image's last token come back as zeros (finite)
#pragma unroll
          for (int n = 0; n < 4; ++n)
            tma_load_5d(stage + n * kChunkBytes, &tmTok, &tok_full[s], n * 64, 1 + ty0 * F, b, 0, 0);
          if (++s == kTokStages) {
            s = 0;
            ++k;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: MMA 1 of unit u, then MMA 2 of unit u - 1 =================
    // A = U (bf16, K-major), B = tokens (bf16, MN-major)
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64) | (1u << 16);
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, kJPad);
    const uint32_t ubase = smem_u32(smem + kOffU), wbase = smem_u32(smem + kOffW);
    int s = 0, ks = 0;
    for (int u = 0; u < num_units + 1; ++u) {
      if (u < num_units) {
        const int ub = u & 1;
        mbar_wait(&tok_full[s], ks & 1);
        mbar_wait(&u_full[ub], (u >> 1) & 1);
        if (u >= 2) mbar_wait(&buf_free[ub], ((u >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tok = smem_u32(smem + s * kTokBytes);
        const uint32_t d1 = tmem_base + ub * kBufCols;
        if (elect_one_sync()) {
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int k = 0; k < kKSteps; ++k)
              umma_bf16_ss(d1 + n * 64, umma_desc_sw128(ubase + ub * kUBytes, 1024) + 2 * k,
                           umma_desc_mn_sw128(tok + n * kChunkBytes + k * 2048, 1024), idesc1, k != 0 ? 1u : 0u);
          umma_commit(&d1_full[ub]);
          umma_commit(&u_empty[ub]);
          if (u + kTokStages < num_units) umma_commit(&tok_empty[s]);
        }
        __syncwarp();
        if (++s == kTokStages) {
          s = 0;
          ++ks;
        }
      }
      if (u >= 1) {
        const int v = u - 1, vb = v & 1;
        mbar_wait(&a2_ready[vb], (v >> 1) & 1);
        tc_fence_after();
        const uint32_t buf = tmem_base + vb * kBufCols;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kDim / 16; ++k)
            umma_bf16_ts(buf + kD2Col, buf + (k < 8 ? 8 * k : kA2Hi + 8 * (k - 8)),
                         umma_desc_sw128(wbase + (k >> 2) * (kJPad * 128), 1024) + 2 * (k & 3), idesc2,
                         k != 0 ? 1u : 0u);
          umma_commit(&d2_full[vb]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================= aux warps: U tile of unit u, then E2 (bias, store, arg-max) of unit u - 2 =================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // pixel of the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    TOut* heat = static_cast<TOut*>(p.heat);
    float* red_v = reinterpret_cast<float*>(smem + kOffRed);
    int* red_i = reinterpret_cast<int*>(smem + kOffRed + 4 * kJPad * 4);
    float best_v[kJPad > 24 ? 24 : kJPad];
    int best_i[kJPad > 24 ? 24 : kJPad];
    const bool decode = p.preds != nullptr;
    float bias_r[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) bias_r[j] = j < p.J ? __ldg(p.bias + j) : 0.f;
    for (int u = 0; u < num_units + 2; ++u) {
      if (u < num_units) {
        const int ub = u & 1, t = u % kTiles;
        const int pix = t * 128 + r, oy = pix / So, ox = pix - oy * So;
        const int ty0 = (int)(scale * (float)((t * 128) / So));
        // ATen upsample_bilinear2d, align_corners: src = dst * (in - 1) / (out - 1)
        const float sy = scale * (float)oy, sx = scale * (float)ox;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < F - 1 ? 1 : 0), x1 = x0 + (x0 < F - 1 ? 1 : 0);
        const float ly1 = sy - (float)y0, ly0 = 1.0f - ly1, lx1 = sx - (float)x0, lx0 = 1.0f - lx1;
        float w00 = ly0 * lx0, w01 = ly0 * lx1, w10 = ly1 * lx0, w11 = ly1 * lx1;
        if (x1 == x0) {  // clamped at the right border: both taps are the same token
          w00 += w01;
          w10 += w11;
          w01 = w11 = 0.f;
        }
        if (y1 == y0) {
          w00 += w10;
          w01 += w11;
          w10 = w11 = 0.f;
        }
        uint8_t* row = smem + kOffU + ub * kUBytes + r * 128;
        mbar_wait(&u_empty[ub], ((u >> 1) & 1) ^ 1);
        // rows are 128 B apart: rotate the chunk order with the lane so that the eight lanes of a store phase
        // cover all 32 banks
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(row + (((c + lane) & 7) << 4)) = make_uint4(0, 0, 0, 0);
        auto put = [&](int k, float wgt) {
          if (wgt != 0.f)
            *reinterpret_cast<__nv_bfloat16*>(row + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(wgt);
        };
        put((y0 - ty0) * F + x0, w00);
        put((y0 - ty0) * F + x1, w01);
        put((y1 - ty0) * F + x0, w10);
        put((y1 - ty0) * F + x1, w11);
        fence_proxy_async_smem();
        mbar_arrive(&u_full[ub]);
      }
      if (u >= 2) {
        const int v = u - 2, vb = v & 1, t = v % kTiles;
        const int b = (int)blockIdx.x + (v / kTiles) * (int)gridDim.x;
        mbar_wait(&d2_full[vb], (v >> 1) & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_lane + vb * kBufCols + kD2Col, o);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&buf_free[vb]);
        const int pix = t * 128 + r;
        float val[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) val[j] = __uint_as_float(o[j]) + bias_r[j];
        if (heat != nullptr) {
          // NCHW: for one joint the 32 lanes of a warp cover 32 consecutive pixels
          TOut* hp = heat + (size_t)b * p.J * (So * So) + pix;
#pragma unroll
          for (int j = 0; j < 24; ++j)
            if (j < p.J) store_out<TOut>(hp + (size_t)j * (So * So), val[j]);
        }
        if (decode) {
          if (t == 0) {
#pragma unroll
            for (int j = 0; j < 24; ++j) {
              best_v[j] = as_stored<TOut>(val[j]);
              best_i[j] = pix;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 24; ++j) {
              const float sv = as_stored<TOut>(val[j]);
              // pix grows from tile to tile, so on equal values the earlier index stays (numpy's first-occurrence rule)
              if (beats(sv, pix, best_v[j], best_i[j])) {
                best_v[j] = sv;
                best_i[j] = pix;
              }
            }
          }
        }
        if (decode && t == kTiles - 1) {
          // ---- end of the image: warp shuffles, then the four warps through shared memory ----
#pragma unroll
          for (int j = 0; j < 24; ++j) {
            if (j < p.J) {
              float bv = best_v[j];
              int bi = best_i[j];
#pragma unroll
              for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (beats(ov, oi, bv, bi)) {
                  bv = ov;
                  bi = oi;
                }
              }
              if (lane == 0) {
                red_v[q * kJPad + j] = bv;
                red_i[q * kJPad + j] = bi;
              }
            }
          }
          bar_sync(1, 128);
          if (r < p.J) {
            float bv = red_v[r];
            int bi = red_i[r];
#pragma unroll
            for (int w = 1; w < 4; ++w)
              if (beats(red_v[w * kJPad + r], red_i[w * kJPad + r], bv, bi)) {
                bv = red_v[w * kJPad + r];
                bi = red_i[w * kJPad + r];
              }
            // libs/utils.py:21-30 in fp32, like tail.cu
            const float fi = (float)bi, fw = (float)So;
            float x = fmodf(fi, fw);
            float y = floorf(__fdiv_rn(fi, fw));
            const float mask = bv > 0.0f ? 1.0f : 0.0f;
            p.preds[((size_t)b * p.J + r) * 2] = __fmul_rn(x, mask);
            p.preds[((size_t)b * p.J + r) * 2 + 1] = __fmul_rn(y, mask);
            p.maxvals[(size_t)b * p.J + r] = bv;
          }
          bar_sync(1, 128);  // the exchange buffer is free for the next image
        }
      }
    }
  } else if (warp >= 8) {
    // ================= E1: ReLU + bf16 of the interpolated tile, TMEM -> TMEM =================
    const int q = warp & 3, half = (warp - 8) >> 2;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int u = 0; u < num_units; ++u) {
      const int ub = u & 1;
      const uint32_t d1 = t_lane + ub * kBufCols + half * 128;  // this warp's 128 channels
      const uint32_t a2 = t_lane + ub * kBufCols + (half ? kA2Hi : 0);
      mbar_wait(&d1_full[ub], (u >> 1) & 1);
      tc_fence_after();
      uint32_t sb[2][32];
      tmem_ld_32x32b_x32(d1, sb[0]);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const uint32_t(&s)[32] = sb[cb & 1];
        tmem_ld_wait();
        if (cb < 3) tmem_ld_32x32b_x32(d1 + (cb + 1) * 32, sb[(cb + 1) & 1]);
        uint32_t packed[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) packed[e] = pack_relu_bf16x2(__uint_as_float(s[2 * e]), __uint_as_float(s[2 * e + 1]));
        // channels 32 cb .. 32 cb + 31 of this half -> 16 columns inside the 32 just consumed
        tmem_st_32x32b_x16(a2 + cb * 16, packed);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&a2_ready[ub]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <typename TOut, int F>
int launch_impl(const __nv_bfloat16* tokens, const PoseTcParams& p, int num_sms, cudaStream_t stream) {
  constexpr int kK = (3 * F + 15) / 16 * 16;
  constexpr int kTokBytes = 4 * kK * 128;
  constexpr int smem = kTokStages * kTokBytes + 2 * kUBytes + kWBytes + 4 * kJPad * 8 + 19 * 8 + 16;
  static_assert(smem <= 227 * 1024, "pose_head_tc shared-memory plan exceeds one CTA");
  HGR_CHECK_CUDA(cudaFuncSetAttribute(pose_head_tc_kernel<TOut, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int T = F * F + 1;
  CUtensorMap tm;
  {
    const uint64_t dims[5] = {256, (uint64_t)T, (uint64_t)p.B, 1, 1};
    const uint64_t row = 256 * 2;
    const uint64_t strides[4] = {row, row * T, row * T * p.B, row * T * p.B};
    const uint32_t box[5] = {64, (uint32_t)kK, 1, 1, 1};
    if (int r = make_tensor_map_bf16(&tm, tokens, 5, dims, strides, box)) return r;
  }
  const int grid = p.B < num_sms ? p.B : num_sms;
  if (grid <= 0) return 0;
  HGR_CHECK_CUDA(launch_pdl(pose_head_tc_kernel<TOut, F>, dim3(grid), dim3(kThreads), smem, stream, tm, p));
  return 0;
}

template <typename TOut>
int launch_f(const __nv_bfloat16* tokens, const PoseTcParams& p, int F, int num_sms, cudaStream_t stream) {
  switch (F) {
    case 4: return launch_impl<TOut, 4>(tokens, p, num_sms, stream);
    case 8: return launch_impl<TOut, 8>(tokens, p, num_sms, stream);
    case 12: return launch_impl<TOut, 12>(tokens, p, num_sms, stream);
    case 16: return launch_impl<TOut, 16>(tokens, p, num_sms, stream);
    case 20: return launch_impl<TOut, 20>(tokens, p, num_sms, stream);
    default: set_error("pose_head_tc: feature side %d not instantiated (4, 8, 12, 16, 20)", F); return -1;
  }
}

}  // namespace

bool pose_head_tc_supported(int F, int J) { return F >= 4 && F <= 20 && F % 4 == 0 && J >= 1 && J <= 24; }

int launch_pose_head_tc(const __nv_bfloat16* tokens, const __nv_bfloat16* w, const float* bias, void* heatmaps,
                        int out_dtype, float* preds, float* maxvals, int B, int F, int J, int num_sms,
                        cudaStream_t stream) {
  if (!pose_head_tc_supported(F, J)) {
    set_error("pose_head_tc: unsupported F=%d J=%d", F, J);
    return -1;
  }
  if (heatmaps == nullptr && preds == nullptr) {
    set_error("pose_head_tc: neither heatmaps nor keypoints requested");
    return -1;
  }
  if ((preds == nullptr) != (maxvals == nullptr)) {
    set_error("pose_head_tc: preds and maxvals go together");
    return -1;
  }
  PoseTcParams p;
  p.w = w;
  p.bias = bias;
  p.heat = heatmaps;
  p.preds = preds;
  p.maxvals = maxvals;
  p.B = B;
  p.J = J;
  if (out_dtype == DT_F32) return launch_f<float>(tokens, p, F, num_sms, stream);
  return launch_f<__nv_bfloat16>(tokens, p, F, num_sms, stream);
}

}  // namespace hgr
