// Stem convolution encoder.conv1 = Conv(3, 64, k=3, s=2) + BN + SiLU (reference model/gelan.py:155 via
// Conv.forward :55-56) with the contraction on the 5th-gen tensor cores.
//
// conv1.cu (mma.sync) spends 236 instructions per 16-pixel tile - fragment gathers through 2-byte shared loads, 16
// HMMA, the SiLU epilogue, a staging round trip - and is bound by its issue slots (58 % busy, 3.7 TB/s of the 6.45
// TB/s its bytes could move at).  Here the im2col rows are built ONCE per pixel as a K-major operand tile, the MMA is
// two tcgen05 instructions per 128 pixels, and the epilogue reads the accumulator row of its pixel straight out of
// TMEM, so per output element only the affine, the SiLU and half a pack remain:
//
//   gather teams (2 x 4 warps, alternating tiles):  a CTA walks bands of 4 output rows; the 9 x 3 input rows of a band are staged in shared memory
//       as bf16 (cp.async for bf16 input, load + convert for fp32), then every thread builds the 27 taps of ONE pixel
//       as a 64-byte row of the [128 pixels][K = 32] A tile (SWIZZLE_128B, K order chosen so that the three taps of a
//       (kh, c) group arrive as one 16-bit and one 32-bit load and need five byte-permutes per pixel to pack);
//   MMA warp:  D[128 pixels][64 ch] = A[128][32] * W[64][32]^T, fp32 in TMEM, two accumulator stages;
//   epilogue groups (2 x 4 warps, alternating tiles):  TMEM -> + shift -> SiLU -> bf16 -> swizzled staging ->
//       one TMA store of the tile's 128 x 128 B (the pixels of a band are contiguous in NHWC).
// Bound: HBM (221 KB in + 1.18 MB out per image at 192 x 192), then the MUFU pipe (one tanh per output).
//
// BN's scale is folded into the weights (packing.py) and K is padded 27 -> 32 exactly as for conv1.cu.  Two of the
// five padding slots carry the BN shift INTO the MMA: the A rows hold 1.0 there and the weight rows shift / 2 as a
// bf16 hi + lo pair (exact to 2^-17), and all weights are halved (exact in bf16), so the accumulator already is
// h = (conv + shift) / 2, the argument the SiLU form h + h tanh(h) wants: the epilogue is one MUFU, one FMA and half a
// pack per output.  (The first version loaded the shift from shared memory per 8 channels and had one staging buffer
// per group: 0.384 ms against conv1.cu's 0.304 ms - the epilogue threads sat in short-scoreboard stalls behind those
// loads and a quarter of their time in the barrier behind cp.async.bulk.wait_group.read of the previous tile's store.)
// RAW = true (training): the un-normalised convolution output is written, `shift` is ignored.
#include <cstdlib>

#include "epilogue_math.cuh"
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 640;     // warps 0-7 two gather teams, 8 MMA, 9 TMEM allocator, 12-19 two epilogue groups
constexpr int kGatherThreads = 256;
constexpr int kBandRows = 4;      // output rows per band
constexpr int kRowsIn = 2 * kBandRows + 1;
constexpr int kLeft = 8;          // patch index of input column 0 (index 7 is column -1, the zero padding)
constexpr int kATileBytes = 128 * 128;
constexpr int kWBytes = 64 * 128;
constexpr int kOutBytes = 128 * 128;
constexpr int kOffA = 0;
constexpr int kOffW = 2 * kATileBytes;
constexpr int kOffOut = kOffW + kWBytes;
constexpr int kOffBars = kOffOut + 4 * kOutBytes;  // two staging buffers per epilogue group
constexpr int kOffTmemPtr = kOffBars + 8 * 8;
constexpr int kOffPatch = kOffTmemPtr + 64;  // 2 x [3][9][S + 16] bf16 follow

// position of tap (kh, kw, c) in a row of the A tile (and in the re-ordered weight rows): group g = kh * 3 + c;
// kw = 1, 2 -> slots 2 g, 2 g + 1 (one 32-bit load), kw = 0 -> slot 18 + g (one 16-bit load); slots 27, 28 carry the
// BN shift (A holds 1.0, W holds shift / 2 as bf16 hi + lo), slots 29-31 are zero
__host__ __device__ constexpr int tap_slot(int kh, int kw, int c) {
  const int g = kh * 3 + c;
  return kw == 0 ? 18 + g : 2 * g + (kw - 1);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ uint4 load8(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

// Stages the 9-row input patch of band `rb` (zero padded) into `patch` with the 256 gather threads.
template <typename TIn>
__device__ __forceinline__ void stage_patch(__nv_bfloat16* patch, const TIn* xb, int rb, int S, int pitch, int tid) {
  const int ih0 = 2 * rb * kBandRows - 1;
  const int cpr = pitch >> 3;  // 16-byte chunks per patch row, including one padding chunk on each side
  const int nchunks = 3 * kRowsIn * cpr;
  if constexpr (sizeof(TIn) == 2) {
    for (int i = tid; i < nchunks; i += kGatherThreads) {
      const int ck = i % cpr, rr = i / cpr;  // rr = c * kRowsIn + r
      const int r = rr % kRowsIn, c = rr / kRowsIn;
      const int ih = ih0 + r;
      const bool inside = ck >= 1 && ck <= (S >> 3) && ih >= 0 && ih < S;
      const TIn* src = inside ? xb + ((size_t)c * S + ih) * S + (ck - 1) * 8 : xb;
      cp_async_16(patch + rr * pitch + ck * 8, src, inside ? 16u : 0u);
    }
  } else {
    for (int i0 = tid; i0 < nchunks; i0 += 4 * kGatherThreads) {
      uint4 v[4];  // four requests in flight per thread before the first shared-memory store
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kGatherThreads;
        const int ck = i % cpr, rr = i / cpr;
        const int r = rr % kRowsIn, c = rr / kRowsIn;
        const int ih = ih0 + r;
        v[u] = make_uint4(0, 0, 0, 0);
        if (i < nchunks && ck >= 1 && ck <= (S >> 3) && ih >= 0 && ih < S)
          v[u] = load8(xb + ((size_t)c * S + ih) * S + (ck - 1) * 8);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kGatherThreads;
        if (i < nchunks) *reinterpret_cast<uint4*>(patch + (i / cpr) * pitch + (i % cpr) * 8) = v[u];
      }
    }
  }
}

template <typename TIn, bool RAW>
__global__ void __launch_bounds__(kThreads, 1)
conv1_tc_kernel(const __grid_constant__ CUtensorMap tmO, const TIn* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                const float* __restrict__ shift, int S, int num_bands) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int So = S >> 1;
  const int pitch = S + 16;
  const int patch_elems = 3 * kRowsIn * pitch;
  const int tiles_per_band = (kBandRows * So) >> 7;
  const int bands_per_img = So / kBandRows;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* a_full = bars;         // [2] 128 arrivals (gather threads)
  uint64_t* a_empty = bars + 2;    // [2] MMA commit
  uint64_t* acc_full = bars + 4;   // [2] MMA commit
  uint64_t* acc_empty = bars + 6;  // [2] 128 arrivals (epilogue group)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  __nv_bfloat16* patch0 = reinterpret_cast<__nv_bfloat16*>(smem + kOffPatch);

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmO);
  if (warp == 8 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr_smem, 128);
    tmem_relinquish();
  }
  // ---- weights -> B operand rows [64 co][32 slots], K-major SWIZZLE_128B (only the first 64 bytes of a row are used).
  // SiLU works on h = x / 2 (epilogue_math.cuh): weights and shift are halved here, the shift rides in slots 27, 28
  for (int i = threadIdx.x; i < 64 * 32; i += kThreads) {
    const int co = i >> 5, slot = i & 31;
    float v = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (tap_slot(kh, kw, c) == slot) v = (RAW ? 1.0f : 0.5f) * __bfloat162float(w[co * 32 + (kh * 3 + kw) * 3 + c]);
    if (!RAW && (slot == 27 || slot == 28)) {
      const float hs = 0.5f * shift[co];
      const float hi = __bfloat162float(__float2bfloat16_rn(hs));
      v = slot == 27 ? hi : hs - hi;
    }
    *reinterpret_cast<__nv_bfloat16*>(smem + kOffW + co * 128 + (((slot >> 3) ^ (co & 7)) << 4) + (slot & 7) * 2) =
        __float2bfloat16_rn(v);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  const int my_bands = (int)blockIdx.x < num_bands ? (num_bands - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int my_tiles = my_bands * tiles_per_band;

  if (warp < 8) {
    // ================= gather: all eight warps stage the band's input rows; team (warp / 4) builds the A tiles with
    // tile counter & 1 == team, one row per thread, so that two tiles are under construction at any time (a tile is a
    // chain of shared loads, stores, a proxy fence and a barrier arrive: ~1100 cycles of latency, few instructions) ====
    const int tid = threadIdx.x;
    const int team = warp >> 2, ttid = tid & 127;
    const uint32_t sw = static_cast<uint32_t>(ttid & 7);
    auto band_of = [&](int k) { return (int)blockIdx.x + k * (int)gridDim.x; };
    if (my_bands > 0) {
      const int band = band_of(0);
      stage_patch<TIn>(patch0, x + (size_t)(band / bands_per_img) * 3 * S * S, band % bands_per_img, S, pitch, tid);
      cp_async_commit();
    }
    int tc = 0;
    for (int k = 0; k < my_bands; ++k) {
      if (k + 1 < my_bands) {
        const int nb = band_of(k + 1);
        stage_patch<TIn>(patch0 + ((k + 1) & 1) * patch_elems, x + (size_t)(nb / bands_per_img) * 3 * S * S,
                         nb % bands_per_img, S, pitch, tid);
        cp_async_commit();
        cp_async_wait<1>();  // band k has landed, band k + 1 may still be in flight
      } else {
        cp_async_wait<0>();
      }
      bar_sync(1, kGatherThreads);  // every gather thread's part of the patch is visible
      const unsigned short* pu = reinterpret_cast<const unsigned short*>(patch0 + (k & 1) * patch_elems);
      // pixel (orow, ow) of this thread in tile 0 of the band; the next tile is 128 pixels further
      int orow = ttid / So, ow = ttid - orow * So;
      for (int t = 0; t < tiles_per_band; ++t, ++tc) {
        if ((tc & 1) == team) {
          // tap (kh, kw) of channel c: patch[(c * 9 + 2 orow + kh) * pitch + kLeft - 1 + 2 ow + kw]
          const int base = (2 * orow) * pitch + (kLeft - 1) + 2 * ow;
          uint32_t lo[9], pr[9];  // kw = 0 (16 bits) and kw = 1, 2 (32 bits) of the nine (kh, c) groups
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int off = base + (c * kRowsIn + kh) * pitch;
              lo[kh * 3 + c] = pu[off];
              pr[kh * 3 + c] = *reinterpret_cast<const uint32_t*>(pu + off + 1);
            }
          uint32_t row[16];
#pragma unroll
          for (int g = 0; g < 9; ++g) row[g] = pr[g];  // slots 2 g, 2 g + 1
#pragma unroll
          for (int g = 0; g < 8; g += 2) row[9 + (g >> 1)] = lo[g] | (lo[g + 1] << 16);  // slots 18 + g, 19 + g
          row[13] = lo[8] | 0x3F800000u;  // slot 26, and 1.0 in slot 27 (times shift_hi / 2)
          row[14] = 0x00003F80u;          // 1.0 in slot 28 (times shift_lo / 2), slot 29 zero
          row[15] = 0u;
          mbar_wait(&a_empty[team], ((tc >> 1) & 1) ^ 1);
          uint8_t* dst = smem + kOffA + team * kATileBytes + ttid * 128;
#pragma unroll
          for (int v = 0; v < 4; ++v)
            *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(v) ^ sw) << 4)) =
                make_uint4(row[4 * v], row[4 * v + 1], row[4 * v + 2], row[4 * v + 3]);
          fence_proxy_async_smem();
          mbar_arrive(&a_full[team]);
        }
        ow += 128;
        while (ow >= So) {
          ow -= So;
          ++orow;
        }
      }
      bar_sync(1, kGatherThreads);  // everyone is done with this patch buffer before band k + 2 is copied into it
    }
  } else if (warp == 8) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint32_t abase = smem_u32(smem + kOffA), wbase = smem_u32(smem + kOffW);
    for (int tc = 0; tc < my_tiles; ++tc) {
      const int ab = tc & 1;
      const uint32_t ph = (tc >> 1) & 1;
      mbar_wait(&acc_empty[ab], ph ^ 1);
      mbar_wait(&a_full[ab], ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16_ss(tmem_base + ab * 64, umma_desc_sw128(abase + ab * kATileBytes, 1024) + 2 * k,
                       umma_desc_sw128(wbase, 1024) + 2 * k, idesc, k != 0 ? 1u : 0u);
        umma_commit(&a_empty[ab]);
        umma_commit(&acc_full[ab]);
      }
      __syncwarp();
    }
  } else if (warp >= 12) {
    // ================= epilogue groups: group g takes the tiles with tile counter & 1 == g =================
    const int group = (warp - 12) >> 2;
    const int q = warp & 3;
    const int rowi = q * 32 + lane;
    const int gtid = threadIdx.x - 384 - group * 128;
    const uint32_t sw = static_cast<uint32_t>(rowi & 7);
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + group * 64;
    for (int tc = group; tc < my_tiles; tc += 2) {
      const uint32_t ph = (tc >> 1) & 1;
      const int band = (int)blockIdx.x + (tc / tiles_per_band) * (int)gridDim.x;
      const int m0 = band * (kBandRows * So) + (tc % tiles_per_band) * 128;  // first pixel of the tile in NHWC order
      // two staging buffers per group: only the store before the previous one must have left shared memory
      uint8_t* stage = smem + kOffOut + (group * 2 + ((tc >> 1) & 1)) * kOutBytes;
      if (gtid == 0) tma_store_wait_read<1>();
      mbar_wait(&acc_full[group], ph);
      tc_fence_after();
      bar_sync(2 + group, 128);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(t_row + half * 32, acc);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&acc_empty[group]);  // the registers hold the rest of the tile: the MMA warp may reuse the stage
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = v * 8 + 2 * e;
            float y0 = __uint_as_float(acc[c]), y1 = __uint_as_float(acc[c + 1]);
            if constexpr (!RAW) {
              y0 = apply_act<ACT_SILU>(y0);  // the accumulator is (conv + shift) / 2 already
              y1 = apply_act<ACT_SILU>(y1);
            }
            pk[e] = pack_bf16x2(y0, y1);
          }
          *reinterpret_cast<uint4*>(stage + rowi * 128 + ((static_cast<uint32_t>(half * 4 + v) ^ sw) << 4)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async_smem();
      bar_sync(2 + group, 128);
      if (gtid == 0) {
        tma_store_2d(&tmO, stage, 0, m0);
        tma_store_commit();
      }
    }
    if (gtid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 128);
}

template <typename TIn, bool RAW>
int launch_t(const void* x, __nv_bfloat16* out, const __nv_bfloat16* w, const float* shift, int B, int S, int num_sms,
             cudaStream_t stream) {
  const int So = S / 2;
  const size_t smem = (size_t)kOffPatch + (size_t)2 * 3 * kRowsIn * (S + 16) * 2;
  if (smem > 227 * 1024) {
    set_error("conv1_tc: image side %d does not fit shared memory", S);
    return -1;
  }
  HGR_CHECK_CUDA(cudaFuncSetAttribute(conv1_tc_kernel<TIn, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap tm;
  {
    const uint64_t dims[2] = {64, (uint64_t)B * So * So};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 128};
    if (int r = make_tensor_map_bf16(&tm, out, 2, dims, strides, box)) return r;
  }
  const int bands = B * (So / kBandRows);
  const int grid = bands < num_sms ? bands : num_sms;
  if (grid <= 0) return 0;
  HGR_CHECK_CUDA(launch_pdl(conv1_tc_kernel<TIn, RAW>, dim3(grid), dim3(kThreads), smem, stream, tm,
                            static_cast<const TIn*>(x), w, shift, S, bands));
  return 0;
}

}  // namespace

// 4 output rows per band must be a whole number of 128-pixel tiles: S / 2 a multiple of 32
bool conv1_tc_supported(int S) { return S >= 64 && S % 64 == 0 && S <= 1024; }

int launch_conv1_tc(const void* x, int x_dtype, __nv_bfloat16* out, const __nv_bfloat16* w, const float* shift, int B,
                    int S, bool raw, cudaStream_t stream) {
  if (!conv1_tc_supported(S)) {
    set_error("conv1_tc: image side %d must be a multiple of 64 in [64, 1024]", S);
    return -1;
  }
  const int sms = device_sm_count();
  if (x_dtype == DT_F32)
    return raw ? launch_t<float, true>(x, out, w, shift, B, S, sms, stream)
               : launch_t<float, false>(x, out, w, shift, B, S, sms, stream);
  return raw ? launch_t<__nv_bfloat16, true>(x, out, w, shift, B, S, sms, stream)
             : launch_t<__nv_bfloat16, false>(x, out, w, shift, B, S, sms, stream);
}

}  // namespace hgr
