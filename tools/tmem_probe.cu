// Hardware probe (development tool, not part of the library).  Two questions the attention kernel's design
// depends on:
//   1. tcgen05.mma with the A operand in TENSOR MEMORY: a [128 x 160] bf16 matrix written by one thread per row with
//      tcgen05.st.32x32b (two bf16 per 32-bit column, even k in the low half) against an MN-major SWIZZLE_128B B
//      operand ([160 keys][64 dims] box as TMA delivers it) must give D = P V exactly (small-integer data).
//   2. Tensor-memory read / write bandwidth per SM: cycles per tcgen05.ld.32x32b.x32 (4 KiB per warp instruction) and
//      per tcgen05.st.32x32b.x32 with 1, 4, 8 and 16 warps issuing back to back.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../hand-gesture-recognition_b200/csrc/ptx.cuh"

using namespace hgr;

constexpr int kKeys = 160;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// head: which 32-column half of the V box is the B operand; col0: TMEM column where P starts; dcol: where D goes.
__global__ void __launch_bounds__(128, 1)
ts_probe_kernel(const __grid_constant__ CUtensorMap tmV, const __nv_bfloat16* __restrict__ P, float* out, int head,
                int col0, int dcol) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sV = smem;  // 160 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sV + kKeys * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kKeys * 128);
    tma_load_2d(sV, &tmV, bar, 0, 0);
  }
  // every thread writes its row of P into tensor memory: 160 bf16 = 80 columns
  const int row = warp * 32 + lane;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < 80; c += 8) {
    uint32_t v[8];
    for (int e = 0; e < 8; ++e) {
      const __nv_bfloat16 lo = P[row * kKeys + 2 * (c + e)], hi = P[row * kKeys + 2 * (c + e) + 1];
      v[e] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
    }
    tmem_st_32x32b_x8(t_lane + col0 + c, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  mbar_wait(bar, 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 32) | (1u << 16);  // B MN-major
    for (int k = 0; k < kKeys / 16; ++k) {
      const uint64_t bd = desc_mn_sw128(smem_u32(sV) + head * 64 + k * 2048, 1024);
      umma_bf16_ts(tmem + dcol, tmem + col0 + 8 * k, bd, idesc, k != 0);
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32b_x32(t_lane + dcol, v);
  tmem_ld_wait();
  for (int e = 0; e < 32; ++e) out[row * 32 + e] = __uint_as_float(v[e]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// mode 0: loads, mode 1: stores.  Each warp issues `iters` x32 instructions on its own lane quarter.
__global__ void __launch_bounds__(512, 1) bw_probe_kernel(long long* cycles, int iters, int mode, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_lane = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t v[32];
  for (int e = 0; e < 32; ++e) v[e] = threadIdx.x + e;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      tmem_ld_32x32b_x32(t_lane + ((i * 32) & 255) + (warp >> 2) * 32 % 256, v);
      if ((i & 3) == 3) {
        tmem_ld_wait();
        acc ^= v[i & 31];
      }
    }
    tmem_ld_wait();
  } else {
    for (int i = 0; i < iters; ++i) {
      tmem_st_32x32b_x32(t_lane + ((i * 32) & 255), v);
      if ((i & 3) == 3) tmem_st_wait();
    }
    tmem_st_wait();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  if (acc == 0x12345678u) *sink = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}

int main() {
  // ---------------- 1. TS MMA ----------------
  std::vector<__nv_bfloat16> hP(128 * kKeys), hV(kKeys * 64);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < kKeys; ++k) hP[m * kKeys + k] = __float2bfloat16((float)((m * 7 + k * 3) % 11 - 5));
  for (int k = 0; k < kKeys; ++k)
    for (int n = 0; n < 64; ++n) hV[k * 64 + n] = __float2bfloat16((float)((k * 5 + n * 2 + (n >> 5)) % 7 - 3));
  __nv_bfloat16 *dP, *dV;
  float* dOut;
  cudaMalloc(&dP, hP.size() * 2);
  cudaMalloc(&dV, hV.size() * 2);
  cudaMalloc(&dOut, 128 * 32 * 4);
  cudaMemcpy(dP, hP.data(), hP.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dV, hV.data(), hV.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmV;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)kKeys};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)kKeys};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(&tmV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dV, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d\n", (int)r);
      return 1;
    }
  }
  const int smem_bytes = kKeys * 128 + 64;
  cudaFuncSetAttribute(ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<float> hOut(128 * 32);
  const int cfg[4][3] = {{0, 0, 128}, {1, 0, 128}, {0, 160, 288}, {1, 320, 448}};
  for (int c = 0; c < 4; ++c) {
    cudaMemset(dOut, 0xff, 128 * 32 * 4);
    ts_probe_kernel<<<1, 128, smem_bytes>>>(tmV, dP, dOut, cfg[c][0], cfg[c][1], cfg[c][2]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("ts probe: CUDA error %s\n", cudaGetErrorString(e));
      return 2;
    }
    cudaMemcpy(hOut.data(), dOut, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    int bad = 0, first = -1;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 32; ++n) {
        float ref = 0.f;
        for (int k = 0; k < kKeys; ++k)
          ref += __bfloat162float(hP[m * kKeys + k]) * __bfloat162float(hV[k * 64 + cfg[c][0] * 32 + n]);
        if (hOut[m * 32 + n] != ref) {
          if (first < 0) first = m * 32 + n;
          ++bad;
        }
      }
    printf("ts_mma head %d pcol %3d dcol %3d: mismatches %d / 4096", cfg[c][0], cfg[c][1], cfg[c][2], bad);
    if (first >= 0) printf("  first at m=%d n=%d got %.1f", first / 32, first % 32, hOut[first]);
    printf("\n");
  }
  // ---------------- 2. bandwidth ----------------
  long long* dCyc;
  uint32_t* dSink;
  cudaMalloc(&dCyc, 8);
  cudaMalloc(&dSink, 4);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 2, 4, 8, 16}) {
      long long cyc = 0;
      for (int rep = 0; rep < 2; ++rep) {
        bw_probe_kernel<<<1, warps * 32>>>(dCyc, iters, mode, dSink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("bw probe: CUDA error %s\n", cudaGetErrorString(e));
          return 3;
        }
        cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost);
      }
      const double bytes = (double)warps * iters * 4096.0;
      printf("tmem %s warps %2d: %lld cycles, %.1f cycles per x32 instruction per warp, %.1f B/clk/SM\n",
             mode ? "st" : "ld", warps, cyc, (double)cyc / iters, bytes / (double)cyc);
    }
  return 0;
}
