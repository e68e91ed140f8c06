// Hardware probe (development tool, not part of the library): how does tcgen05.mma
// address a K-major SWIZZLE_128B operand whose start address is NOT 1024-byte aligned
// and whose 8-row groups are NOT 1024 bytes apart?
//
// A [512 rows x 64] bf16 array is TMA-loaded (SWIZZLE_128B) into 1024-aligned smem as
// consecutive 128-byte rows.  B is a 64x64 identity, so D[m][n] = A_view[m][n]: with
// A[r][c] = r the output tells which smem row each MMA row read, with A[r][c] = c it
// tells which 16-byte chunk each logical column came from (a swizzle-phase mismatch
// shows up as a chunk permutation).  Sweeps the start shift (in rows), the stride
// byte offset and the descriptor's base_offset field.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../hand-gesture-recognition_b200/csrc/ptx.cuh"

using namespace hgr;

constexpr int kRows = 512;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
             int shift_rows, int sbo_bytes, int base_offset) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                       // 512 rows x 128 B = 64 KiB
  uint8_t* sB = smem + kRows * 128;         // 64 rows x 128 B = 8 KiB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kRows * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);            // rows 0..255
    tma_load_2d(sA + 256 * 128, &tmA, bar, 0, 256);  // rows 256..511
    tma_load_2d(sB, &tmB, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(sA) + shift_rows * 128;
    const uint32_t b_addr = smem_u32(sB);
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = umma_desc_sw128(a_addr + k * 32, sbo_bytes, base_offset);
      const uint64_t bd = umma_desc_sw128(b_addr + k * 32, 1024);
      umma_bf16_ss(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[row * 64 + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}

static CUtensorMap make_map(void* base, uint64_t rows, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {64, rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

int main() {
  std::vector<__nv_bfloat16> hA(kRows * 64), hB(64 * 64);
  for (int i = 0; i < 64 * 64; ++i) hB[i] = __float2bfloat16((i / 64) == (i % 64) ? 1.0f : 0.0f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA = make_map(dA, kRows, 256), tmB = make_map(dB, 64, 64);
  const int smem_bytes = kRows * 128 + 64 * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<float> rowsOut(128 * 64), colsOut(128 * 64);
  int ok_total = 0, n_total = 0;
  const int sbos[3] = {1024, 1280, 2304};
  for (int si = 0; si < 3; ++si)
    for (int shift = 0; shift < 12; ++shift)
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        const int sbo = sbos[si];
        const int base_offset = bo_mode == 0 ? 0 : (shift & 7);
        for (int pass = 0; pass < 2; ++pass) {
          for (int r = 0; r < kRows; ++r)
            for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16(pass == 0 ? (float)(r % 256) : (float)c);
          cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
          cudaMemset(dOut, 0xff, 128 * 64 * 4);
          probe_kernel<<<1, 128, smem_bytes>>>(tmA, tmB, dOut, shift, sbo, base_offset);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("sbo %d shift %d bo %d: CUDA error %s\n", sbo, shift, base_offset, cudaGetErrorString(e));
            return 2;
          }
          cudaMemcpy(pass == 0 ? rowsOut.data() : colsOut.data(), dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        }
        // expected: MMA row m reads smem row shift + (m/8)*(sbo/128) + m%8, columns in order
        int bad_rows = 0, bad_cols = 0, mixed = 0;
        int first_bad = -1;
        for (int m = 0; m < 128; ++m) {
          const int want = (shift + (m / 8) * (sbo / 128) + m % 8) % 256;
          bool row_ok = true, col_ok = true, uniform = true;
          for (int n = 0; n < 64; ++n) {
            if (rowsOut[m * 64 + n] != rowsOut[m * 64]) uniform = false;
            if (rowsOut[m * 64 + n] != (float)want) row_ok = false;
            if (colsOut[m * 64 + n] != (float)n) col_ok = false;
          }
          if (!uniform) ++mixed;
          if (!row_ok) ++bad_rows;
          if (!col_ok) ++bad_cols;
          if ((!row_ok || !col_ok) && first_bad < 0) first_bad = m;
        }
        ++n_total;
        if (!bad_rows && !bad_cols) ++ok_total;
        printf("sbo %4d shift %2d base_offset %d : bad_rows %3d bad_cols %3d mixed_rows %3d", sbo, shift, base_offset,
               bad_rows, bad_cols, mixed);
        if (first_bad >= 0) {
          printf("  first bad m=%d got row %.0f (want %d) cols:", first_bad, rowsOut[first_bad * 64],
                 (shift + (first_bad / 8) * (sbo / 128) + first_bad % 8) % 256);
          for (int n = 0; n < 64; n += 8) printf(" %.0f", colsOut[first_bad * 64 + n]);
        }
        printf("\n");
      }
  printf("configs fully as expected: %d / %d\n", ok_total, n_total);
  return 0;
}
