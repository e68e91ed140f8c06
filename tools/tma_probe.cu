// Hardware probe (development tool): which unswizzled TMA boxes over a bf16 NCHW batch the hardware accepts.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_probe.cu
// usage: tma_probe <box0> <box1> <box2> <swizzle 0|1> <cluster 0|1> <c0> <c1>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__global__ void k(const __grid_constant__ CUtensorMap tm, int bytes, int c0, int c1, unsigned* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            (uint32_t)__cvta_generic_to_shared(smem)),
        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1), "r"(0), "r"(0)
        : "memory");
  }
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done && clock64() - t0 < (1ll << 28)) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b) : "memory");
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = done;
    const unsigned short* s = reinterpret_cast<const unsigned short*>(smem);
    for (int i = 0; i < 8; ++i) out[1 + i] = s[i];
  }
}

int main(int argc, char** argv) {
  const int b0 = atoi(argv[1]), b1 = atoi(argv[2]), b2 = atoi(argv[3]), sw = atoi(argv[4]), cl = atoi(argv[5]);
  const int c0 = atoi(argv[6]), c1 = atoi(argv[7]);
  const int S = 192, B = 2;
  uint16_t* x;
  cudaMalloc(&x, (size_t)B * 3 * S * S * 2);
  uint16_t* h = (uint16_t*)malloc((size_t)B * 3 * S * S * 2);
  for (int i = 0; i < B * 3 * S * S; ++i) h[i] = (uint16_t)(i & 0xffff);
  cudaMemcpy(x, h, (size_t)B * 3 * S * S * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  CUtensorMap tm;
  cuuint64_t gd[4] = {(cuuint64_t)S, (cuuint64_t)S, 3, (cuuint64_t)B};
  cuuint64_t gs[3] = {(cuuint64_t)S * 2, (cuuint64_t)S * S * 2, (cuuint64_t)3 * S * S * 2};
  cuuint32_t bd[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %d %d %d sw %d cluster %d c %d %d: encode %d; ", b0, b1, b2, sw, cl, c0, c1, (int)r);
  unsigned* out;
  cudaMalloc(&out, 64);
  cudaMemset(out, 0, 64);
  const int bytes = b0 * b1 * b2 * 2;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = 100 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl ? 2 : 1;
  at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, tm, bytes, c0, c1, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned ho[16] = {0};
  if (e == cudaSuccess) cudaMemcpy(ho, out, 64, cudaMemcpyDeviceToHost);
  printf("sync %s done %u first %u %u %u %u\n", cudaGetErrorName(e), ho[0], ho[1], ho[2], ho[3], ho[4]);
  return 0;
}
