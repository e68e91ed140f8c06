// Hardware probe (development tool): MUFU.EX2 issue rate per scheduler with 1, 2, 3, 4, 6 warps per scheduler,
// as a pure stream and inside the softmax instruction mix (FFMA -> EX2 -> FADD, one F2FP per two elements).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mufu_probe mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanhf_(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf_(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned pack2(float lo, float hi) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(long long* cyc, float* sink, int iters, float c, float m) {
  float v[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) v[e] = (float)(threadIdx.x + e) * 1e-3f;
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = ex2f(v[e]);
    } else if (MODE == 2) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = tanhf_(v[e]);
    } else if (MODE == 3) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = rcpf_(v[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float p0 = ex2f(fmaf(v[e], c, -m)), p1 = ex2f(fmaf(v[e + 1], c, -m));
        const float p2 = ex2f(fmaf(v[e + 2], c, -m)), p3 = ex2f(fmaf(v[e + 3], c, -m));
        l0 += p0; l1 += p1; l2 += p2; l3 += p3;
        acc ^= pack2(p0, p1) + pack2(p2, p3);
        v[e] = p0 - 1.f; v[e + 1] = p1 - 1.f; v[e + 2] = p2 - 1.f; v[e + 3] = p3 - 1.f;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  float s = l0 + l1 + l2 + l3 + __uint_as_float(acc);
#pragma unroll
  for (int e = 0; e < 32; ++e) s += v[e];
  if (s == 1234.5f) *sink = s;
}

int main() {
  long long* d;
  float* s;
  cudaMalloc(&d, 8);
  cudaMalloc(&s, 4);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int wps : {1, 2, 3, 4, 6, 8}) {
      long long c = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, wps * 128>>>(d, s, iters, 1.01f, 0.5f);
        else if (mode == 1) k<1><<<1, wps * 128>>>(d, s, iters, 1.01f, 0.5f);
        else if (mode == 2) k<2><<<1, wps * 128>>>(d, s, iters, 1.01f, 0.5f);
        else k<3><<<1, wps * 128>>>(d, s, iters, 1.01f, 0.5f);
        cudaDeviceSynchronize();
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      }
      printf("%s warps/scheduler %d: %.2f clk per warp-level EX2 per scheduler (%.2f clk per EX2 of one warp)\n",
             mode == 0 ? "pure ex2   " : (mode == 1 ? "softmax mix" : (mode == 2 ? "pure tanh  " : "pure rcp   ")), wps, (double)c / (iters * 32.0 * wps), (double)c / (iters * 32.0));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
