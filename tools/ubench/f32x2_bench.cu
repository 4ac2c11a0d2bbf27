// Microbenchmark: throughput of add.rn.f32x2 / fma.rn.f32x2 against scalar add.rn.f32 / fma.rn.f32 per SM sub-partition.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void bench(int iters, float seed, long long *out, float *sink) {
  float a[64];
#pragma unroll
  for (int j = 0; j < 64; j++) a[j] = seed * j + threadIdx.x;
  float b0 = seed, b1 = seed * 2;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {  // 64 scalar adds
#pragma unroll
      for (int j = 0; j < 64; j++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(b0));
    } else if (MODE == 1) {  // 32 packed adds (same 64 results)
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        unsigned long long x, y;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[j]), "f"(a[j + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b0), "f"(b1));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[j]), "=f"(a[j + 1]) : "l"(x));
      }
    } else if (MODE == 2) {  // 64 scalar fmas
#pragma unroll
      for (int j = 0; j < 64; j++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b0), "f"(b1));
    } else {  // 32 packed fmas
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        unsigned long long x, y;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[j]), "f"(a[j + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b0), "f"(b1));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x) : "l"(y));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[j]), "=f"(a[j + 1]) : "l"(x));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 64; j++) s += a[j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  long long *out;
  float *sink;
  cudaMallocManaged(&out, 148 * sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  const char *names[] = {"64 x add.rn.f32", "32 x add.rn.f32x2", "64 x fma.rn.f32", "32 x fma.rn.f32x2"};
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 4; mode++) {
      if (mode == 0) bench<0><<<148, warps * 32>>>(iters, 1e-9f, out, sink);
      if (mode == 1) bench<1><<<148, warps * 32>>>(iters, 1e-9f, out, sink);
      if (mode == 2) bench<2><<<148, warps * 32>>>(iters, 1e-9f, out, sink);
      if (mode == 3) bench<3><<<148, warps * 32>>>(iters, 1e-9f, out, sink);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("warps/SM %2d %-20s: %7.1f clk per 64 results per warp\n", warps, names[mode], (double)out[0] / iters);
    }
  }
  return 0;
}
