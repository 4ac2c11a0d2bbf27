// Microbenchmark: throughput of tcgen05.ld / tcgen05.st (TMEM <-> registers) per SM, by warps per CTA and vector width.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu && ./tmem_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}

// mode 0: x32 load + wait each; 1: two x32 loads per wait; 2: four x16 loads per wait; 3: x32 store + wait; 4: x32 load, adds, no other work
__global__ void bench(int mode, int iters, long long *out, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t z[32];
#pragma unroll
  for (int j = 0; j < 32; j++) z[j] = threadIdx.x + j;
  // initialise the columns this warp reads
  for (int c = 0; c < 512; c += 32) st32(base + c, z);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const uint32_t col = (uint32_t)((it * 64 + (warp >> 2) * 32) & 511);
    if (mode == 0 || mode == 4) {
      uint32_t v[32];
      ld32(base + (col & 480), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j++) acc += v[j];
    } else if (mode == 1) {
      uint32_t v[32], w[32];
      ld32(base + (col & 448), v);
      ld32(base + (col & 448) + 32, w);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j++) acc += v[j] ^ w[j];
    } else if (mode == 2) {
      uint32_t a[16], b[16], c[16], d[16];
      ld16(base + (col & 448), a);
      ld16(base + (col & 448) + 16, b);
      ld16(base + (col & 448) + 32, c);
      ld16(base + (col & 448) + 48, d);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; j++) acc += a[j] ^ b[j] ^ c[j] ^ d[j];
    } else if (mode == 3) {
      z[0] = acc + it;
      st32(base + (col & 480), z);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  long long *out;
  uint32_t *sink;
  cudaMallocManaged(&out, 148 * 32 * sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * sizeof(uint32_t));
  const int iters = 2000;
  const char *names[] = {"x32 load + wait", "2 x x32 loads per wait", "4 x x16 loads per wait", "x32 store + wait"};
  for (int mode = 0; mode < 4; mode++)
    for (int warps : {1, 4, 8, 16}) {
      bench<<<148, warps * 32>>>(mode, iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long mx = 0;
      for (int w = 0; w < warps; w++) mx = out[w] > mx ? out[w] : mx;
      const double bytes_per_iter = (mode == 0 || mode == 3 ? 4096.0 : 8192.0) * warps;
      printf("%-26s warps/CTA %2d: %7.1f clk per iteration per warp, %7.1f B/clk/SM\n", names[mode], warps, (double)mx / iters,
             bytes_per_iter * iters / (double)mx);
    }
  return 0;
}
