// Microbenchmark: issue rate of tcgen05.mma.kind::f16 (M = 128, K = 16, operands K-major SWIZZLE_128B in shared memory)
// for the access patterns of the split-precision GEMM, with and without concurrent shared-memory writes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu && ./mma_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3ffff) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

// pattern 0: one MMA repeated on one accumulator (same operands)
// pattern 1: the kernel's K step: hi*hi -> D0, lo*hi -> D1, hi*lo -> D1, walking 4 K sub-steps and 3 stages
// pattern 2: like 1 but each MMA on its own accumulator (no accumulator dependency)
// pattern 3: like 1 with A operands from TMEM
// writers > 0: that many extra warps stream st.shared into an unrelated 32 KB region (stand-in for TMA writes)
__global__ void __launch_bounds__(512, 1) bench(int pattern, int n, int iters, int writers, int readers, long long *out, unsigned *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + (s0 - smem_u32(smem)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t stage_bytes = 2u * 16384u + 2u * (uint32_t)n * 128u;
      const long long t0 = clock64();
      for (int it = 0; it < iters; it++) {
        const uint32_t sa = s0 + (uint32_t)(it % 3) * stage_bytes;
        const uint64_t a_hi = desc_sw128(sa), a_lo = desc_sw128(sa + 16384), b_hi = desc_sw128(sa + 32768),
                       b_lo = desc_sw128(sa + 32768 + n * 128);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const uint64_t adv = (uint64_t)(k * 32 >> 4);
          if (pattern == 9) {
            __nanosleep(200);
          } else if (pattern == 0) {
            mma(tm, a_hi, b_hi, idesc, 1);
            mma(tm, a_hi, b_hi, idesc, 1);
            mma(tm, a_hi, b_hi, idesc, 1);
          } else if (pattern == 1) {
            mma(tm + (uint32_t)((it & 1) * n), a_hi + adv, b_hi + adv, idesc, k & 1);
            mma(tm + 2 * n, a_lo + adv, b_hi + adv, idesc, 1);
            mma(tm + 2 * n, a_hi + adv, b_lo + adv, idesc, 1);
          } else if (pattern == 2) {
            mma(tm, a_hi + adv, b_hi + adv, idesc, 1);
            mma(tm + n, a_lo + adv, b_hi + adv, idesc, 1);
            mma(tm + (n < 256 ? 2 * n : 0), a_hi + adv, b_lo + adv, idesc, 1);
          } else {
            mma_ts(tm + (uint32_t)((it & 1) * n), tm + 384 + k * 8, b_hi + adv, idesc, k & 1);
            mma_ts(tm + 2 * n, tm + 416 + k * 8, b_hi + adv, idesc, 1);
            mma_ts(tm + 2 * n, tm + 384 + k * 8, b_lo + adv, idesc, 1);
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      const long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
      stop = 1;
    }
  } else if (warp >= 8 && warp < 8 + readers) {
    // TMEM readers: tcgen05.ld x32 + wait in a loop on columns the MMAs do not touch (stand-in for the fold warps)
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + 448;
    unsigned acc = 0, loops = 0;
    const long long r0 = clock64();
    while (!stop) {
      loops++;
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(base)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j++) acc += v[j];
    }
    sink[blockIdx.x * 512 + threadIdx.x] = acc;
    if (blockIdx.x == 0 && lane == 0 && (warp == 8 || warp == 12)) printf("   reader warp %d: %.1f clk per x32 load+wait (%u loads)\n", warp, (double)(clock64() - r0) / loops, loops);
  } else if (warp <= writers) {
    // stream 16-byte stores over a 24 KB window far from the operand stages
    uint4 *w = reinterpret_cast<uint4 *>(smem + (s0 - smem_u32(smem)) + 200 * 1024);
    uint4 v = make_uint4(lane, warp, 0, 0);
    int i = 0;
    while (!stop) {
#pragma unroll
      for (int u = 0; u < 8; u++) w[((i + u) * 32 + lane) % 1536] = v;
      i += 8;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long *out;
  cudaMallocManaged(&out, 148 * sizeof(long long));
  const int smem = 226 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  const char *names[] = {"same operands, one accumulator", "split pattern (kernel)", "split pattern, 3 accumulators", "split pattern, A from TMEM"};
  unsigned *sink;
  cudaMalloc(&sink, 148 * 512 * 4);
  for (int n : {128})
    for (int pattern : {9, 1, 0})
      for (int writers : {0, 4})
      for (int readers : {0, 4, 8}) {
        printf("readers %d: ", readers);
        bench<<<148, 512, smem>>>(pattern, n, iters, writers, readers, out, sink);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long mx = 0;
        for (int b = 0; b < 148; b++) mx = out[b] > mx ? out[b] : mx;
        const double per = (double)mx / (iters * 12.0);
        printf("N=%3d %-34s writers %d: %6.1f clk per MMA (floor %d)  -> %.0f TFLOP/s at 1.9 GHz x 148\n", n, (pattern == 9 ? "no MMAs (sleep)" : names[pattern]), writers, per, n / 2,
               2.0 * 128 * n * 16 / per * 1.9e9 * 148 / 1e12);
      }
  return 0;
}
