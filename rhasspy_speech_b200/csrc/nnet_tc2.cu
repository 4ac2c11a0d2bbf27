// Stage (ii), second generation of the tensor-core affine kernel (TdnnComponent::Propagate,
// kaldi/src/nnet3/nnet-tdnn-component.cc:181-211; Affine / Linear components of nnet3/nnet-simple-component.cc).
// Same arithmetic as gemm_tc_kernel (nnet_tc.cu: three fp16 MMAs per product, the main partial sums leave TMEM
// every `fold` MMAs and are summed in fp32 registers with round-to-nearest, the cross sum is folded once with
// the exact factor 2^-11, the tail runs the layer's ops in the reference's order) -- what changes is who does what:
//
//   warp 0        TMA producer      (one lane)
//   warp 1        MMA issuer        (one lane)
//   warps 4..7    FOLD warps        one per TMEM lane quadrant; a thread owns one output row and the 128 fp32 running
//                                   sums of the tile; every published partial sum is tcgen05.ld-ed and added; at the
//                                   end of the tile the cross accumulator is updated IN PLACE in TMEM to
//                                   sum + cross * 2^-11  (tcgen05.st) and handed to the tail warps -- the fold warps
//                                   go straight on to the next tile
//   warps 8..15   TAIL warps        two per TMEM lane quadrant: each takes two 32-column chunks of the finished tile
//                                   out of TMEM and runs ReLU / BatchNorm / bypass / split / store on them; the global
//                                   inputs of the next chunk are requested while the current one is finished
// (Measured: eight fold warps of 64 columns each + four tail warps: the K = 2048 layers do not change, 104 vs 103 us --
//  they are not fold-bound -- and the 128 -> 1024 layers become tail-bound, 162-177 vs 146-155 us.)
//
// In gemm_tc_kernel the eight epilogue warps held the running sums AND ran the tail: 232 registers were not
// enough (spills), the tail was unrolled four times (a 420 KB kernel, 14 % of the stalls were instruction
// fetches), and with two warps per scheduler only 29 % of the issue slots were used (profiles/r2_gemm_details.txt).
// Here the tail is one chunk long and loops, per-column vectors are warp-uniform 16-byte loads instead of
// shuffles.  Measured on the way (tools/ubench): a tcgen05.ld x32 round trip is ~160 clk and a second load in
// flight adds ~70; the MMA rate in this operand pattern is 71 clk per 128x128x16 MMA alone and 90-110 with
// concurrent shared-memory writes; FADD2 / FFMA2 have the result rate of FADD / FFMA (half the issue slots).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "engine.h"
#include "model.h"
#include "nnet_tc.h"
#include "split.cuh"
#include "tc_ptx.cuh"

namespace rs {

namespace {

constexpr int kT2Threads = 512;
constexpr int kT2StageA = kTcBM * 128;  // one plane of the activation tile: 128 rows x 128 B

// per-column vector, four columns starting at c (warp-uniform address: one broadcast load)
template <bool FULL>
__device__ __forceinline__ float4 ldvec4(const float *v, int c, int n) {
  if constexpr (FULL) {
    return __ldg(reinterpret_cast<const float4 *>(v + c));
  } else {
    float4 r;
    r.x = c + 0 < n ? __ldg(v + c + 0) : 0.f;
    r.y = c + 1 < n ? __ldg(v + c + 1) : 0.f;
    r.z = c + 2 < n ? __ldg(v + c + 2) : 0.f;
    r.w = c + 3 < n ? __ldg(v + c + 3) : 0.f;
    return r;
  }
}

// the split of split.cuh for two values: hi = fp16(x), lo = fp16((x - hi) * 2048); x - hi and the scaling are exact
__device__ __forceinline__ void split2x(float x0, float x1, uint32_t &hi, uint32_t &lo, float &amax) {
  amax = fmaxf(amax, fmaxf(fabsf(x0), fabsf(x1)));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
  float d0 = x0, d1 = x1;
  add2(d0, d1, -f.x, -f.y);
  mul2(d0, d1, kSplitScale, kSplitScale);
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d1), "f"(d0));
}

// {a0 * b0 + c0, a1 * b1 + c1}, one rounding each (FFMA2)
__device__ __forceinline__ void fma2(float &a0, float &a1, float b0, float b1, float c0, float c1) {
  unsigned long long a, b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}

}  // namespace

// PAT >= 0: the op sequence is a compile-time constant (4 bits per op: EpiOp::Type + 1, first op in the low
// bits); PAT < 0: run-time op list.  FULL: bn == 128 and n % 128 == 0 (no column guards anywhere).
// PROF (RS_B200_TC_PROFILE=1, the two hot instantiations only): block 0 prints where each role spent its clocks.
template <int PAT, bool FULL, bool PROF = false>
__global__ void __launch_bounds__(kT2Threads, 1) gemm_tc2_kernel(const __grid_constant__ TcParams p) {
  // mbarrier wait that adds the waiting time to a counter in the profiling build
  auto timed_wait = [&](uint32_t bar, uint32_t parity, long long &acc) {
    if constexpr (PROF) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      acc += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  long long prof_t0 = 0, prof_a = 0, prof_b = 0, prof_c = 0;
  if constexpr (PROF) prof_t0 = clock64();
  constexpr bool kStatic = PAT >= 0;
  constexpr int kTypes[4] = {kStatic ? ((PAT >> 0) & 15) - 1 : -1, kStatic ? ((PAT >> 4) & 15) - 1 : -1,
                             kStatic ? ((PAT >> 8) & 15) - 1 : -1, kStatic ? ((PAT >> 12) & 15) - 1 : -1};
  // index of the (first) BatchNorm scale / offset op of a static list
  constexpr int kSoIdx = kTypes[0] == EpiOp::kScaleOffset ? 0 : kTypes[1] == EpiOp::kScaleOffset ? 1 : kTypes[2] == EpiOp::kScaleOffset ? 2
                         : kTypes[3] == EpiOp::kScaleOffset ? 3 : -1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.bn * 128u;
  const uint32_t stage_bytes = 2u * kT2StageA + 2u * b_bytes;
  // [stages][8 x 4 KB tail staging tiles][barriers]
  const uint32_t epi0 = smem0 + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar0 = epi0 + 8u * 4096u;
  // barriers: full[stages] | empty[stages] | setf[2] | sete[2] | xfull[2] | xfree[2] | tmem slot
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(p.stages + s); };
  auto setf_bar = [&](uint32_t a) { return bar0 + 8u * (uint32_t)(2 * p.stages + a); };      // partial sum published
  auto sete_bar = [&](uint32_t a) { return bar0 + 8u * (uint32_t)(2 * p.stages + 2 + a); };  // main accumulator drained
  auto xfull_bar = [&](uint32_t a) { return bar0 + 8u * (uint32_t)(2 * p.stages + 4 + a); }; // finished tile in TMEM
  auto xfree_bar = [&](uint32_t a) { return bar0 + 8u * (uint32_t)(2 * p.stages + 6 + a); }; // ... taken by the tail warps
  const uint32_t tmem_slot = bar0 + 8u * (uint32_t)(2 * p.stages + 8);
  // [2][128] floats: the bias slice of the fold warps' next tile (16-byte aligned: bar0 is, and the barrier block is 8 * even)
  float *bias_s = reinterpret_cast<float *>(smem_raw + (bar0 - smem_u32(smem_raw)) + 8u * (uint32_t)(2 * p.stages + 10));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; a++) {
      mbar_init(setf_bar(a), 1);
      mbar_init(sete_bar(a), 4);   // the four fold warps
      mbar_init(xfull_bar(a), 4);  // the four fold warps
      mbar_init(xfree_bar(a), 8);  // the eight tail warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int num_tiles = p.tiles_m * p.tiles_n;
  int total_kb = 0;
  for (int s = 0; s < p.n_slabs; s++) total_kb += p.slabs[s].kblocks;
  const int total_sums = total_kb * 2;  // partial sums per tile: one per two main MMAs (K = 16 each), see the MMA issuer
  // the bias, when it is the first op, is the start value of the running sums (AffineComponent / TdnnComponent::Propagate
  // copy the bias into the output and let the GEMM accumulate onto it)
  const bool bias_first = kStatic ? kTypes[0] == EpiOp::kBias : (p.n_ops > 0 && p.ops[0].type == EpiOp::kBias);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      if (lane == 0) {
        // ------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        bool uniform = true;
        for (int s = 1; s < p.n_slabs; s++) uniform = uniform && p.slabs[s].kblocks == p.slabs[0].kblocks;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const int m0 = (tile / p.tiles_n) * kTcBM, n0 = (tile % p.tiles_n) * p.bn;
          // K blocks block-major across equally long slabs: the time-offset slabs of a TDNN layer read the same
          // source rows shifted by a few rows, so back-to-back loads hit in L2
          for (int it = 0; it < total_kb; it++) {
            int s, kb;
            if (uniform) {
              s = it % p.n_slabs;
              kb = it / p.n_slabs;
            } else {
              s = 0;
              kb = it;
              while (kb >= p.slabs[s].kblocks) kb -= p.slabs[s++].kblocks;
            }
            const TcSlab sl = p.slabs[s];
            timed_wait(empty_bar(stage), phase ^ 1u, prof_a);
            const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes, fb = full_bar(stage);
            mbar_expect_tx(fb, stage_bytes);
            tma_load_2d(sa, &p.a_hi[s], kb * kTcBK, m0 + sl.yshift, fb);
            tma_load_2d(sa + kT2StageA, &p.a_lo[s], kb * kTcBK, m0 + sl.yshift, fb);
            tma_load_2d(sa + 2 * kT2StageA, &p.w_hi, sl.wk0 + kb * kTcBK, n0, fb);
            tma_load_2d(sa + 2 * kT2StageA + b_bytes, &p.w_lo, sl.wk0 + kb * kTcBK, n0, fb);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        if (PROF && blockIdx.x == 0)
          printf("tc2 profile (n=%d k-blocks/tile=%d tiles=%d): TMA thread total %lld clk, waiting for a free stage %lld\n", p.n, total_kb,
                 num_tiles, clock64() - prof_t0, prof_a);
      }
      __syncwarp();
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer
      // All 32 lanes walk the loop (warp-uniform control flow, so barrier addresses and matrix descriptors live in
      // uniform registers) and one elected lane issues.  The issuing thread's own instruction stream is what bounds
      // this kernel once the epilogue keeps up: measured with one lane walking the loop alone, ~78 instructions per MMA
      // at ~5 clk per dependent instruction = 160 clk per MMA against 70-100 clk of tensor work (profiles/r2_gemm3_*).
      // fold is fixed at 2 here: partial sums are published after MMAs 1 and 3 of a K block.
      const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
      // upper word of a shared-memory matrix descriptor: SBO = 1024 B, version 1, SWIZZLE_128B
      constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      auto desc = [&](uint32_t lo) {
        uint64_t d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(kDescHi));
        return d;
      };
      const uint32_t bn = (uint32_t)p.bn;
      uint32_t stage = 0, phase = 0, fcount = 0, tcount = 0;  // fcount: partial sums published so far (main set = fcount & 1)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
        const uint32_t xb = tcount & 1u;
        const uint32_t d_cross = tmem_base + (2u + xb) * bn;
        mbar_wait_lean(xfree_bar(xb), ((tcount >> 1) & 1u) ^ 1u);  // the tail warps have taken the tile before last
        tc_fence_after();
        for (int kb = 0; kb < total_kb; kb++) {
          mbar_wait_lean(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = ((smem0 + stage * stage_bytes) & 0x3ffffu) >> 4;  // descriptor address field, 16-byte units
          const uint32_t a_hi = sa, a_lo = sa + (kT2StageA >> 4), b_hi = sa + (2 * kT2StageA >> 4), b_lo = b_hi + (b_bytes >> 4);
  #pragma unroll
          for (int half = 0; half < 2; half++) {  // one published partial sum = two K steps of 16 (32 bytes inside the swizzle atom)
            const uint32_t set = fcount & 1u;
            mbar_wait_lean(sete_bar(set), ((fcount >> 1) & 1u) ^ 1u);  // the fold warps have drained this accumulator
            tc_fence_after();
            const uint32_t d_main = tmem_base + set * bn;
            if (elect_one()) {
  #pragma unroll
              for (int kk = 0; kk < 2; kk++) {
                const uint32_t adv = (uint32_t)((half * 2 + kk) * 32 >> 4);
                tc_mma_f16(d_main, desc(a_hi + adv), desc(b_hi + adv), idesc, kk);
                tc_mma_f16(d_cross, desc(a_lo + adv), desc(b_hi + adv), idesc, (kb | half | kk) != 0 ? 1u : 0u);
                tc_mma_f16(d_cross, desc(a_hi + adv), desc(b_lo + adv), idesc, 1u);
              }
              if (half == 1) tc_commit(empty_bar(stage));  // frees the smem stage when these MMAs have read it
              tc_commit(setf_bar(set));                    // this partial sum (and, on the tile's last one, the cross sum) complete
            }
            __syncwarp();
            fcount++;
          }
          if (++stage == (uint32_t)p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------------------------ fold warps: one per lane quadrant
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t fcount = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
      const int n0 = (tile % p.tiles_n) * p.bn;
      float acc[kTcMaxBN];
      if (bias_first) {
        // the bias slice was staged in shared memory during the previous tile (one column per fold thread), so the
        // start values cost 32 broadcast LDS instead of 32 L2 round trips at the head of the tile
        const float *bs = bias_s + (tcount & 1u) * 128u;
        if (tcount == 0) {
          const int c = n0 + (int)threadIdx.x - 128;
          bias_s[threadIdx.x - 128] = (int)threadIdx.x - 128 < p.bn && c < p.n ? __ldg(p.ops[0].v0 + c) : 0.f;
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < kTcMaxBN; j += 4) {
          const float4 b = *reinterpret_cast<const float4 *>(bs + j);
          acc[j] = b.x;
          acc[j + 1] = b.y;
          acc[j + 2] = b.z;
          acc[j + 3] = b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kTcMaxBN; j++) acc[j] = 0.f;
      }
      float bias_next = 0.f;  // this thread's column of the next tile's bias slice
      const int next_tile = tile + (int)gridDim.x;
      if (bias_first && next_tile < num_tiles) {
        const int c = (next_tile % p.tiles_n) * p.bn + (int)threadIdx.x - 128;
        if ((int)threadIdx.x - 128 < p.bn && c < p.n) bias_next = __ldg(p.ops[0].v0 + c);
      }
#pragma unroll 1
      for (int f = 0; f < total_sums; f++, fcount++) {
        const uint32_t set = fcount & 1u;
        timed_wait(setf_bar(set), (fcount >> 1) & 1u, prof_a);
        tc_fence_after();
        const uint32_t taddr = lane_base + set * (uint32_t)p.bn;
        // two TMEM loads in flight per wait: the load -> wait round trip, not the adds, bounds a fold
#pragma unroll
        for (int jc = 0; jc < 4; jc += 2) {
          if (!FULL && jc * 32 >= p.bn) break;
          uint32_t raw0[32], raw1[32];
          const bool two = FULL || (jc + 1) * 32 < p.bn;
          long long tl0 = 0;
          if constexpr (PROF) tl0 = clock64();
          if constexpr (PROF) {
            if (p.profile & 2) continue;  // experiment: no TMEM loads at all (results are garbage)
          }
          tmem_ld32_nowait(taddr + jc * 32, raw0);
          if (two) tmem_ld32_nowait(taddr + (jc + 1) * 32, raw1);
          tmem_wait_ld();
          if constexpr (PROF) prof_b += clock64() - tl0;
          if constexpr (PROF) {
            if (p.profile & 4) {  // experiment: loads but one add per load instead of 32
              acc[jc * 32] += __uint_as_float(raw0[0]) + __uint_as_float(raw1[1]);
              continue;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            add2(acc[jc * 32 + j], acc[jc * 32 + j + 1], __uint_as_float(raw0[j]), __uint_as_float(raw0[j + 1]));
          if (two) {
#pragma unroll
            for (int j = 0; j < 32; j += 2)
              add2(acc[(jc + 1) * 32 + j], acc[(jc + 1) * 32 + j + 1], __uint_as_float(raw1[j]), __uint_as_float(raw1[j + 1]));
          }
        }
        long long ta0 = 0;
        if constexpr (PROF) ta0 = clock64();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sete_bar(set));
        if constexpr (PROF) prof_c += clock64() - ta0;
      }
      // the tile's cross sum (the last partial sum's commit covers it): X <- sum + cross * 2^-11, in place
      const uint32_t xb = tcount & 1u;
      const uint32_t xaddr = lane_base + (2u + xb) * (uint32_t)p.bn;
#pragma unroll
      for (int jc = 0; jc < 4; jc += 2) {  // two chunks per round trip
        if (!FULL && jc * 32 >= p.bn) break;
        uint32_t raw0[32], raw1[32];
        const bool two = FULL || (jc + 1) * 32 < p.bn;
        tmem_ld32_nowait(xaddr + jc * 32, raw0);
        if (two) tmem_ld32_nowait(xaddr + (jc + 1) * 32, raw1);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j++) raw0[j] = __float_as_uint(fmaf(__uint_as_float(raw0[j]), 1.f / kSplitScale, acc[jc * 32 + j]));
        tmem_st32(xaddr + jc * 32, raw0);
        if (two) {
#pragma unroll
          for (int j = 0; j < 32; j++) raw1[j] = __float_as_uint(fmaf(__uint_as_float(raw1[j]), 1.f / kSplitScale, acc[(jc + 1) * 32 + j]));
          tmem_st32(xaddr + (jc + 1) * 32, raw1);
        }
      }
      if (bias_first && next_tile < num_tiles) {
        bias_s[((tcount + 1u) & 1u) * 128u + threadIdx.x - 128] = bias_next;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(xfull_bar(xb));
    }
    if (PROF && blockIdx.x == 0 && threadIdx.x == 128)
      printf("tc2 profile: fold warp total %lld clk, waiting for partial sums %lld, in TMEM loads %lld, fence + arrive %lld\n", clock64() - prof_t0, prof_a, prof_b, prof_c);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    // ------------------------------------------------------------------ tail warps: two per lane quadrant, two chunks each
    const int q = warp & 3, half = (warp - 8) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // this warp's 32 x 32 fp32 staging tile (float4 columns XOR-swizzled by row: conflict-free both ways)
    float4 *stg = reinterpret_cast<float4 *>(smem_raw + (epi0 - smem_u32(smem_raw)) + (uint32_t)(warp - 8) * 4096u);
    const int first_op = bias_first ? 1 : 0;
    int ib = -1;  // the op whose split bypass input is prefetched (first kAddScaled with a split source)
    for (int i = 0; i < p.n_ops && ib < 0; i++)
      if (p.ops[i].type == EpiOp::kAddScaled && p.ops[i].buf_lo) ib = i;
    // The warp walks "units" = (tile, 32-column chunk).  The global inputs of unit u + 1 -- the bypass rows and this
    // lane's column of the BatchNorm vectors -- are requested in the middle of unit u, as soon as unit u has consumed
    // its own, so that their latency hides behind the rest of unit u and the TMEM load of unit u + 1.
    uint4 pf_h[4], pf_l[4];        // bypass input of the coming unit: 4 lanes x 16 B cover a row segment, 8 rows per instruction
    float so_s = 0.f, so_o = 0.f;  // column `lane` of the BatchNorm scale / offset of the coming unit
    const int my_tiles = blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_units = my_tiles * 2;
    // coordinates of a unit: one integer division per unit, shared by everything that needs them
    struct Unit {
      int m0, c0, jc;
      bool valid;
    };
    auto unit_at = [&](int u) {
      Unit w;
      const int t = (int)blockIdx.x + (u >> 1) * (int)gridDim.x, tm = t / p.tiles_n, tn = t - tm * p.tiles_n;
      w.jc = half * 2 + (u & 1);
      w.m0 = tm * kTcBM;
      w.c0 = tn * p.bn + w.jc * 32;
      w.valid = u < n_units && (FULL || (w.jc * 32 < p.bn && w.c0 < p.n));
      return w;
    };
    auto prefetch_bypass = [&](int i, const Unit &w) {
      const DevOp &op = p.ops[i];
      const int c8 = lane & 3;
      const bool ok = FULL || w.c0 + c8 * 8 < p.n;
#pragma unroll
      for (int it = 0; it < 4; it++) {
        int ri = w.m0 + q * 32 + it * 8 + (lane >> 2);
        if (ri >= p.m) ri = p.m - 1;
        long long orow = op.den == op.num ? ri : ((long long)ri * op.num) / op.den;
        if (orow >= op.buf_rows) orow = op.buf_rows - 1;
        const size_t off = (size_t)orow * op.buf_ld + w.c0 + c8 * 8;
        pf_h[it] = make_uint4(0u, 0u, 0u, 0u);
        pf_l[it] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
          pf_h[it] = __ldcs(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(op.buf) + off));
          pf_l[it] = __ldcs(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(op.buf_lo) + off));
        }
      }
    };
    auto prefetch_so = [&](const Unit &w) {
      if constexpr (kStatic && kSoIdx >= 0) {
        so_s = so_o = 0.f;
        if (FULL || w.c0 + lane < p.n) {
          so_s = __ldg(p.ops[kSoIdx].v0 + w.c0 + lane);
          so_o = __ldg(p.ops[kSoIdx].v1 + w.c0 + lane);
        }
      }
    };
    Unit nxt = unit_at(0);
    if (nxt.valid) {
      if (ib >= 0) prefetch_bypass(ib, nxt);
      prefetch_so(nxt);
    }
#pragma unroll 1
    for (int u = 0; u < n_units; u++) {
      {
        const Unit cur = nxt;
        nxt = unit_at(u + 1);
        const int jc = cur.jc, m0 = cur.m0, c0 = cur.c0;
        const uint32_t tcount = (uint32_t)(u >> 1);
        const uint32_t xb = tcount & 1u;
        const uint32_t xaddr = lane_base + (2u + xb) * (uint32_t)p.bn;
        const int r = m0 + q * 32 + lane;
        const int rr = r < p.m ? r : p.m - 1;
        const bool valid = cur.valid, next_valid = nxt.valid;
        if ((u & 1) == 0) {
          timed_wait(xfull_bar(xb), (tcount >> 1) & 1u, prof_a);
          tc_fence_after();
        }
        float v[32];
        if (valid) {
          uint32_t raw[32];
          tmem_ld32_nowait(xaddr + jc * 32, raw);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = __uint_as_float(raw[j]);
        }
        if ((u & 1) == 1) {  // this warp's last chunk of the tile is in registers: the MMA issuer may overwrite this accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(xfree_bar(xb));
        }
        if (!valid) {
          if (next_valid) {
            if (ib >= 0) prefetch_bypass(ib, nxt);
            prefetch_so(nxt);
          }
          continue;
        }
        bool bypass_requested = false;

        auto apply = [&](const int i, const int type) {
          const DevOp &op = p.ops[i];
          switch (type) {
            case EpiOp::kBias:
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = ldvec4<FULL>(op.v0, c0 + j, p.n);
                add2(v[j], v[j + 1], b.x, b.y);
                add2(v[j + 2], v[j + 3], b.z, b.w);
              }
              break;
            case EpiOp::kRelu:
#pragma unroll
              for (int j = 0; j < 32; j++) v[j] = v[j] > 0.f ? v[j] : 0.f;
              break;
            case EpiOp::kScaleOffset:
              // y = x * scale + offset as one fused multiply-add per element (the reference's MulColsVec + AddVecToRows
              // round twice; the fused form is at most half an ulp closer to the exact value)
              if constexpr (kStatic && kSoIdx >= 0) {
                if (i == kSoIdx) {
                  float *vs = reinterpret_cast<float *>(stg);
                  __syncwarp();
                  vs[lane] = so_s;
                  vs[32 + lane] = so_o;
                  __syncwarp();
                  if (next_valid) prefetch_so(nxt);
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    const float4 s = *reinterpret_cast<const float4 *>(vs + j), o = *reinterpret_cast<const float4 *>(vs + 32 + j);
                    fma2(v[j], v[j + 1], s.x, s.y, o.x, o.y);
                    fma2(v[j + 2], v[j + 3], s.z, s.w, o.z, o.w);
                  }
                  break;
                }
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 s = ldvec4<FULL>(op.v0, c0 + j, p.n), o = ldvec4<FULL>(op.v1, c0 + j, p.n);
                fma2(v[j], v[j + 1], s.x, s.y, o.x, o.y);
                fma2(v[j + 2], v[j + 3], s.z, s.w, o.z, o.w);
              }
              break;
            case EpiOp::kScale:
#pragma unroll
              for (int j = 0; j < 32; j += 2) mul2(v[j], v[j + 1], op.alpha, op.alpha);
              break;
            case EpiOp::kAddScaled: {
              // bypass input: read with full-row coalescing, summed hi + lo, transposed through the warp's
              // staging tile so that each thread gets the 32 values of its own row
              __syncwarp();
              if (op.buf_lo) {
                const int c8 = lane & 3;
                if (i != ib) prefetch_bypass(i, cur);  // a second bypass op: loaded now
#pragma unroll
                for (int it = 0; it < 4; it++) {
                  const int ii = it * 8 + (lane >> 2);
                  const __half2 *hh = reinterpret_cast<const __half2 *>(&pf_h[it]), *ll = reinterpret_cast<const __half2 *>(&pf_l[it]);
                  float x[8];
#pragma unroll
                  for (int e = 0; e < 4; e++) {
                    const float2 fh = __half22float2(hh[e]), fl = __half22float2(ll[e]);
                    x[2 * e] = fmaf(fl.x, 1.f / kSplitScale, fh.x);  // exact: hi + lo / 2048
                    x[2 * e + 1] = fmaf(fl.y, 1.f / kSplitScale, fh.y);
                  }
                  stg[ii * 8 + ((2 * c8) ^ (ii & 7))] = make_float4(x[0], x[1], x[2], x[3]);
                  stg[ii * 8 + ((2 * c8 + 1) ^ (ii & 7))] = make_float4(x[4], x[5], x[6], x[7]);
                }
                if (i == ib && next_valid) {  // registers free again: request the next unit's rows
                  prefetch_bypass(ib, nxt);
                  bypass_requested = true;
                }
              } else {
                // plain fp32 source: 8 lanes x 16 B cover a 128-byte row segment, 4 rows per instruction
                const int c4 = lane & 7;
                const bool col_ok = FULL || c0 + c4 * 4 < p.n;
#pragma unroll
                for (int it = 0; it < 8; it++) {
                  const int ii = it * 4 + (lane >> 3);
                  int ri = m0 + q * 32 + ii;
                  if (ri >= p.m) ri = p.m - 1;
                  long long orow = op.den == op.num ? ri : ((long long)ri * op.num) / op.den;
                  if (orow >= op.buf_rows) orow = op.buf_rows - 1;
                  float4 bf = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (col_ok) bf = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(op.buf) + (size_t)orow * op.buf_ld + c0 + c4 * 4));
                  stg[ii * 8 + (c4 ^ (ii & 7))] = bf;
                }
              }
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = stg[lane * 8 + ((j >> 2) ^ (lane & 7))];
                if (op.alpha != 1.f) {
                  mul2(o.x, o.y, op.alpha, op.alpha);
                  mul2(o.z, o.w, op.alpha, op.alpha);
                }
                add2(v[j], v[j + 1], o.x, o.y);
                add2(v[j + 2], v[j + 3], o.z, o.w);
              }
              break;
            }
            case EpiOp::kUttBias: {
              const int u = p.row_utt[(size_t)rr * op.num];
              const float *b = reinterpret_cast<const float *>(op.buf) + (size_t)u * op.buf_ld + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (FULL || c0 + j < p.n) {
                  const float4 o = *reinterpret_cast<const float4 *>(b + j);
                  add2(v[j], v[j + 1], o.x, o.y);
                  add2(v[j + 2], v[j + 3], o.z, o.w);
                }
              break;
            }
          }
        };
        if constexpr (kStatic) {
          if constexpr (kTypes[0] >= 0 && kTypes[0] != EpiOp::kBias) apply(0, kTypes[0]);
          if constexpr (kTypes[1] >= 0) apply(1, kTypes[1]);
          if constexpr (kTypes[2] >= 0) apply(2, kTypes[2]);
          if constexpr (kTypes[3] >= 0) apply(3, kTypes[3]);
        } else {
#pragma unroll 1
          for (int i = first_op; i < p.n_ops; i++) apply(i, p.ops[i].type);
        }
        if (next_valid) {
          if (ib >= 0 && !bypass_requested) prefetch_bypass(ib, nxt);
        }
        // store through the staging tile so that every instruction writes whole row segments
        __syncwarp();
        if (p.out_lo) {
          // two fp16 planes: hi tile in the first 2 KB of the staging tile, lo tile in the second;
          // 16-byte chunks XOR-swizzled by row pair (conflict-free for both access patterns)
          uint4 *st16 = reinterpret_cast<uint4 *>(stg);
          float amax = 0.f;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            uint4 hh, ll;
            split2x(v[8 * c + 0], v[8 * c + 1], hh.x, ll.x, amax);
            split2x(v[8 * c + 2], v[8 * c + 3], hh.y, ll.y, amax);
            split2x(v[8 * c + 4], v[8 * c + 5], hh.z, ll.z, amax);
            split2x(v[8 * c + 6], v[8 * c + 7], hh.w, ll.w, amax);
            const int slot = lane * 4 + (c ^ ((lane >> 1) & 3));
            st16[slot] = hh;
            st16[128 + slot] = ll;
          }
          if (amax > 65504.f && r < p.m) *p.range_flag = 1;
          __syncwarp();
          const int c8 = lane & 3;
#pragma unroll
          for (int it = 0; it < 4; it++) {
            const int i = it * 8 + (lane >> 2);
            const int ri = m0 + q * 32 + i;
            if (ri < p.m && (FULL || c0 + c8 * 8 < p.n)) {
              const int slot = i * 4 + (c8 ^ ((i >> 1) & 3));
              const size_t off = (size_t)ri * p.out_ld + c0 + c8 * 8;
              *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(p.out_hi) + off) = st16[slot];
              *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(p.out_lo) + off) = st16[128 + slot];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) stg[lane * 8 + ((j >> 2) ^ (lane & 7))] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; it++) {
            const int i = it * 4 + (lane >> 3), c4 = lane & 7;
            const int ri = m0 + q * 32 + i;
            if (ri < p.m && (FULL || c0 + c4 * 4 < p.n))
              *reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out_hi) + (size_t)ri * p.out_ld + c0 + c4 * 4) = stg[i * 8 + (c4 ^ (i & 7))];
          }
        }
        __syncwarp();
      }
    }
    if (PROF && blockIdx.x == 0 && lane == 0 && (warp == 8 || warp == 12))
      printf("tc2 profile: tail warp %d total %lld clk, waiting for a finished tile %lld\n", warp, clock64() - prof_t0, prof_a);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
namespace {

int Pattern(const TcParams &p) {
  if (p.n_ops > 4) return -1;
  int pat = 0;
  for (int i = 0; i < p.n_ops; i++) pat |= (p.ops[i].type + 1) << (4 * i);
  return pat;
}
constexpr int PatOf(int a = -1, int b = -1, int c = -1, int d = -1) { return (a + 1) | ((b + 1) << 4) | ((c + 1) << 8) | ((d + 1) << 12); }
constexpr int kPatNone = PatOf();
constexpr int kPatBias = PatOf(EpiOp::kBias);
constexpr int kPatBRS = PatOf(EpiOp::kBias, EpiOp::kRelu, EpiOp::kScaleOffset);
constexpr int kPatBRSA = PatOf(EpiOp::kBias, EpiOp::kRelu, EpiOp::kScaleOffset, EpiOp::kAddScaled);

template <int PAT, bool FULL, bool PROF = false>
void Launch(const TcParams &p, int grid, int smem, int smem_limit, cudaStream_t stream) {
  static int configured_dev = -1;  // opt-in shared memory size is a per-device function attribute
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<PAT, FULL, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit);
    if (e != cudaSuccess) RS_FAIL("cudaFuncSetAttribute(gemm_tc2_kernel): " << cudaGetErrorString(e));
    configured_dev = dev;
  }
  gemm_tc2_kernel<PAT, FULL, PROF><<<grid, kT2Threads, smem, stream>>>(p);
}

}  // namespace

void LaunchGemmTc2(const TcParams &p, int num_sms, int smem_limit, cudaStream_t stream) {
  const int stage_bytes = 2 * kT2StageA + 2 * p.bn * 128;
  const int smem = 1024 + p.stages * stage_bytes + 8 * 4096 + 8 * (2 * p.stages + 10) + 1024;
  int grid = p.tiles_m * p.tiles_n;
  if (grid > num_sms) grid = num_sms;
  const bool full = p.bn == 128 && p.n % 128 == 0;
  const int pat = Pattern(p);
  if (p.profile && full && pat == kPatNone) Launch<kPatNone, true, true>(p, grid, smem, smem_limit, stream);
  else if (p.profile && full && pat == kPatBRSA) Launch<kPatBRSA, true, true>(p, grid, smem, smem_limit, stream);
  else if (full && pat == kPatNone) Launch<kPatNone, true>(p, grid, smem, smem_limit, stream);
  else if (full && pat == kPatBRS) Launch<kPatBRS, true>(p, grid, smem, smem_limit, stream);
  else if (full && pat == kPatBRSA) Launch<kPatBRSA, true>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatNone) Launch<kPatNone, false>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatBias) Launch<kPatBias, false>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatBRS) Launch<kPatBRS, false>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatBRSA) Launch<kPatBRSA, false>(p, grid, smem, smem_limit, stream);
  else Launch<-1, false>(p, grid, smem, smem_limit, stream);
}

}  // namespace rs
