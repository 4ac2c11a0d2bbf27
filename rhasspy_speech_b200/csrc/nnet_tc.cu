// Stage (ii) on the 5th-generation tensor cores: the affine layers of the TDNN(-F) forward
// (TdnnComponent::Propagate, kaldi/src/nnet3/nnet-tdnn-component.cc:181-211, and the Affine /
// Linear components of nnet3/nnet-simple-component.cc, which the reference runs as cblas_sgemm,
// kaldi/src/matrix/kaldi-matrix.cc:171-183) as ONE persistent, warp-specialised kernel per layer:
//
//   warp 0      TMA producer : cp.async.bulk.tensor tiles of the activations (one tensor map per
//                              time-offset slab: the TDNN splice is a row-shifted / row-strided view
//                              of the producing layer's buffer, never materialised) and of the weights
//   warp 1      MMA issuer   : tcgen05.mma.kind::f16, per-K-block partial sums in TMEM
//   warps 2..9  epilogue     : tcgen05.ld of every K-block partial sum -> fp32 running sums in
//                              registers -> bias / ReLU / BatchNorm scale+offset / bypass add in the
//                              reference's order -> split store
//
// Numerics.  The reference computes in fp32; the tolerance on the log-likelihoods is 1e-4.  One
// fp16 (or TF32) product carries an 11-bit significand, ~1e-3.  So every operand is carried as two
// fp16 planes (split.cuh)
//   x ~ hi + lo / 2048,   hi = fp16(x),   lo = fp16((x - hi) * 2048)          (22+ significant bits)
// and a product is three MMAs:  hi*hi  into a "main" accumulator and  hi*lo + lo*hi  into a "cross"
// accumulator that is folded with the exact factor 2^-11.  The dropped lo*lo term and the rounding
// of lo are O(2^-22) relative per product and unbiased.  The scaling keeps lo in fp16's normal range
// whatever the magnitude of x (an unscaled remainder of a weight of 0.03 would be subnormal).
// fp16 rather than TF32 planes: the kernel is bound by the bytes each SM can pull from L2 per MMA
// (measured with the TF32 variant: 35 B/clk/SM, tensor pipe 30 % busy), and fp16 halves the bytes
// per product and doubles the MMA rate.  Activations are stored by the producing epilogue already
// split, weights are split once at model load; values beyond +-65504 saturate and raise a flag
// that fails the call (the fp32 CUDA-core path, RS_B200_GEMM=simt, has no such limit).
#include <cuda.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "engine.h"
#include "model.h"
#include "nnet_tc.h"
#include "split.cuh"
#include "tc_ptx.cuh"

namespace rs {

// ------------------------------------------------------------------------------------ the kernel
// Accumulation.  The tensor core adds each MMA into the fp32 TMEM accumulator with truncation, not
// round-to-nearest (measured with whole-K accumulation in TMEM: the error of a K = 2048 dot product
// grows linearly with K and is biased towards zero, ~1e-5 relative, which breaks the 1e-4 gate after
// 30 layers).  So TMEM only ever holds the partial sums of ONE 64-wide K block: per block the issuer
// starts a fresh main accumulator (4 MMAs) and hands it to the epilogue warps, which add it into
// fp32 registers with round-to-nearest -- the blocked summation a CPU sgemm micro-kernel performs.
// The cross terms (8 MMAs per block) are 2^-11 of the result, so their accumulator stays in TMEM
// for the whole tile and is folded once (its truncation error is 2^-11 * 1e-5: nothing).  TMEM holds
// a ring of two main accumulators and two cross accumulators (4 x bn columns), so the issuer runs
// up to two K blocks ahead, also across the tile boundary while the epilogue warps run the bias /
// ReLU / BatchNorm / bypass / split-store tail of the previous tile.
constexpr int kStageABytes = kTcBM * 128;  // one plane of the activation tile: 128 rows x 128 B
// main MMAs (K = 16 each) accumulated inside TMEM before the epilogue warps fold the sum in registers: 4 = one fold per
// 64-wide K block, 1 = every MMA starts from zero (no in-TMEM accumulation of the main term at all)
// Measured on the bench model (scripts/debug_ll.py, log-likelihoods against an fp64 forward; the reference's own
// nnet3-compute is 8e-6 rms / 7e-5 max away from it): 4 -> 1.5e-5 rms / 1.4e-4 max (the truncation inside TMEM is biased
// towards zero and the bias adds up coherently over the layers), 2 -> 9e-6 / 7e-5, 1 -> 7e-6 / 5e-5.
constexpr int kTcFold = 2;

// 12 warps = 3 per SM sub-partition (16 K registers each): 168 registers per thread at launch, then
// re-allocated by setmaxnreg to 40 (TMA / MMA warpgroup) and 232 (epilogue warpgroups)
// PAT >= 0: the epilogue op sequence is a compile-time constant (4 bits per op: EpiOp::Type + 1,
// first op in the low bits, see TcPattern); PAT < 0: run-time op list.
template <int PAT>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr bool kStatic = PAT >= 0;
  constexpr int kT0 = kStatic ? ((PAT >> 0) & 15) - 1 : -1, kT1 = kStatic ? ((PAT >> 4) & 15) - 1 : -1;
  constexpr int kT2 = kStatic ? ((PAT >> 8) & 15) - 1 : -1, kT3 = kStatic ? ((PAT >> 12) & 15) - 1 : -1;
  constexpr bool kPairLoads = PAT == 0;  // no tail ops (the 2048 -> 128 bottleneck layers)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.bn * 128u;
  const uint32_t stage_bytes = 2u * kStageABytes + 2u * b_bytes;
  // [stages][8 x 4 KB epilogue staging tiles][barriers]
  const uint32_t epi0 = smem0 + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar0 = epi0 + 8u * 4096u;
  // barriers: full[stages] | empty[stages] | set_full[4] | set_empty[4] | tmem slot
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(p.stages + s); };
  auto setf_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * p.stages + a); };
  // main-accumulator barriers are indexed (epilogue group, physical set) = g * 2 + s: an mbarrier waiter
  // may be at most one phase behind, so the two groups, which alternate tiles, cannot share a barrier
  auto sete_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * p.stages + 4 + a); };
  auto crosse_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * p.stages + 8 + a); };  // cross accumulator a drained
  const uint32_t tmem_slot = bar0 + 8u * (uint32_t)(2 * p.stages + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 4; a++) {
      mbar_init(setf_bar(a), 1);
      mbar_init(sete_bar(a), 4);  // released by the 4 warps of one group
    }
    mbar_init(crosse_bar(0), 4);
    mbar_init(crosse_bar(1), 4);
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

  // register re-allocation between the warpgroups: 40 for {TMA, MMA, 2 idle warps}, 232 for the epilogue groups
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      long long prof_wait = 0, prof_t0 = clock64();
      bool uniform = true;
      for (int s = 1; s < p.n_slabs; s++) uniform = uniform && p.slabs[s].kblocks == p.slabs[0].kblocks;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * kTcBM, n0 = (tile % p.tiles_n) * p.bn;
        // K blocks are visited block-major across the slabs when the slabs are equally long: the
        // time-offset slabs of a TDNN layer read the same source rows shifted by a few rows, so
        // back-to-back loads hit in L2 (slab-major order re-read the whole source from HBM)
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
          long long tw0 = p.profile ? clock64() : 0;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (p.profile) prof_wait += clock64() - tw0;
          const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes, fb = full_bar(stage);
          mbar_expect_tx(fb, stage_bytes);
          tma_load_2d(sa, &p.a_hi[s], kb * kTcBK, m0 + sl.yshift, fb);
          tma_load_2d(sa + kStageABytes, &p.a_lo[s], kb * kTcBK, m0 + sl.yshift, fb);
          tma_load_2d(sa + 2 * kStageABytes, &p.w_hi, sl.wk0 + kb * kTcBK, n0, fb);
          tma_load_2d(sa + 2 * kStageABytes + b_bytes, &p.w_lo, sl.wk0 + kb * kTcBK, n0, fb);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (p.profile && blockIdx.x == 0)
        printf("gemm_tc profile (n=%d k-blocks/tile=%d tiles=%d): TMA thread total %lld clk, waiting for a free stage %lld\n", p.n, total_kb,
               num_tiles, clock64() - prof_t0, prof_wait);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      // (Measured, RS_B200_TC_PROFILE: this thread spends 60-75 % of the kernel outside any wait.  Walking the loop with
      // the whole warp and issuing from an elected lane removes the ELECT + R2UR sequences in front of every UTCHMMA but
      // not the time: the MMA issue itself blocks -- three MMAs per K step read 24 KB of operands from shared memory
      // next to 16 KB of TMA writes, and the tile is bound by that traffic, not by the tensor pipe.)
      constexpr bool leader = true;
      // instruction descriptor: D fp32, A/B fp16, both K-major, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, kbc = 0;  // kbc: K blocks issued so far; physical main set = kbc & 1
      uint32_t use_par = 0;         // bit b: parity of the number of blocks committed on barrier pair b = group * 2 + set
      int last_b[2] = {-1, -1};     // barrier pair of the block that last occupied each physical set
      long long prof_full = 0, prof_sete = 0, prof_cross = 0, prof_t0 = clock64();
      constexpr int fold = kTcFold;
      uint32_t tcount = 0;  // tiles issued so far: cross accumulator = tcount & 1
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
        const uint32_t d_cross = tmem_base + (uint32_t)((2 + (tcount & 1)) * p.bn);
        {
          const long long tw0 = p.profile ? clock64() : 0;
          mbar_wait(crosse_bar(tcount & 1), ((tcount >> 1) & 1u) ^ 1u);
          if (p.profile) prof_cross += clock64() - tw0;
        }
        for (int kb = 0; kb < total_kb; kb++) {
          {
            const long long tw0 = p.profile ? clock64() : 0;
            mbar_wait(full_bar(stage), phase);
            if (p.profile) prof_full += clock64() - tw0;
          }
          tc_fence_after();
          const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes;
          const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sa + kStageABytes);
          const uint64_t b_hi = smem_desc_sw128(sa + 2 * kStageABytes), b_lo = smem_desc_sw128(sa + 2 * kStageABytes + b_bytes);
          // fold main MMAs (16 K each) share one fresh accumulator; with fold = 1 every MMA starts from zero and the
          // epilogue warps do ALL the accumulation in fp32 registers with round-to-nearest
#pragma unroll
          for (int k = 0; k < kTcBK / 16; k++) {  // 16 fp16 = 32 bytes per step inside the swizzle atom
            const int set = kbc & 1, bsel = (int)(tcount & 1) * 2 + set;
            if (k % fold == 0 && last_b[set] >= 0) {
              const long long tw0 = p.profile ? clock64() : 0;
              mbar_wait(sete_bar(last_b[set]), ((use_par >> last_b[set]) & 1u) ^ 1u);  // set drained
              if (p.profile) prof_sete += clock64() - tw0;
            }
            const uint32_t d_main = tmem_base + (uint32_t)(set * p.bn);
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            if (leader) {
              tc_mma_f16(d_main, a_hi + adv, b_hi + adv, idesc, k % fold != 0 ? 1u : 0u);
              tc_mma_f16(d_cross, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0 ? 1u : 0u);
              tc_mma_f16(d_cross, a_hi + adv, b_lo + adv, idesc, 1u);
              if (k == kTcBK / 16 - 1) tc_commit(empty_bar(stage));  // frees the smem stage when these MMAs have read it
            }
            if (k % fold == fold - 1) {
              if (leader) tc_commit(setf_bar(bsel));  // this partial sum (and, on the last one, the cross sum) complete
              use_par ^= 1u << bsel;
              last_b[set] = bsel;
              kbc++;
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (p.profile && blockIdx.x == 0 && leader)
        printf("gemm_tc profile: MMA thread total %lld clk, waiting for operands %lld, for a drained accumulator %lld, for the cross accumulator %lld\n",
               clock64() - prof_t0, prof_full, prof_sete, prof_cross);
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ------------------------------------------------- epilogue groups: warps 4..7 and 8..11
    // The groups alternate tiles, so the tail of tile i (bias .. split store) overlaps the main loop and
    // the folds of tile i+1.  warp -> TMEM lane quadrant q; a thread owns one output row and all (<= 128)
    // columns of the tile: 128 fp32 running sums in registers (the warpgroup holds 232 registers per
    // thread after setmaxnreg, the TMA / MMA warpgroup 40).
    const int q = warp & 3, group = (warp - 4) >> 2;
    constexpr int h = 0;
    // this warp's 32 x 32 fp32 staging tile (float4 columns XOR-swizzled by row: conflict-free both ways)
    float4 *stg = reinterpret_cast<float4 *>(smem_raw + (epi0 - smem_u32(smem_raw)) + (uint32_t)(warp - 4) * 4096u);
    uint32_t tcount = group;  // local index of this group's current tile: K-block ring position = tcount * total_kb
    long long prof_setf = 0, prof_fold = 0, prof_tail = 0, prof_t0 = clock64();
    uint32_t cnt_par = 0;     // bit s: parity of the number of blocks this group has taken from physical set s
    int ib = -1;  // the op whose split bypass input is prefetched (first kAddScaled with a split source)
    for (int i = 0; i < p.n_ops && ib < 0; i++)
      if (p.ops[i].type == EpiOp::kAddScaled && p.ops[i].buf_lo) ib = i;
    for (int tile = blockIdx.x + group * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, tcount += 2) {
      const int m0 = (tile / p.tiles_n) * kTcBM, n0 = (tile % p.tiles_n) * p.bn;
      const int total_sums = total_kb * (kTcBK / 16 / kTcFold);  // partial sums the issuer publishes per tile
      uint32_t kbc = tcount * (uint32_t)total_sums;
      // bypass input of one 32-column chunk: 32 columns = 64 bytes per plane and row; 4 lanes x 16 B
      // cover a row segment, 8 rows per instruction; issued early so that HBM latency is hidden
      uint4 pf_h[4], pf_l[4];
      auto prefetch = [&](int jc) {
        const DevOp &op = p.ops[ib];
        const int c0 = n0 + h * 64 + jc * 32, c8 = lane & 3;
        const bool ok = h * 64 + jc * 32 < p.bn && c0 + c8 * 8 < p.n;
#pragma unroll
        for (int it = 0; it < 4; it++) {
          int ri = m0 + q * 32 + it * 8 + (lane >> 2);
          if (ri >= p.m) ri = p.m - 1;
          long long orow = op.den == op.num ? ri : ((long long)ri * op.num) / op.den;
          if (orow >= op.buf_rows) orow = op.buf_rows - 1;
          const size_t off = (size_t)orow * op.buf_ld + c0 + c8 * 8;
          pf_h[it] = make_uint4(0u, 0u, 0u, 0u);
          pf_l[it] = make_uint4(0u, 0u, 0u, 0u);
          if (ok) {
            pf_h[it] = __ldcs(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(op.buf) + off));
            pf_l[it] = __ldcs(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(op.buf_lo) + off));
          }
        }
      };
      // static op list: lane l keeps column l of each per-column vector (bias, BatchNorm scale /
      // offset) of both chunks in a register, loaded here so that the latency hides behind the main
      // loop; the tail broadcasts a column with a shuffle (a warp works on one column set for 32 rows)
      float vr0[4][kTcChunks], vr1[4][kTcChunks];
      if constexpr (kStatic) {
        constexpr int types[4] = {kT0, kT1, kT2, kT3};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int jc = 0; jc < kTcChunks; jc++) {
            const int c = n0 + h * 64 + jc * 32 + lane;
            const bool ok = h * 64 + jc * 32 + lane < p.bn && c < p.n;
            vr0[i][jc] = vr1[i][jc] = 0.f;
            if (types[i] == EpiOp::kBias || types[i] == EpiOp::kScaleOffset) vr0[i][jc] = ok ? __ldg(p.ops[i].v0 + c) : 0.f;
            if (types[i] == EpiOp::kScaleOffset) vr1[i][jc] = ok ? __ldg(p.ops[i].v1 + c) : 0.f;
          }
      }
      // AffineComponent / TdnnComponent::Propagate copy the bias into the output and let the GEMM accumulate
      // onto it (nnet-simple-component.cc, nnet-tdnn-component.cc:181-211): same order here when the op list
      // starts with the bias
      constexpr bool kBiasFirst = kStatic && kT0 == EpiOp::kBias;
      float acc[kTcMaxBN];
#pragma unroll
      for (int j = 0; j < kTcMaxBN; j++) acc[j] = kBiasFirst ? __shfl_sync(0xffffffffu, vr0[0][j >> 5], j & 31) : 0.f;
      for (int kb = 0; kb < total_sums; kb++, kbc++) {
        if (kb == total_sums - 1 && ib >= 0) prefetch(0);
        const int set = kbc & 1, bsel = group * 2 + set;
        long long tw0 = p.profile ? clock64() : 0;
        mbar_wait(setf_bar(bsel), (cnt_par >> set) & 1u);
        if (p.profile) {
          const long long now = clock64();
          prof_setf += now - tw0;
          tw0 = now;
        }
        cnt_par ^= 1u << set;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * p.bn + h * 64);
        if constexpr (kPairLoads) {
          // two TMEM loads in flight per wait (the load -> wait round trip, not the adds, bounds the fold); only where
          // the tail leaves the registers for 64 raw values (measured with spills on the BatchNorm / bypass patterns)
#pragma unroll
          for (int jc = 0; jc < kTcChunks; jc += 2)
            if (h * 64 + jc * 32 < p.bn) {
              uint32_t raw0[32], raw1[32];
              const bool two = h * 64 + (jc + 1) * 32 < p.bn;
              tmem_ld32_nowait(taddr + jc * 32, raw0);
              if (two) tmem_ld32_nowait(taddr + (jc + 1) * 32, raw1);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; j += 2)
                add2(acc[jc * 32 + j], acc[jc * 32 + j + 1], __uint_as_float(raw0[j]), __uint_as_float(raw0[j + 1]));
              if (two) {
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                  add2(acc[(jc + 1) * 32 + j], acc[(jc + 1) * 32 + j + 1], __uint_as_float(raw1[j]), __uint_as_float(raw1[j + 1]));
              }
            }
        } else {
#pragma unroll
          for (int jc = 0; jc < kTcChunks; jc++)
            if (h * 64 + jc * 32 < p.bn) {
              uint32_t raw[32];
              tmem_ld32_nowait(taddr + jc * 32, raw);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; j += 2)
                add2(acc[jc * 32 + j], acc[jc * 32 + j + 1], __uint_as_float(raw[j]), __uint_as_float(raw[j + 1]));
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sete_bar(bsel));
        if (p.profile) prof_fold += clock64() - tw0;
      }
      const long long tail0 = p.profile ? clock64() : 0;
      {  // the tile's cross sum: acc += cross * 2^-11 (the last block's commit covers it)
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((2 + (tcount & 1)) * p.bn + h * 64);
#pragma unroll
        for (int jc = 0; jc < kTcChunks; jc++)
          if (h * 64 + jc * 32 < p.bn) {
            uint32_t raw[32];
            tmem_ld32_nowait(taddr + jc * 32, raw);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j++) acc[jc * 32 + j] = fmaf(__uint_as_float(raw[j]), 1.f / kSplitScale, acc[jc * 32 + j]);
          }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(crosse_bar(tcount & 1));
      }
      const int r = m0 + q * 32 + lane;
      const int rr = r < p.m ? r : p.m - 1;
#pragma unroll
      for (int jc = 0; jc < kTcChunks; jc++) {
        const int c0 = n0 + h * 64 + jc * 32;
        if (h * 64 + jc * 32 >= p.bn || c0 >= p.n) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = acc[jc * 32 + j];
        auto apply = [&](const int i, const int type, const float vb0, const float vb1) {
          const DevOp &op = p.ops[i];
          switch (type) {
            case EpiOp::kBias:
              if constexpr (kStatic) {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = __fadd_rn(v[j], __shfl_sync(0xffffffffu, vb0, j));
                break;
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (c0 + j < p.n) {
                  const float4 b = __ldg(reinterpret_cast<const float4 *>(op.v0 + c0 + j));
                  v[j] = __fadd_rn(v[j], b.x);
                  v[j + 1] = __fadd_rn(v[j + 1], b.y);
                  v[j + 2] = __fadd_rn(v[j + 2], b.z);
                  v[j + 3] = __fadd_rn(v[j + 3], b.w);
                }
              break;
            case EpiOp::kRelu:
#pragma unroll
              for (int j = 0; j < 32; j++) v[j] = v[j] > 0.f ? v[j] : 0.f;
              break;
            case EpiOp::kScaleOffset:
              if constexpr (kStatic) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {  // y = x * scale, then + offset: two roundings as the reference, two columns per instruction
                  mul2(v[j], v[j + 1], __shfl_sync(0xffffffffu, vb0, j), __shfl_sync(0xffffffffu, vb0, j + 1));
                  add2(v[j], v[j + 1], __shfl_sync(0xffffffffu, vb1, j), __shfl_sync(0xffffffffu, vb1, j + 1));
                }
                break;
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (c0 + j < p.n) {
                  const float4 s = __ldg(reinterpret_cast<const float4 *>(op.v0 + c0 + j));
                  const float4 o = __ldg(reinterpret_cast<const float4 *>(op.v1 + c0 + j));
                  v[j] = __fadd_rn(__fmul_rn(v[j], s.x), o.x);
                  v[j + 1] = __fadd_rn(__fmul_rn(v[j + 1], s.y), o.y);
                  v[j + 2] = __fadd_rn(__fmul_rn(v[j + 2], s.z), o.z);
                  v[j + 3] = __fadd_rn(__fmul_rn(v[j + 3], s.w), o.w);
                }
              break;
            case EpiOp::kScale:
#pragma unroll
              for (int j = 0; j < 32; j++) v[j] = __fmul_rn(v[j], op.alpha);
              break;
            case EpiOp::kAddScaled: {
              // bypass input: read with full-row coalescing (8 lanes x 16 B cover a 128-byte row
              // segment, 4 rows per instruction), summed hi + lo, transposed through the warp's
              // staging tile so that each thread gets the 32 values of its own row
              __syncwarp();
              if (op.buf_lo) {
                // split source: registers filled by prefetch() (or loaded now for a second bypass op)
                const int c8 = lane & 3;
                if (i != ib) {
                  const int keep = ib;
                  ib = i;
                  prefetch(jc);
                  ib = keep;
                }
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
                if (i == ib && jc + 1 < kTcChunks) prefetch(jc + 1);  // next chunk's bypass while this one is finished
              } else {
                // plain fp32 source: 8 lanes x 16 B cover a 128-byte row segment, 4 rows per instruction
                float4 bf[8];
                const int c4 = lane & 7;
                const bool col_ok = c0 + c4 * 4 < p.n;
#pragma unroll
                for (int it = 0; it < 8; it++) {
                  int ri = m0 + q * 32 + it * 4 + (lane >> 3);
                  if (ri >= p.m) ri = p.m - 1;
                  long long orow = op.den == op.num ? ri : ((long long)ri * op.num) / op.den;
                  if (orow >= op.buf_rows) orow = op.buf_rows - 1;
                  bf[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (col_ok)
                    bf[it] = __ldcs(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(op.buf) + (size_t)orow * op.buf_ld + c0 + c4 * 4));
                }
#pragma unroll
                for (int it = 0; it < 8; it++) {
                  const int i = it * 4 + (lane >> 3);
                  stg[i * 8 + (c4 ^ (i & 7))] = bf[it];
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
                if (c0 + j < p.n) {
                  const float4 o = *reinterpret_cast<const float4 *>(b + j);
                  v[j] = __fadd_rn(v[j], o.x);
                  v[j + 1] = __fadd_rn(v[j + 1], o.y);
                  v[j + 2] = __fadd_rn(v[j + 2], o.z);
                  v[j + 3] = __fadd_rn(v[j + 3], o.w);
                }
              break;
            }
          }
        };
        if constexpr (kStatic) {
          if constexpr (kT0 >= 0 && !kBiasFirst) apply(0, kT0, vr0[0][jc], vr1[0][jc]);
          if constexpr (kT1 >= 0) apply(1, kT1, vr0[1][jc], vr1[1][jc]);
          if constexpr (kT2 >= 0) apply(2, kT2, vr0[2][jc], vr1[2][jc]);
          if constexpr (kT3 >= 0) apply(3, kT3, vr0[3][jc], vr1[3][jc]);
        } else {
#pragma unroll 1
          for (int i = 0; i < p.n_ops; i++) apply(i, p.ops[i].type, 0.f, 0.f);
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
            split2(v[8 * c + 0], v[8 * c + 1], hh.x, ll.x, amax);
            split2(v[8 * c + 2], v[8 * c + 3], hh.y, ll.y, amax);
            split2(v[8 * c + 4], v[8 * c + 5], hh.z, ll.z, amax);
            split2(v[8 * c + 6], v[8 * c + 7], hh.w, ll.w, amax);
            const int slot = lane * 4 + (c ^ ((lane >> 1) & 3));
            st16[slot] = hh;
            st16[128 + slot] = ll;
          }
          const bool sat = amax > 65504.f;
          if (sat && r < p.m) *p.range_flag = 1;
          __syncwarp();
          const int c8 = lane & 3;
#pragma unroll
          for (int it = 0; it < 4; it++) {
            const int i = it * 8 + (lane >> 2);
            const int ri = m0 + q * 32 + i;
            if (ri < p.m && c0 + c8 * 8 < p.n) {
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
            if (ri < p.m && c0 + c4 * 4 < p.n)
              *reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out_hi) + (size_t)ri * p.out_ld + c0 + c4 * 4) = stg[i * 8 + (c4 ^ (i & 7))];
          }
        }
        __syncwarp();
      }
      if (p.profile) prof_tail += clock64() - tail0;
    }
    if (p.profile && blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 8))
      printf("gemm_tc profile: epilogue warp %d total %lld clk, waiting for partial sums %lld, folding %lld, cross fold + tail %lld\n", warp,
             clock64() - prof_t0, prof_setf, prof_fold, prof_tail);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn GetEncodeTiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) RS_FAIL("cuTensorMapEncodeTiled is not available from the CUDA driver");
  return fn;
}

}  // namespace

// 2-D fp16 tensor [rows x cols], row pitch `pitch_elems`, box = 64 columns x box_rows, 128-byte
// swizzle, out-of-bounds elements read as zero.
void TcEncodeMap(CUtensorMap *map, const __half *base, long long rows, int cols, long long pitch_elems, int box_rows) {
  if (rows < 1) rows = 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * 2) % 16) RS_FAIL("tensor map operand is not 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = GetEncodeTiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    RS_FAIL("cuTensorMapEncodeTiled failed (" << (int)rc << ") rows " << rows << " cols " << cols << " pitch " << pitch_elems);
}

bool TcSplitHost(float x, __half *hi, __half *lo) {
  bool ok = true;
  if (std::fabs(x) > 65504.f) {
    x = std::copysign(65504.f, x);
    ok = false;
  }
  *hi = __float2half_rn(x);
  *lo = __float2half_rn((x - __half2float(*hi)) * kSplitScale);
  return ok;
}

int TcTileN(int n) {
  const int tiles = (n + kTcMaxBN - 1) / kTcMaxBN;  // fewest tiles, then the smallest tile that covers n
  const int per = (n + tiles - 1) / tiles;
  return (per + 31) / 32 * 32;
}

void TcPackWeights(const float *w, int n, int ktot, const std::vector<std::pair<int, int>> &slabs, std::vector<__half> *hi,
                   std::vector<__half> *lo, std::vector<int> *k0, int *kp) {
  int total = 0;
  k0->clear();
  for (const auto &s : slabs) {
    k0->push_back(total);
    total += (s.second + kTcBK - 1) / kTcBK * kTcBK;
  }
  *kp = total;
  hi->assign((size_t)n * total, __float2half_rn(0.f));
  lo->assign((size_t)n * total, __float2half_rn(0.f));
  for (size_t si = 0; si < slabs.size(); si++) {
    const int wcol = slabs[si].first, k = slabs[si].second;
    if (wcol + k > ktot) RS_FAIL("weight slab out of range");
    for (int r = 0; r < n; r++)
      for (int c = 0; c < k; c++)
        if (!TcSplitHost(w[(size_t)r * ktot + wcol + c], &(*hi)[(size_t)r * total + (*k0)[si] + c], &(*lo)[(size_t)r * total + (*k0)[si] + c]))
          RS_FAIL("a weight exceeds the fp16 range of the tensor-core path (set RS_B200_GEMM=simt)");
  }
}

static int g_tc_smem_limit = 0;

void TcConfigure(TcParams *p) {
  if (g_tc_smem_limit == 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_tc_smem_limit = v;
  }
  const int stage_bytes = 2 * kStageABytes + 2 * p->bn * 128;
  int stages = (g_tc_smem_limit - 1024 - 8 * 4096 - 256) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) RS_FAIL("not enough shared memory for the tensor-core GEMM pipeline");
  p->stages = stages;
  int cols = 32;
  while (cols < 4 * p->bn) cols <<= 1;
  p->tmem_cols = cols;
  static const int prof = getenv("RS_B200_TC_PROFILE") ? atoi(getenv("RS_B200_TC_PROFILE")) : 0;
  p->profile = prof;
  static const int fold_env = getenv("RS_B200_TC_FOLD") ? atoi(getenv("RS_B200_TC_FOLD")) : 0;
  p->fold = fold_env == 1 || fold_env == 2 || fold_env == 4 ? fold_env : kTcFold;
  // short contractions (the 128 -> 1024 layers: K = 256) may use a different fold: few partial sums per output
  static const int fold_short = getenv("RS_B200_TC_FOLD_SHORT") ? atoi(getenv("RS_B200_TC_FOLD_SHORT")) : 0;
  int total_kb = 0;
  for (int s = 0; s < p->n_slabs; s++) total_kb += p->slabs[s].kblocks;
  if (total_kb <= 8 && (fold_short == 1 || fold_short == 2 || fold_short == 4)) p->fold = fold_short;
  p->tiles_m = (p->m + kTcBM - 1) / kTcBM;
  p->tiles_n = (p->n + p->bn - 1) / p->bn;
}

// op sequence -> template pattern (4 bits per op, type + 1); -1 if it has no static instantiation
static int TcPattern(const TcParams &p) {
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
constexpr int kPatS = PatOf(EpiOp::kScaleOffset);
constexpr int kPatU = PatOf(EpiOp::kUttBias);

template <int PAT>
static void LaunchPattern(const TcParams &p, int grid, int smem, cudaStream_t stream) {
  static int configured_dev = -1;  // opt-in shared memory size is a per-device function attribute
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_tc_smem_limit);
    if (e != cudaSuccess) RS_FAIL("cudaFuncSetAttribute(gemm_tc_kernel): " << cudaGetErrorString(e));
    configured_dev = dev;
  }
  gemm_tc_kernel<PAT><<<grid, kTcThreads, smem, stream>>>(p);
}

void LaunchGemmTc(const TcParams &p, int num_sms, cudaStream_t stream) {
  if (p.m <= 0 || p.n <= 0) return;
  // RS_B200_TC = v1 | v2 keep the earlier arrangements of the kernel selectable for A/B measurements
  static const int ver = !getenv("RS_B200_TC") ? 3 : !strcmp(getenv("RS_B200_TC"), "v1") ? 1 : !strcmp(getenv("RS_B200_TC"), "v2") ? 2 : 3;
  if (ver == 3) {
    LaunchGemmTc3(p, num_sms, g_tc_smem_limit, stream);
    return;
  }
  if (ver == 2) {
    LaunchGemmTc2(p, num_sms, g_tc_smem_limit, stream);
    return;
  }
  const int stage_bytes = 2 * kStageABytes + 2 * p.bn * 128;
  const int smem = 1024 + p.stages * stage_bytes + 8 * 4096 + 8 * (2 * p.stages + 10) + 16;
  int grid = p.tiles_m * p.tiles_n;
  if (grid > num_sms) grid = num_sms;
  switch (TcPattern(p)) {
    case kPatNone: LaunchPattern<kPatNone>(p, grid, smem, stream); break;
    case kPatBias: LaunchPattern<kPatBias>(p, grid, smem, stream); break;
    case kPatBRS: LaunchPattern<kPatBRS>(p, grid, smem, stream); break;
    case kPatBRSA: LaunchPattern<kPatBRSA>(p, grid, smem, stream); break;
    case kPatS: LaunchPattern<kPatS>(p, grid, smem, stream); break;
    case kPatU: LaunchPattern<kPatU>(p, grid, smem, stream); break;
    default: LaunchPattern<-1>(p, grid, smem, stream); break;
  }
}

}  // namespace rs
