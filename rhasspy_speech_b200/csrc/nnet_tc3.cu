// Stage (ii), third arrangement of the tensor-core affine kernel (TdnnComponent::Propagate,
// kaldi/src/nnet3/nnet-tdnn-component.cc:181-211; Affine / Linear components of nnet3/nnet-simple-component.cc).
// Arithmetic as in nnet_tc.cu (three fp16 MMAs per product; the main partial sums leave TMEM every `fold` MMAs and
// are summed in fp32 registers with round-to-nearest; the cross sum is folded once with the exact factor 2^-11).
//
//   warp 0        TMA producer (one lane)
//   warp 1        MMA issuer   (one lane)
//   warps 4..19   sixteen EPILOGUE warps = 4 TMEM lane quadrants x 4 column chunks: a thread owns one output row
//                 and 32 columns of the tile.  Per published partial sum a warp issues ONE tcgen05.ld (x32) and 16
//                 FADD2; after the tile's last partial sum it folds its cross chunk and runs the layer's tail on its
//                 32 values straight from registers.
//
// Why: what bounds the earlier arrangements is the time a warp needs to take one partial sum out of TMEM -- a
// tcgen05.ld round trip is ~160-250 clk and a warp that owns 128 columns needs four of them (two in flight) per
// partial sum, 800-1000 clk against ~600 clk of MMA work per partial sum (tools/ubench, profiles/r2_gemm2_*).  With 32
// columns per warp the round trip is paid once per partial sum and four warps per scheduler overlap theirs; no
// finished tile has to travel through TMEM to another warp, and 32 running sums leave the registers for the tail
// (640 threads: 96 registers at launch, 40 for the TMA / MMA warpgroup and 104 for the epilogue warpgroups after
// setmaxnreg).
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

constexpr int kT3Threads = 640;  // warpgroup 0: TMA warp, MMA warp, 2 idle; warpgroups 1-4: the epilogue warps
constexpr int kT3EpiWarps = 16;
constexpr int kT3StageA = kTcBM * 128;  // one plane of the activation tile: 128 rows x 128 B
constexpr int kT3StagingBytes = 2048;   // per epilogue warp: one fp16 plane of a 32 x 32 chunk

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
template <int PAT, bool FULL>
__global__ void __launch_bounds__(kT3Threads, 1) gemm_tc3_kernel(const __grid_constant__ TcParams p) {
  constexpr bool kStatic = PAT >= 0;
  constexpr int kTypes[4] = {kStatic ? ((PAT >> 0) & 15) - 1 : -1, kStatic ? ((PAT >> 4) & 15) - 1 : -1,
                             kStatic ? ((PAT >> 8) & 15) - 1 : -1, kStatic ? ((PAT >> 12) & 15) - 1 : -1};
  // index of the (first) BatchNorm scale / offset op of a static list
  constexpr int kSoIdx = kTypes[0] == EpiOp::kScaleOffset ? 0 : kTypes[1] == EpiOp::kScaleOffset ? 1 : kTypes[2] == EpiOp::kScaleOffset ? 2
                         : kTypes[3] == EpiOp::kScaleOffset ? 3 : -1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.bn * 128u;
  const uint32_t stage_bytes = 2u * kT3StageA + 2u * b_bytes;
  // [stages][16 x 2 KB epilogue staging tiles][barriers]
  const uint32_t epi0 = smem0 + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar0 = epi0 + (uint32_t)(kT3EpiWarps * kT3StagingBytes);
  // barriers: full[stages] | empty[stages] | setf[2] | sete[2] | xfree[2] | tmem slot
  auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(p.stages + s); };
  auto setf_bar = [&](uint32_t a) { return bar0 + 8u * (uint32_t)(2 * p.stages + a); };      // partial sum published
  auto sete_bar = [&](uint32_t a) { return bar0 + 8u * (uint32_t)(2 * p.stages + 2 + a); };  // main accumulator drained
  auto xfree_bar = [&](uint32_t a) { return bar0 + 8u * (uint32_t)(2 * p.stages + 4 + a); }; // cross accumulator taken
  const uint32_t tmem_slot = bar0 + 8u * (uint32_t)(2 * p.stages + 6);

  // thread coordinates read once (a volatile read is not re-materialised inside the loops)
  uint32_t tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = (int)(tid >> 5), lane = (int)(tid & 31);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; a++) {
      mbar_init(setf_bar(a), 1);
      mbar_init(sete_bar(a), kT3EpiWarps);
      mbar_init(xfree_bar(a), kT3EpiWarps);
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
  const int total_kb = p.total_kb;      // host-computed: summed here it was spilled to local memory and re-read in the fold loop
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
      const int step_m = (int)gridDim.x / p.tiles_n, step_n = (int)gridDim.x % p.tiles_n;
      int tm = (int)blockIdx.x / p.tiles_n, tn = (int)blockIdx.x % p.tiles_n;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = tm * kTcBM, n0 = tn * p.bn;
        tm += step_m;
        tn += step_n;
        if (tn >= p.tiles_n) {
          tn -= p.tiles_n;
          tm++;
        }
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
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes, fb = full_bar(stage);
          // (experiment, RS_B200_TC_PROFILE=8 / 16: leave out the weight / activation loads -- wrong results, shows what the L2 -> SM traffic costs)
          const bool skip_w = (p.profile & 8) && tile != (int)blockIdx.x, skip_a = (p.profile & 16) && tile != (int)blockIdx.x;
          mbar_expect_tx(fb, stage_bytes - (skip_w ? 2u * b_bytes : 0u) - (skip_a ? 2u * kT3StageA : 0u));
          if (!skip_a) {
            tma_load_2d(sa, &p.a_hi[s], kb * kTcBK, m0 + sl.yshift, fb);
            tma_load_2d(sa + kT3StageA, &p.a_lo[s], kb * kTcBK, m0 + sl.yshift, fb);
          }
          if (!skip_w) {
            tma_load_2d(sa + 2 * kT3StageA, &p.w_hi, sl.wk0 + kb * kTcBK, n0, fb);
            tma_load_2d(sa + 2 * kT3StageA + b_bytes, &p.w_lo, sl.wk0 + kb * kTcBK, n0, fb);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
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
      mbar_wait_lean(xfree_bar(xb), ((tcount >> 1) & 1u) ^ 1u);  // the epilogue warps have taken the cross sum of the tile before last
      tc_fence_after();
      for (int kb = 0; kb < total_kb; kb++) {
        mbar_wait_lean(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = ((smem0 + stage * stage_bytes) & 0x3ffffu) >> 4;  // descriptor address field, 16-byte units
        const uint32_t a_hi = sa, a_lo = sa + (kT3StageA >> 4), b_hi = sa + (2 * kT3StageA >> 4), b_lo = b_hi + (b_bytes >> 4);
#pragma unroll
        for (int half = 0; half < 2; half++) {  // one published partial sum = two K steps of 16 (32 bytes inside the swizzle atom)
          const uint32_t set = fcount & 1u;
          mbar_wait_lean(sete_bar(set), ((fcount >> 1) & 1u) ^ 1u);  // the epilogue warps have drained this accumulator
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
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ------------------------------------------------------------------ epilogue warps: (lane quadrant q, column chunk jc)
    const int q = warp & 3, jc = (warp - 4) >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(jc * 32);
    // this warp's 2 KB staging tile: one fp16 plane of the 32 x 32 chunk, 16-byte pieces XOR-swizzled by row pair
    uint4 *st16 = reinterpret_cast<uint4 *>(smem_raw + (epi0 - smem_u32(smem_raw)) + (uint32_t)(warp - 4) * kT3StagingBytes);
    const int first_op = bias_first ? 1 : 0;
    int ib = -1;  // the op whose split bypass input is prefetched (first kAddScaled with a split source)
    if constexpr (kStatic) {
      constexpr int kAs = kTypes[0] == EpiOp::kAddScaled ? 0 : kTypes[1] == EpiOp::kAddScaled ? 1 : kTypes[2] == EpiOp::kAddScaled ? 2
                          : kTypes[3] == EpiOp::kAddScaled ? 3 : -1;
      if constexpr (kAs >= 0) ib = p.ops[kAs].buf_lo ? kAs : -1;
    } else {
      for (int i = 0; i < p.n_ops && ib < 0; i++)
        if (p.ops[i].type == EpiOp::kAddScaled && p.ops[i].buf_lo) ib = i;
    }
    const bool chunk_in_tile = FULL || jc * 32 < p.bn;
    uint32_t fcount = 0, tcount = 0;
    // per-lane column of the per-column vectors of the coming tile: bias (start value), BatchNorm scale / offset
    auto load_cols = [&](int tile_n, bool in_range, float &b, float &s, float &o) {
      const int c = tile_n * p.bn + jc * 32 + lane;
      const bool ok = in_range && chunk_in_tile && (FULL || c < p.n);
      b = ok && bias_first ? __ldg(p.ops[0].v0 + c) : 0.f;
      s = o = 0.f;
      if constexpr (kStatic && kSoIdx >= 0) {
        if (ok) {
          s = __ldg(p.ops[kSoIdx].v0 + c);
          o = __ldg(p.ops[kSoIdx].v1 + c);
        }
      }
    };
    float col_b, col_s, col_o;
    // tile coordinates advance by a fixed (rows, columns) step: no division per tile
    const int step_m = (int)gridDim.x / p.tiles_n, step_n = (int)gridDim.x % p.tiles_n;
    int tm = (int)blockIdx.x / p.tiles_n, tn = (int)blockIdx.x % p.tiles_n;
    load_cols(tn, blockIdx.x < num_tiles, col_b, col_s, col_o);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
      const int m0 = tm * kTcBM, n0 = tn * p.bn;
      tm += step_m;
      tn += step_n;
      if (tn >= p.tiles_n) {
        tn -= p.tiles_n;
        tm++;
      }
      const int c0 = n0 + jc * 32;
      const bool valid = chunk_in_tile && (FULL || c0 < p.n);
      const int r = m0 + q * 32 + lane;
      const int rr = r < p.m ? r : p.m - 1;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = __shfl_sync(0xffffffffu, col_b, j);
      const float so_s = col_s, so_o = col_o;
      load_cols(tn, tile + (int)gridDim.x < num_tiles, col_b, col_s, col_o);  // next tile's columns: a whole tile of latency to hide in
      // bypass input of the chunk: 32 columns = 64 bytes per plane and row; 4 lanes x 16 B cover a row segment,
      // 8 rows per instruction
      uint4 pf_h[4], pf_l[4];
      auto prefetch_bypass = [&](int i) {
        const DevOp &op = p.ops[i];
        const int c8 = lane & 3;
        const bool ok = valid && (FULL || c0 + c8 * 8 < p.n) && !(p.profile & 64);  // (experiment 64: no bypass loads)
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
      // two partial sums before the end of the tile the bypass rows are pulled into L2 (no registers held during the
      // folds); the tail's loads then find them there
      auto prefetch_l2 = [&](int i) {
        const DevOp &op = p.ops[i];
        const int c8 = lane & 3;
        if (!(valid && (FULL || c0 + c8 * 8 < p.n))) return;
#pragma unroll
        for (int it = 0; it < 4; it++) {
          int ri = m0 + q * 32 + it * 8 + (lane >> 2);
          if (ri >= p.m) ri = p.m - 1;
          long long orow = op.den == op.num ? ri : ((long long)ri * op.num) / op.den;
          if (orow >= op.buf_rows) orow = op.buf_rows - 1;
          const size_t off = (size_t)orow * op.buf_ld + c0 + c8 * 8;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __half *>(op.buf) + off));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __half *>(op.buf_lo) + off));
        }
      };
      const int pf_at = total_sums >= 3 ? total_sums - 3 : 0;
#pragma unroll 1
      for (int f = 0; f < total_sums; f++, fcount++) {
        if (f == pf_at && ib >= 0) prefetch_l2(ib);
        const uint32_t set = fcount & 1u;
        mbar_wait(setf_bar(set), (fcount >> 1) & 1u);
        tc_fence_after();
        if (valid) {
          uint32_t raw[32];
          tmem_ld32_nowait(lane_base + set * (uint32_t)p.bn, raw);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 2) add2(v[j], v[j + 1], __uint_as_float(raw[j]), __uint_as_float(raw[j + 1]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sete_bar(set));
      }
      // the tile's cross sum (the last partial sum's commit covers it): v += cross * 2^-11
      const uint32_t xb = tcount & 1u;
      if (valid) {
        uint32_t raw[32];
        tmem_ld32_nowait(lane_base + (2u + xb) * (uint32_t)p.bn, raw);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmaf(__uint_as_float(raw[j]), 1.f / kSplitScale, v[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(xfree_bar(xb));
      if (!valid) continue;
      if (p.profile & 128) continue;  // (experiment 128: no tail at all)

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
                float *vs = reinterpret_cast<float *>(st16);
                __syncwarp();
                vs[lane] = so_s;
                vs[32 + lane] = so_o;
                __syncwarp();
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
            if (op.buf_lo) {
              // split source, read with full-row coalescing (prefetch_bypass); each plane goes through the staging
              // tile so that a thread gets the 32 halves of its own row, then value = hi + lo / 2048 (exact)
              const int c8 = lane & 3;
              prefetch_bypass(i);  // (requesting the rows before the cross fold / ReLU / BatchNorm was measured: slower, spills)
              uint4 rh[4], rl[4];
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 4; it++) {
                const int ii = it * 8 + (lane >> 2);
                st16[ii * 4 + (c8 ^ ((ii >> 1) & 3))] = pf_h[it];
              }
              __syncwarp();
#pragma unroll
              for (int c = 0; c < 4; c++) rh[c] = st16[lane * 4 + (c ^ ((lane >> 1) & 3))];
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 4; it++) {
                const int ii = it * 8 + (lane >> 2);
                st16[ii * 4 + (c8 ^ ((ii >> 1) & 3))] = pf_l[it];
              }
              __syncwarp();
#pragma unroll
              for (int c = 0; c < 4; c++) rl[c] = st16[lane * 4 + (c ^ ((lane >> 1) & 3))];
#pragma unroll
              for (int c = 0; c < 4; c++) {
                const __half2 *hh = reinterpret_cast<const __half2 *>(&rh[c]), *ll = reinterpret_cast<const __half2 *>(&rl[c]);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  const float2 fh = __half22float2(hh[e]), fl = __half22float2(ll[e]);
                  float x0 = fmaf(fl.x, 1.f / kSplitScale, fh.x), x1 = fmaf(fl.y, 1.f / kSplitScale, fh.y);
                  if (op.alpha != 1.f) mul2(x0, x1, op.alpha, op.alpha);
                  add2(v[c * 8 + 2 * e], v[c * 8 + 2 * e + 1], x0, x1);
                }
              }
            } else {
              // plain fp32 source: every thread reads its own row (rare: only the generic op list gets here)
              int ri = rr;
              long long orow = op.den == op.num ? ri : ((long long)ri * op.num) / op.den;
              if (orow >= op.buf_rows) orow = op.buf_rows - 1;
              const float *src = reinterpret_cast<const float *>(op.buf) + (size_t)orow * op.buf_ld + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (FULL || c0 + j < p.n) {
                  float4 o = __ldcs(reinterpret_cast<const float4 *>(src + j));
                  if (op.alpha != 1.f) {
                    mul2(o.x, o.y, op.alpha, op.alpha);
                    mul2(o.z, o.w, op.alpha, op.alpha);
                  }
                  add2(v[j], v[j + 1], o.x, o.y);
                  add2(v[j + 2], v[j + 3], o.z, o.w);
                }
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
      // store through the staging tile so that every instruction writes whole row segments
      if (p.out_lo) {
        // two fp16 planes, one after the other through the 2 KB tile
        uint4 hh[4], ll[4];
        float amax = 0.f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          split2x(v[8 * c + 0], v[8 * c + 1], hh[c].x, ll[c].x, amax);
          split2x(v[8 * c + 2], v[8 * c + 3], hh[c].y, ll[c].y, amax);
          split2x(v[8 * c + 4], v[8 * c + 5], hh[c].z, ll[c].z, amax);
          split2x(v[8 * c + 6], v[8 * c + 7], hh[c].w, ll[c].w, amax);
        }
        if (amax > 65504.f && r < p.m) *p.range_flag = 1;
        const int c8 = lane & 3;
#pragma unroll
        for (int plane = 0; plane < 2; plane++) {
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; c++) st16[lane * 4 + (c ^ ((lane >> 1) & 3))] = plane == 0 ? hh[c] : ll[c];
          __syncwarp();
          __half *out = reinterpret_cast<__half *>(plane == 0 ? p.out_hi : p.out_lo);
#pragma unroll
          for (int it = 0; it < 4; it++) {
            const int i = it * 8 + (lane >> 2);
            const int ri = m0 + q * 32 + i;
            if (ri < p.m && (FULL || c0 + c8 * 8 < p.n) && !(p.profile & 32))  // (experiment 32: no stores)
              *reinterpret_cast<uint4 *>(out + (size_t)ri * p.out_ld + c0 + c8 * 8) = st16[i * 4 + (c8 ^ ((i >> 1) & 3))];
          }
        }
      } else {
        // plain fp32 output (the network's last layer): two halves of 16 columns through the 2 KB tile,
        // 4 lanes x 16 B cover a 64-byte row segment, 8 rows per instruction
        const int c4 = lane & 3;
#pragma unroll
        for (int hcol = 0; hcol < 2; hcol++) {
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; c++)
            st16[lane * 4 + (c ^ ((lane >> 1) & 3))] =
                make_uint4(__float_as_uint(v[hcol * 16 + 4 * c]), __float_as_uint(v[hcol * 16 + 4 * c + 1]),
                           __float_as_uint(v[hcol * 16 + 4 * c + 2]), __float_as_uint(v[hcol * 16 + 4 * c + 3]));
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; it++) {
            const int i = it * 8 + (lane >> 2);
            const int ri = m0 + q * 32 + i;
            const int col = c0 + hcol * 16 + c4 * 4;
            if (ri < p.m && (FULL || col < p.n))
              *reinterpret_cast<uint4 *>(reinterpret_cast<float *>(p.out_hi) + (size_t)ri * p.out_ld + col) = st16[i * 4 + (c4 ^ ((i >> 1) & 3))];
          }
        }
      }
      __syncwarp();
    }
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

template <int PAT, bool FULL>
void Launch(const TcParams &p, int grid, int smem, int smem_limit, cudaStream_t stream) {
  static int configured_dev = -1;  // opt-in shared memory size is a per-device function attribute
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc3_kernel<PAT, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit);
    if (e != cudaSuccess) RS_FAIL("cudaFuncSetAttribute(gemm_tc3_kernel): " << cudaGetErrorString(e));
    configured_dev = dev;
  }
  gemm_tc3_kernel<PAT, FULL><<<grid, kT3Threads, smem, stream>>>(p);
}

}  // namespace

void LaunchGemmTc3(const TcParams &p, int num_sms, int smem_limit, cudaStream_t stream) {
  const int stage_bytes = 2 * kT3StageA + 2 * p.bn * 128;
  const int smem = 1024 + p.stages * stage_bytes + kT3EpiWarps * kT3StagingBytes + 8 * (2 * p.stages + 6) + 16;
  int grid = p.tiles_m * p.tiles_n;
  if (grid > num_sms) grid = num_sms;
  const bool full = p.bn == 128 && p.n % 128 == 0;
  const int pat = Pattern(p);
  if (full && pat == kPatNone) Launch<kPatNone, true>(p, grid, smem, smem_limit, stream);
  else if (full && pat == kPatBRS) Launch<kPatBRS, true>(p, grid, smem, smem_limit, stream);
  else if (full && pat == kPatBRSA) Launch<kPatBRSA, true>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatNone) Launch<kPatNone, false>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatBias) Launch<kPatBias, false>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatBRS) Launch<kPatBRS, false>(p, grid, smem, smem_limit, stream);
  else if (pat == kPatBRSA) Launch<kPatBRSA, false>(p, grid, smem, smem_limit, stream);
  else Launch<-1, false>(p, grid, smem, smem_limit, stream);
}

}  // namespace rs
