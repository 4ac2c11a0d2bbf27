// Stage (ii): nnet3 TDNN(-F) forward on the global time axis.
//
// Replaces NnetComputer::Run over the compiled looped computation
// (kaldi/src/nnet3/decodable-online-looped.cc:118-236, nnet-compute.cc) for the component types a
// TDNN-F chain model is made of:
//   TdnnComponent::Propagate            nnet3/nnet-tdnn-component.cc:181-211  (sum over time offsets)
//   Affine/NaturalGradientAffine/Linear/FixedAffine::Propagate  nnet3/nnet-simple-component.cc
//   RectifiedLinearComponent, BatchNormComponent (test mode, nnet-normalize-component.cc:453-463),
//   dropout/no-op (identity in test mode), Sum(Scale(a, x), y) bypass descriptors.
// A layer is ONE launch: out[r,:] = epilogue( sum_slabs  A_slab[row(r),:] * W[:, slab cols]^T ) where
// the slabs are the time-offset / Append blocks of the component input (no im2col copy) and the
// epilogue applies bias -> ReLU -> BatchNorm scale/offset -> bypass add in the reference's order,
// each with its own rounding (no FMA contraction) so the only numerical difference to the CPU
// path is the summation order inside the dot products.
//
// This file holds the fp32 CUDA-core path (exact fp32 products, fp32 FMA accumulation).
#include "engine.h"
#include "model.h"
#include "split.cuh"

namespace rs {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
constexpr int kGemmThreads = 256;
constexpr int kChunkTiles = 4;  // 4 x BK = 64 products per first-level sum

__device__ __forceinline__ float apply_ops(float v, int r, int c, const GemmParams &p) {
#pragma unroll 1
  for (int i = 0; i < p.n_ops; i++) {
    const DevOp &op = p.ops[i];
    switch (op.type) {
      case EpiOp::kBias:
        v = __fadd_rn(v, op.v0[c]);
        break;
      case EpiOp::kRelu:
        v = v > 0.f ? v : 0.f;
        break;
      case EpiOp::kScaleOffset:
        v = __fadd_rn(__fmul_rn(v, op.v0[c]), op.v1[c]);
        break;
      case EpiOp::kScale:
        v = __fmul_rn(v, op.alpha);
        break;
      case EpiOp::kAddScaled: {
        long long orow = ((long long)r * op.num) / op.den;
        if (orow >= op.buf_rows) orow = op.buf_rows - 1;
        const float o = ld_act(op.buf, op.buf_lo, (size_t)orow * op.buf_ld + c);
        v = __fadd_rn(op.alpha == 1.f ? o : __fmul_rn(op.alpha, o), v);
        break;
      }
      case EpiOp::kUttBias: {
        int u = p.row_utt[(size_t)r * op.num];
        v = __fadd_rn(v, reinterpret_cast<const float *>(op.buf)[(size_t)u * op.buf_ld + c]);
        break;
      }
    }
  }
  return v;
}

__device__ __forceinline__ long long slab_row(const GemmSlab &s, int r) {
  long long t = (long long)r * s.num + s.shift;
  long long q = t >= 0 ? t / s.den : -((-t + s.den - 1) / s.den);
  if (q < 0) q = 0;
  if (q >= s.rows) q = s.rows - 1;
  return q;
}

template <bool kVec>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const __grid_constant__ GemmParams p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  // loader mapping: each thread moves two float4 of A and two of B per k-tile
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  // k-tiles over all slabs
  int n_tiles = 0;
  for (int s = 0; s < p.n_slabs; s++) n_tiles += (p.slabs[s].k + BK - 1) / BK;

  float4 ra[2], rb[2];
  int cur_slab = 0, cur_k0 = 0;
  const float *arow[2] = {nullptr, nullptr};  // plain fp32 source rows
  size_t aoff[2] = {0, 0};                    // element offset of the row (split sources)
  auto set_slab = [&](int s) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int r = row0 + lrow + h * 64;
      if (r >= p.m) r = p.m - 1;
      aoff[h] = (size_t)slab_row(p.slabs[s], r) * p.slabs[s].ld;
      arow[h] = reinterpret_cast<const float *>(p.slabs[s].src) + aoff[h];
    }
  };
  auto load_tile = [&]() {
    const GemmSlab &sl = p.slabs[cur_slab];
    const int k = cur_k0 + lk;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sl.src_lo) {  // split source (two fp16 planes)
        if (k + 0 < sl.k) v.x = ld_act(sl.src, sl.src_lo, aoff[h] + k + 0);
        if (k + 1 < sl.k) v.y = ld_act(sl.src, sl.src_lo, aoff[h] + k + 1);
        if (k + 2 < sl.k) v.z = ld_act(sl.src, sl.src_lo, aoff[h] + k + 2);
        if (k + 3 < sl.k) v.w = ld_act(sl.src, sl.src_lo, aoff[h] + k + 3);
      } else if (kVec) {
        if (k < sl.k) v = *reinterpret_cast<const float4 *>(arow[h] + k);
      } else {
        if (k + 0 < sl.k) v.x = arow[h][k + 0];
        if (k + 1 < sl.k) v.y = arow[h][k + 1];
        if (k + 2 < sl.k) v.z = arow[h][k + 2];
        if (k + 3 < sl.k) v.w = arow[h][k + 3];
      }
      ra[h] = v;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      int c = col0 + lrow + h * 64;
      if (c < p.n) {
        const float *wp = p.w + (size_t)c * p.ktot + sl.wcol + k;
        if (kVec) {
          if (k < sl.k) w = *reinterpret_cast<const float4 *>(wp);
        } else {
          if (k + 0 < sl.k) w.x = wp[0];
          if (k + 1 < sl.k) w.y = wp[1];
          if (k + 2 < sl.k) w.z = wp[2];
          if (k + 3 < sl.k) w.w = wp[3];
        }
      }
      rb[h] = w;
    }
    // advance to the next k-tile
    cur_k0 += BK;
    if (cur_k0 >= sl.k) {
      cur_k0 = 0;
      cur_slab++;
      if (cur_slab < p.n_slabs) set_slab(cur_slab);
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x;
      As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z;
      As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x;
      Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z;
      Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  // Two-level summation: products are accumulated in `part` over kChunkTiles k-tiles (64 terms) and
  // then folded into `acc`, the way a K-blocked BLAS kernel folds register sums into C.  This keeps
  // the rounding error of long dot products (K up to 2048) at the level of the reference's sgemm.
  float acc[8][8], part[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = part[i][j] = 0.f;
  const int tx = tid & 15, ty = tid >> 4;

  set_slab(0);
  load_tile();
  store_tile(0);
  __syncthreads();
  for (int it = 0; it < n_tiles; it++) {
    const int buf = it & 1;
    if (it + 1 < n_tiles) load_tile();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
    if ((it % kChunkTiles) == kChunkTiles - 1 || it + 1 == n_tiles) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          acc[i][j] += part[i][j];
          part[i][j] = 0.f;
        }
    }
    if (it + 1 < n_tiles) store_tile(buf ^ 1);
    __syncthreads();
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= p.m) continue;
#pragma unroll
    for (int jh = 0; jh < 2; jh++) {
      int c = col0 + jh * 64 + tx * 4;
      if (c >= p.n) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = (c + j < p.n) ? apply_ops(acc[i][jh * 4 + j], r, c + j, p) : 0.f;
      float *o = reinterpret_cast<float *>(p.out) + (size_t)r * p.out_ld + c;
      if (p.out_lo) {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (c + j < p.n) st_act(p.out, p.out_lo, (size_t)r * p.out_ld + c + j, v[j], p.range_flag);
      } else if (c + 3 < p.n && (p.out_ld & 3) == 0) {
        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (c + j < p.n) o[j] = v[j];
      }
    }
  }
}

void LaunchGemm(const GemmParams &p, cudaStream_t stream) {
  if (p.m <= 0 || p.n <= 0) return;
  bool vec = (p.ktot % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.w) & 15) == 0);
  for (int s = 0; s < p.n_slabs; s++) {
    const GemmSlab &sl = p.slabs[s];
    if (!sl.src_lo && (sl.k % 4 || sl.ld % 4 || sl.wcol % 4 || (reinterpret_cast<uintptr_t>(sl.src) & 15))) vec = false;
    if (sl.src_lo && (sl.wcol % 4)) vec = false;
  }
  dim3 grid((p.n + BN - 1) / BN, (p.m + BM - 1) / BM);
  if (vec)
    gemm_kernel<true><<<grid, kGemmThreads, 0, stream>>>(p);
  else
    gemm_kernel<false><<<grid, kGemmThreads, 0, stream>>>(p);
}

// --------------------------------------------------------------------------- elementwise
// out[r, col_offset + c] = ops( s0*x0[row0(r), c] (+ s1*x1[row1(r), c] ...) ): nodes that could not be
// fused into a producing GEMM.  Descriptor sums follow nnet3's copy-then-add order.
struct ElemScales {
  float s[kMaxSlabs];
};
__global__ void __launch_bounds__(256) elementwise_kernel(const __grid_constant__ GemmParams p, ElemScales sc, int col_offset) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.m * p.n) return;
  int r = (int)(i / p.n), c = (int)(i - (long long)r * p.n);
  float v = 0.f;
  for (int s = 0; s < p.n_slabs; s++) {
    const GemmSlab &sl = p.slabs[s];
    const size_t idx = (size_t)slab_row(sl, r) * sl.ld + sl.wcol + c;
    const float x = ld_act(sl.src, sl.src_lo, idx);
    float t = sc.s[s] == 1.f ? x : __fmul_rn(sc.s[s], x);
    v = s == 0 ? t : __fadd_rn(v, t);
  }
  v = apply_ops(v, r, c, p);
  const size_t oidx = (size_t)r * p.out_ld + col_offset + c;
  st_act(p.out, p.out_lo, oidx, v, p.range_flag);
}

void LaunchElementwise(const GemmParams &p, const float *term_scale_host, int col_offset, cudaStream_t stream) {
  if (p.m <= 0 || p.n <= 0) return;
  ElemScales sc;
  for (int s = 0; s < kMaxSlabs; s++) sc.s[s] = s < p.n_slabs ? term_scale_host[s] : 1.f;
  long long total = (long long)p.m * p.n;
  elementwise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p, sc, col_offset);
}

// --------------------------------------------------------------------------- log-softmax (warp/row)
// `p` carries the output matrix and the ops fused behind the log-softmax (log-prior subtraction and
// acoustic scale when the log-softmax is the network output, decodable-online-looped.cc:218-223)
__global__ void __launch_bounds__(256) logsoftmax_kernel(const void *in, const void *in_lo, int in_ld, const __grid_constant__ GemmParams p) {
  float *out = reinterpret_cast<float *>(p.out);
  const int out_ld = p.out_ld, rows = p.m, n = p.n;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  auto at = [&](int c) { return ld_act(in, in_lo, (size_t)row * in_ld + c); };
  float mx = -3.4e38f;
  for (int c = lane; c < n; c += 32) mx = fmaxf(mx, at(c));
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s += expf(at(c) - mx);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float lse = mx + logf(s);
  for (int c = lane; c < n; c += 32) out[(size_t)row * out_ld + c] = apply_ops(at(c) - lse, row, c, p);
}

void LaunchLogSoftmax(const void *in, const void *in_lo, int in_ld, const GemmParams &p, cudaStream_t stream) {
  if (p.m <= 0) return;
  logsoftmax_kernel<<<(p.m + 7) / 8, 256, 0, stream>>>(in, in_lo, in_ld, p);
}

// --------------------------------------------------------------------------- input assembly
// Copies each utterance's features onto the global time axis with the reference's edge handling:
// frames before 0 / after T-1 repeat the first / last frame (decodable-online-looped.cc:150-161).
__global__ void __launch_bounds__(256) assemble_kernel(AssembleParams p) {
  const int u = blockIdx.y;
  const int T = p.num_frames[u];
  if (T <= 0) return;
  const int span = T + p.left + p.right;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)span * p.dim;
       i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i / p.dim), d = (int)(i - (long long)w * p.dim);
    int t = w - p.left;
    int row = p.origin[u] + t;
    if (row < 0 || row >= p.axis_len) continue;
    int tc = t < 0 ? 0 : (t >= T ? T - 1 : t);
    const float x = p.feats[((size_t)p.frame_offset[u] + tc) * p.dim + d];
    st_act(p.dst, p.dst_lo, (size_t)row * p.ld + d, x, p.range_flag);
  }
}

void LaunchAssembleInput(const AssembleParams &p, int n_utts, int max_rows, cudaStream_t stream) {
  if (n_utts == 0) return;
  int blocks = (int)(((long long)max_rows * p.dim + 255) / 256);
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  assemble_kernel<<<dim3(blocks, n_utts), 256, 0, stream>>>(p);
}

}  // namespace rs
