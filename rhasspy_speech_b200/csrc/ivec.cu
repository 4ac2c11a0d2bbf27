// Stage (i), second half: online CMVN, splice + LDA, UBM posteriors, iVector statistics and the
// conjugate-gradient solve -- the work of OnlineIvectorFeature in the reference:
//   kaldi/src/feat/online-feature.cc:337-452      OnlineCmvn (+ transform/cmvn.cc:64-115 ApplyCmvn)
//   kaldi/src/feat/online-feature.cc:504-554      OnlineSpliceFrames, OnlineTransform (LDA)
//   kaldi/src/gmm/diag-gmm.cc:546-562             DiagGmm::LogLikelihoods
//   kaldi/src/hmm/posterior.cc:440-508            VectorToPosteriorEntry
//   kaldi/src/online2/online-ivector-feature.cc:188-279, 327-355   scheduling (offline: one update
//                                                 over all frames, then one CG from the prior)
//   kaldi/src/ivector/ivector-extractor.cc:611-668, 732-756        AccStats, GetIvector
//   kaldi/src/matrix/optimization.cc:453-560      LinearCgd
#include <cfloat>
#include <cstdio>

#include "engine.h"
#include "smem_attr.h"

namespace rs {

// names the kernel whose launch configuration the runtime rejected (the caller only sees cudaGetLastError later)
#define RS_CHECK_LAUNCH(name)                                                                          \
  do {                                                                                                \
    cudaError_t e_ = cudaPeekAtLastError();                                                           \
    if (e_ != cudaSuccess) fprintf(stderr, "rs_b200: launch of %s failed: %s\n", name, cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------ CMVN
// SmoothOnlineCmvnStats with no speaker stats (spk == utt in the reference's invocation) + ApplyCmvn for one value
__device__ __forceinline__ float CmvnApply(const CmvnParams &p, float xin, double s0, double s1, double cnt, double g0, double g1, double gcount) {
  double a0 = s0, a1 = s1, c = cnt;
  if (c < (double)p.cmn_window) {
    double cg = (double)p.cmn_window - c;
    if (cg > (double)p.global_frames) cg = (double)p.global_frames;
    if (cg > 0.0) {
      double f = cg / gcount;
      a0 += f * g0;
      a1 += f * g1;
      c += f * gcount;
    }
  }
  if (!p.normalize_mean) return xin;
  if (!p.normalize_variance) {
    float alpha = (float)(-1.0 / c);          // VectorBase<float>::AddVec(float alpha, Vector<double>)
    float off = (float)((double)alpha * a0);
    return __fadd_rn(xin, off);
  }
  double mean = a0 / c;
  double var = a1 / c - mean * mean;
  if (var < 1.0e-20) var = 1.0e-20;
  double scale = 1.0 / sqrt(var);
  float sc = (float)scale, off = (float)(-(mean * scale));
  return __fadd_rn(__fmul_rn(xin, sc), off);
}

// One thread per (utterance, dim): the sliding-window sums are sequential in double exactly as
// ComputeStatsForFrame accumulates them (add frame t, then subtract frame t - cmn_window).
__global__ void cmvn_kernel(CmvnParams p) {
  const int u = blockIdx.x, d = threadIdx.x;
  if (d >= p.dim) return;
  const int T = p.num_frames[u];
  const float *in = p.in + (size_t)p.frame_offset[u] * p.dim;
  float *out = p.out + (size_t)p.frame_offset[u] * p.dim;
  const double g0 = p.global_stats[d], g1 = p.global_stats[(p.dim + 1) + d], gcount = p.global_stats[p.dim];
  double s0 = 0.0, s1 = 0.0, cnt = 0.0;
  constexpr int kPf = 8;  // frames fetched per batch: one exposed memory latency per 8 sequential updates
  float xbuf[kPf];
  for (int t = 0; t < T; t++) {
    if ((t & (kPf - 1)) == 0) {
#pragma unroll
      for (int i = 0; i < kPf; i++) xbuf[i] = t + i < T ? in[(size_t)(t + i) * p.dim + d] : 0.f;
    }
    float xin_t = 0.f;
#pragma unroll
    for (int i = 0; i < kPf; i++)
      if ((t & (kPf - 1)) == i) xin_t = xbuf[i];
    double x = (double)xin_t;
    s0 += x;
    if (p.normalize_variance) s1 += x * x;
    cnt += 1.0;
    int prev = t - p.cmn_window;
    if (prev >= 0) {
      double y = (double)in[(size_t)prev * p.dim + d];
      s0 -= y;
      if (p.normalize_variance) s1 -= y * y;
      cnt -= 1.0;
    }
    // SmoothOnlineCmvnStats with no speaker stats (spk == utt in the reference's invocation)
    double a0 = s0, a1 = s1, c = cnt;
    if (c < (double)p.cmn_window) {
      double cg = (double)p.cmn_window - c;
      if (cg > (double)p.global_frames) cg = (double)p.global_frames;
      if (cg > 0.0) {
        double f = cg / gcount;
        a0 += f * g0;
        a1 += f * g1;
        c += f * gcount;
      }
    }
    float xin = xin_t, y;
    if (!p.normalize_mean) {
      y = xin;
    } else if (!p.normalize_variance) {
      float alpha = (float)(-1.0 / c);          // VectorBase<float>::AddVec(float alpha, Vector<double>)
      float off = (float)((double)alpha * a0);
      y = __fadd_rn(xin, off);
    } else {
      double mean = a0 / c;
      double var = a1 / c - mean * mean;
      if (var < 1.0e-20) var = 1.0e-20;
      double scale = 1.0 / sqrt(var);
      float sc = (float)scale, off = (float)(-(mean * scale));
      y = __fadd_rn(__fmul_rn(xin, sc), off);
    }
    out[(size_t)t * p.dim + d] = y;
  }
}

// The same statistics for utterances no longer than the window (nothing is ever subtracted): thread (d, segment)
// sums its segment of frames in double, the segment sums are prefix-summed, and every thread then walks its own segment
// from that start value -- the sequential depth is T / segments instead of T.  The running sums are sums of floats in
// double: no rounding happens for features of ordinary dynamic range, so they equal the reference's (and are within
// 1e-16 relative otherwise); everything after the sums is the code above.
constexpr int kCmvnMaxSeg = 32;
__global__ void __launch_bounds__(1024) cmvn_seg_kernel(CmvnParams p, int nseg) {
  extern __shared__ double segsum[];  // [2][nseg][dim]
  const int u = blockIdx.x, d = threadIdx.x % p.dim, seg = threadIdx.x / p.dim;
  const int T = p.num_frames[u];
  const float *in = p.in + (size_t)p.frame_offset[u] * p.dim;
  float *out = p.out + (size_t)p.frame_offset[u] * p.dim;
  const double g0 = p.global_stats[d], g1 = p.global_stats[(p.dim + 1) + d], gcount = p.global_stats[p.dim];
  if (T > p.cmn_window) {
    // longer than the window: the sequential form (add frame t, subtract frame t - window)
    if (seg != 0) return;
    double s0 = 0.0, s1 = 0.0, cnt = 0.0;
    for (int t = 0; t < T; t++) {
      const float xin = in[(size_t)t * p.dim + d];
      double x = (double)xin;
      s0 += x;
      if (p.normalize_variance) s1 += x * x;
      cnt += 1.0;
      const int prev = t - p.cmn_window;
      if (prev >= 0) {
        double y = (double)in[(size_t)prev * p.dim + d];
        s0 -= y;
        if (p.normalize_variance) s1 -= y * y;
        cnt -= 1.0;
      }
      out[(size_t)t * p.dim + d] = CmvnApply(p, xin, s0, s1, cnt, g0, g1, gcount);
    }
    return;
  }
  const int per = (T + nseg - 1) / nseg, t0 = seg * per, t1 = min(T, t0 + per);
  double a0 = 0.0, a1 = 0.0;
  for (int t = t0; t < t1; t++) {
    const double x = (double)in[(size_t)t * p.dim + d];
    a0 += x;
    if (p.normalize_variance) a1 += x * x;
  }
  segsum[seg * p.dim + d] = a0;
  segsum[(nseg + seg) * p.dim + d] = a1;
  __syncthreads();
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < seg; k++) {
    s0 += segsum[k * p.dim + d];
    s1 += segsum[(nseg + k) * p.dim + d];
  }
  for (int t = t0; t < t1; t++) {
    const float xin = in[(size_t)t * p.dim + d];
    const double x = (double)xin;
    s0 += x;
    if (p.normalize_variance) s1 += x * x;
    out[(size_t)t * p.dim + d] = CmvnApply(p, xin, s0, s1, (double)(t + 1), g0, g1, gcount);
  }
}

void LaunchCmvn(const CmvnParams &p, int n_utts, cudaStream_t stream) {
  if (n_utts == 0) return;
  int nseg = 1024 / p.dim;
  if (nseg > kCmvnMaxSeg) nseg = kCmvnMaxSeg;
  if (nseg >= 2) {
    cmvn_seg_kernel<<<n_utts, nseg * p.dim, (size_t)2 * nseg * p.dim * sizeof(double), stream>>>(p, nseg);
    RS_CHECK_LAUNCH("cmvn_seg_kernel");
    return;
  }
  int threads = ((p.dim + 31) / 32) * 32;
  cmvn_kernel<<<n_utts, threads, 0, stream>>>(p);
  RS_CHECK_LAUNCH("cmvn_kernel");
}

// ---------------------------------------------------------------------------- splice + LDA
// CTA = 16 frames of one utterance; the spliced window (16 + left + right frames) is staged in
// shared memory for both the raw and the normalised stream; thread (f, j) -> one output each.
constexpr int kLdaFrames = 16;
__global__ void __launch_bounds__(256) splice_lda_kernel(IvecParams p) {
  extern __shared__ float sm[];
  const int u = blockIdx.y;
  const int T = p.num_frames[u];
  const int t0 = blockIdx.x * kLdaFrames;
  if (t0 >= T) return;
  const int W = kLdaFrames + p.left + p.right;
  float *sraw = sm, *snorm = sm + (size_t)W * p.dim;
  const size_t base = (size_t)p.frame_offset[u];
  for (int i = threadIdx.x; i < W * p.dim; i += blockDim.x) {
    int w = i / p.dim, d = i - w * p.dim;
    int t = t0 - p.left + w;
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);  // OnlineSpliceFrames clamps at both ends
    sraw[i] = p.mfcc[(base + t) * p.dim + d];
    snorm[i] = p.mfcc_norm[(base + t) * p.dim + d];
  }
  __syncthreads();
  const int K = p.dim * (p.left + 1 + p.right);
  // thread = (output column j, group of 4 frames): each LDA weight is loaded once and used for 4 frames x 2
  // streams; the sum over k runs in ascending order exactly as before
  for (int o = threadIdx.x; o < (kLdaFrames / 4) * p.ldim; o += blockDim.x) {
    const int fg = o / p.ldim, j = o - fg * p.ldim;
    const float *xr = sraw + (size_t)fg * 4 * p.dim, *xn = snorm + (size_t)fg * 4 * p.dim;
    float ar[4] = {0.f, 0.f, 0.f, 0.f}, an[4] = {0.f, 0.f, 0.f, 0.f};
    int k = 0;
    if ((p.dim & 3) == 0) {  // 16-byte shared-memory reads along k (rows of the staged window stay 16-byte aligned)
      for (; k + 4 <= K; k += 4) {
        float w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = __ldg(p.lda_t + (size_t)(k + i) * p.ldim + j);
#pragma unroll
        for (int f = 0; f < 4; f++) {
          const float4 a = *reinterpret_cast<const float4 *>(xr + f * p.dim + k);
          const float4 b = *reinterpret_cast<const float4 *>(xn + f * p.dim + k);
          ar[f] = fmaf(w[3], a.w, fmaf(w[2], a.z, fmaf(w[1], a.y, fmaf(w[0], a.x, ar[f]))));
          an[f] = fmaf(w[3], b.w, fmaf(w[2], b.z, fmaf(w[1], b.y, fmaf(w[0], b.x, an[f]))));
        }
      }
    }
    for (; k < K; k++) {
      const float w = __ldg(p.lda_t + (size_t)k * p.ldim + j);
#pragma unroll
      for (int f = 0; f < 4; f++) {
        ar[f] = fmaf(w, xr[f * p.dim + k], ar[f]);
        an[f] = fmaf(w, xn[f * p.dim + k], an[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int t = t0 + fg * 4 + f;
      if (t >= T) continue;
      float r = ar[f], n = an[f];
      if (p.lda_bias) {
        r += p.lda_bias[j];
        n += p.lda_bias[j];
      }
      p.x_raw[(base + t) * p.ldim + j] = r;
      p.x_norm[(base + t) * p.ldim + j] = n;
    }
  }
}

// The same transform for the common layout (dim and ldim multiples of 4): CTA = 64 frames of one utterance, a thread
// owns 4 frames x 4 output columns of both streams, so four weights (one 16-byte load) and eight 16-byte shared-memory
// reads feed 128 FMAs.  Every sum runs over k in ascending order with one FMA per term, as in splice_lda_kernel.
constexpr int kLdaFrames4 = 64;
__global__ void __launch_bounds__(256) splice_lda4_kernel(IvecParams p) {
  extern __shared__ float sm[];
  const int u = blockIdx.y;
  const int T = p.num_frames[u];
  const int t0 = blockIdx.x * kLdaFrames4;
  if (t0 >= T) return;
  const int W = kLdaFrames4 + p.left + p.right;
  float *sraw = sm, *snorm = sm + (size_t)W * p.dim;
  const size_t base = (size_t)p.frame_offset[u];
  const int dim4 = p.dim >> 2;
  for (int i = threadIdx.x; i < W * dim4; i += blockDim.x) {
    const int w = i / dim4, d4 = i - w * dim4;
    int t = t0 - p.left + w;
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);  // OnlineSpliceFrames clamps at both ends
    reinterpret_cast<float4 *>(sraw)[i] = __ldg(reinterpret_cast<const float4 *>(p.mfcc + (base + t) * p.dim) + d4);
    reinterpret_cast<float4 *>(snorm)[i] = __ldg(reinterpret_cast<const float4 *>(p.mfcc_norm + (base + t) * p.dim) + d4);
  }
  __syncthreads();
  const int K = p.dim * (p.left + 1 + p.right), jgroups = p.ldim >> 2;
  for (int o = threadIdx.x; o < (kLdaFrames4 / 4) * jgroups; o += blockDim.x) {
    const int fg = o / jgroups, j0 = (o - fg * jgroups) * 4;
    if (t0 + fg * 4 >= T) continue;
    const float *xr = sraw + (size_t)fg * 4 * p.dim, *xn = snorm + (size_t)fg * 4 * p.dim;
    float ar[4][4], an[4][4];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int j = 0; j < 4; j++) ar[f][j] = an[f][j] = 0.f;
    for (int k = 0; k < K; k += 4) {
      float4 w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) w[i] = __ldg(reinterpret_cast<const float4 *>(p.lda_t + (size_t)(k + i) * p.ldim + j0));
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const float4 a = *reinterpret_cast<const float4 *>(xr + f * p.dim + k);
        const float4 b = *reinterpret_cast<const float4 *>(xn + f * p.dim + k);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          ar[f][0] = fmaf(w[i].x, av[i], ar[f][0]);
          ar[f][1] = fmaf(w[i].y, av[i], ar[f][1]);
          ar[f][2] = fmaf(w[i].z, av[i], ar[f][2]);
          ar[f][3] = fmaf(w[i].w, av[i], ar[f][3]);
          an[f][0] = fmaf(w[i].x, bv[i], an[f][0]);
          an[f][1] = fmaf(w[i].y, bv[i], an[f][1]);
          an[f][2] = fmaf(w[i].z, bv[i], an[f][2]);
          an[f][3] = fmaf(w[i].w, bv[i], an[f][3]);
        }
      }
    }
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.lda_bias) bias = __ldg(reinterpret_cast<const float4 *>(p.lda_bias + j0));
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const int t = t0 + fg * 4 + f;
      if (t >= T) continue;
      float4 r = make_float4(ar[f][0], ar[f][1], ar[f][2], ar[f][3]), n = make_float4(an[f][0], an[f][1], an[f][2], an[f][3]);
      if (p.lda_bias) {
        r.x += bias.x, r.y += bias.y, r.z += bias.z, r.w += bias.w;
        n.x += bias.x, n.y += bias.y, n.z += bias.z, n.w += bias.w;
      }
      *reinterpret_cast<float4 *>(p.x_raw + (base + t) * p.ldim + j0) = r;
      *reinterpret_cast<float4 *>(p.x_norm + (base + t) * p.ldim + j0) = n;
    }
  }
}

// -------------------------------------------------------------------- UBM posteriors
// DiagGmm::LogLikelihoods (gmm/diag-gmm.cc:546-562) for a tile of kUbmFrames frames per CTA as a
// register-tiled product [frames x D] x [D x G] (the UBM tables stream through L2 once per tile,
// not once per frame), then VectorToPosteriorEntry (hmm/posterior.cc:440-508) with one warp per
// frame over the log-likelihood row kept in shared memory.
// (Measured in round 2: 32 frames per CTA, 8 frames x 8 consecutive Gaussians per thread with 16-byte table loads, one row
//  ahead in registers or eight rows ahead through shared memory with cp.async: 205-217 registers, one CTA per SM, 676 / 743 us
//  against 609 us -- the posterior selection below, one warp per frame, needs the occupancy more than the product needs the tile.)
// (32 frames per CTA with 8 frames per thread was measured: 169 registers, one CTA per SM, feature stage 2.31 -> 2.59 ms)
constexpr int kUbmFrames = 16;
__global__ void __launch_bounds__(256) ubm_post_kernel(IvecParams p) {
  extern __shared__ float sm[];
  const int D = p.ldim, G = p.num_gauss;
  float *sx = sm, *sxsq = sx + kUbmFrames * D, *sll = sxsq + kUbmFrames * D;  // [D][F], [D][F], [F][G]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kUbmFrames;
  const int nrows = min(kUbmFrames, p.total_frames - row0);
  // features transposed [d][frame]: the eight frames of a thread are two 16-byte broadcast loads
  for (int i = tid; i < kUbmFrames * D; i += 256) {
    const int f = i / D, d = i - f * D;
    const float v = f < nrows ? p.x_norm[(size_t)row0 * D + i] : 0.f;
    sx[d * kUbmFrames + f] = v;
    sxsq[d * kUbmFrames + f] = __fmul_rn(v, v);
  }
  __syncthreads();
  if ((G & 3) == 0) {
    // 8 frames x 4 CONSECUTIVE Gaussians per thread: one 16-byte load per table row and table, prefetched one row
    // ahead, and four 16-byte broadcast loads of the features feed 64 FMAs (round 1: 24 scalar loads per 64 FMAs, the
    // FMAs waited for the table loads).  Every sum still runs over d in ascending order with one FMA per term.
    const int tg = tid & 127, tf = tid >> 7;  // 128 groups of 4 Gaussians x 2 groups of 8 frames
    for (int g0 = 0; g0 < G; g0 += 512) {
      const int gb = g0 + tg * 4;
      const bool on = gb < G;
      float a1[8][4], a2[8][4];
#pragma unroll
      for (int f = 0; f < 8; f++)
#pragma unroll
        for (int j = 0; j < 4; j++) a1[f][j] = a2[f][j] = 0.f;
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 mn = on ? __ldg(reinterpret_cast<const float4 *>(p.means_invvars_t + gb)) : zero4;
      float4 vn = on ? __ldg(reinterpret_cast<const float4 *>(p.inv_vars_t + gb)) : zero4;
      for (int d = 0; d < D; d++) {
        const float4 m4 = mn, v4 = vn;
        if (d + 1 < D && on) {
          mn = __ldg(reinterpret_cast<const float4 *>(p.means_invvars_t + (size_t)(d + 1) * G + gb));
          vn = __ldg(reinterpret_cast<const float4 *>(p.inv_vars_t + (size_t)(d + 1) * G + gb));
        }
        const float4 xa = *reinterpret_cast<const float4 *>(sx + d * kUbmFrames + tf * 8), xb = *reinterpret_cast<const float4 *>(sx + d * kUbmFrames + tf * 8 + 4);
        const float4 qa = *reinterpret_cast<const float4 *>(sxsq + d * kUbmFrames + tf * 8), qb = *reinterpret_cast<const float4 *>(sxsq + d * kUbmFrames + tf * 8 + 4);
        const float xf[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w}, x2[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        const float m[4] = {m4.x, m4.y, m4.z, m4.w}, iv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int f = 0; f < 8; f++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            a1[f][j] = fmaf(xf[f], m[j], a1[f][j]);
            a2[f][j] = fmaf(x2[f], iv[j], a2[f][j]);
          }
      }
      if (on) {
        const float4 gc4 = __ldg(reinterpret_cast<const float4 *>(p.gconsts + gb));
        const float gc[4] = {gc4.x, gc4.y, gc4.z, gc4.w};
#pragma unroll
        for (int f = 0; f < 8; f++) {
          float4 o;
          o.x = __fadd_rn(__fadd_rn(gc[0], a1[f][0]), __fmul_rn(-0.5f, a2[f][0]));
          o.y = __fadd_rn(__fadd_rn(gc[1], a1[f][1]), __fmul_rn(-0.5f, a2[f][1]));
          o.z = __fadd_rn(__fadd_rn(gc[2], a1[f][2]), __fmul_rn(-0.5f, a2[f][2]));
          o.w = __fadd_rn(__fadd_rn(gc[3], a1[f][3]), __fmul_rn(-0.5f, a2[f][3]));
          *reinterpret_cast<float4 *>(sll + (size_t)(tf * 8 + f) * G + gb) = o;
        }
      }
    }
  } else {
    const int tg = tid & 63, tf = tid >> 6;  // 64 Gaussian lanes x 4 frame groups of 4 frames
    for (int g0 = 0; g0 < G; g0 += 512) {
      float a1[4][8], a2[4][8];
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int j = 0; j < 8; j++) a1[f][j] = a2[f][j] = 0.f;
      for (int d = 0; d < D; d++) {
        float m[8], iv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int g = g0 + tg + 64 * j;
          m[j] = g < G ? __ldg(p.means_invvars_t + (size_t)d * G + g) : 0.f;
          iv[j] = g < G ? __ldg(p.inv_vars_t + (size_t)d * G + g) : 0.f;
        }
#pragma unroll
        for (int f = 0; f < 4; f++) {
          const float xf = sx[d * kUbmFrames + tf * 4 + f], x2 = sxsq[d * kUbmFrames + tf * 4 + f];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            a1[f][j] = fmaf(xf, m[j], a1[f][j]);
            a2[f][j] = fmaf(x2, iv[j], a2[f][j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int g = g0 + tg + 64 * j;
        if (g < G) {
          const float gc = p.gconsts[g];
#pragma unroll
          for (int f = 0; f < 4; f++)
            sll[(tf * 4 + f) * G + g] = __fadd_rn(__fadd_rn(gc, a1[f][j]), __fmul_rn(-0.5f, a2[f][j]));
        }
      }
    }
  }
  __syncthreads();
  // GetMinPost with weight 1.0 (online-ivector-feature.cc:188-199)
  const float min_post = p.min_post > 0.99f ? 0.99f : p.min_post;
  for (int f = warp; f < nrows; f += 8) {
    const int row = row0 + f;
    float *ll = sll + (size_t)f * G;  // lane owns elements lane, lane + 32, ...
    float mx = -FLT_MAX;
    for (int g = lane; g < G; g += 32) mx = fmaxf(mx, ll[g]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float cut = min_post != 0.f ? __fadd_rn(mx, logf(min_post)) : -FLT_MAX;
    // posteriors of the candidates; everything else is marked -1
    int any = 0;
    for (int g = lane; g < G; g += 32) any |= (min_post != 0.f && ll[g] > cut) ? 1 : 0;
    any = __any_sync(0xffffffffu, any);
    for (int g = lane; g < G; g += 32) {
      const float v = ll[g];
      if (any)
        ll[g] = v > cut ? (float)exp((double)__fsub_rn(v, mx)) : -1.f;
      else  // min_post == 0 or nothing above the threshold: all Gaussians are candidates
        ll[g] = expf(__fsub_rn(v, mx));
    }
    // top num_gselect by posterior, in decreasing order (ties: lowest Gaussian index)
    const int ng = p.num_gselect < G ? p.num_gselect : G;
    float sel_post[8];
    int sel_idx[8];
    int nsel = 0;
    for (int s = 0; s < ng && s < 8; s++) {
      float best = -1.f;
      int bi = 0x7fffffff;
      for (int g = lane; g < G; g += 32)
        if (ll[g] > best) {
          best = ll[g];
          bi = g;
        }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
          best = ob;
          bi = oi;
        }
      }
      if (best < 0.f) break;
      sel_post[nsel] = best;
      sel_idx[nsel] = bi;
      nsel++;
      if ((bi & 31) == lane) ll[bi] = -1.f;
    }
    // prune + renormalise (posterior.cc:490-505), all lanes redundantly
    float tot = 0.f;
    for (int s = 0; s < nsel; s++) tot = __fadd_rn(tot, sel_post[s]);
    const float cutoff = __fmul_rn(min_post, tot);
    while (nsel > 1 && sel_post[nsel - 1] < cutoff) {
      tot = __fsub_rn(tot, sel_post[nsel - 1]);
      nsel--;
    }
    const float inv_tot = (float)(1.0 / (double)tot);
    const float sc = p.posterior_scale;  // * weight (1.0)
    if (lane < p.num_gselect) {
      int idx = -1;
      float w = 0.f;
      if (lane < nsel) {
        idx = sel_idx[lane];
        w = __fmul_rn(__fmul_rn(sel_post[lane], inv_tot), sc);
      }
      p.post_idx[(size_t)row * p.num_gselect + lane] = idx;
      p.post_w[(size_t)row * p.num_gselect + lane] = w;
    }
  }
}

// ---------------------------------------------------------------------------- statistics
// wf[u][g][:] += w * x_t  (double; order-insensitive at 1e-16) ; gw[u][g] = float sum over frames in
// frame order (GaussInfo::tot_weight is a float accumulated in frame order).
// Solves of one utterance are cumulative (solve j covers the first v_num_frames[j] frames, online schedule): every frame
// is accumulated ONCE, into the first solve that contains it, and ivec_prefix_kernel then adds the solves up in order.
// (Round 1 re-accumulated all frames of every solve: 3.6 of the 6.8 ms of device time of 64 concurrent streams.)
__global__ void __launch_bounds__(256) ivec_acc_kernel(IvecParams p) {
  const int u = blockIdx.y;  // solve index
  const int T = p.v_num_frames[u];
  const size_t base = (size_t)p.v_frame_offset[u];
  // frames the previous solve of the same utterance already covers (same feature rows <=> same utterance; an utterance
  // without frames shares its offset with the next one and contributes nothing either way)
  const int T0 = u > 0 && p.v_frame_offset[u - 1] == p.v_frame_offset[u] ? min(p.v_num_frames[u - 1], T) : 0;
  const float *feats = p.online_cmvn_iextractor ? p.x_norm : p.x_raw;
  const int D = p.ldim, S = p.num_gselect;
  double *wf = p.wf + (size_t)u * p.num_gauss * D;
  const int per = S * D;
  for (long long i = (long long)T0 * per + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)T * per;
       i += (long long)gridDim.x * blockDim.x) {
    int t = (int)(i / per), r = (int)(i - (long long)t * per);
    int s = r / D, d = r - s * D;
    int g = p.post_idx[(base + t) * S + s];
    if (g < 0) continue;
    double w = (double)p.post_w[(base + t) * S + s];
    atomicAdd(&wf[(size_t)g * D + d], w * (double)feats[(base + t) * D + d]);
  }
}

// wf[j] += wf[j - 1] over the solves of an utterance, in order (one thread per (g, d) element)
__global__ void __launch_bounds__(256) ivec_prefix_kernel(IvecParams p) {
  const int u = blockIdx.y;  // utterance
  const int j0 = p.v_begin[u], j1 = p.v_begin[u + 1];
  const size_t n = (size_t)p.num_gauss * p.ldim;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    double run = p.wf[(size_t)j0 * n + e];
    for (int j = j0 + 1; j < j1; j++) {
      run += p.wf[(size_t)j * n + e];
      p.wf[(size_t)j * n + e] = run;
    }
  }
}

// gw[j][g]: the float sum of the posteriors of Gaussian g over the first v_num_frames[j] frames IN FRAME ORDER
// (GaussInfo::tot_weight): one thread per Gaussian walks the utterance once and records the running sum at every
// solve boundary -- the same sequence of additions as a separate pass per solve.
__global__ void __launch_bounds__(256) ivec_gw_kernel(IvecParams p) {
  const int u = blockIdx.y;  // utterance
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int S = p.num_gselect;
  const int j0 = p.v_begin[u], j1 = p.v_begin[u + 1];
  if (j0 >= j1) return;
  const size_t base = (size_t)p.v_frame_offset[j0];
  extern __shared__ int spost[];  // tile of posteriors: idx then weight bits
  float acc = 0.f;
  const int tile = 256;
  int done = 0;  // posteriors consumed so far
  for (int j = j0; j < j1; j++) {
    const int upto = max(p.v_num_frames[j], 0) * S;
    while (done < upto) {
      const int n = min(tile, upto - done);
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        spost[i] = p.post_idx[base * S + done + i];
        spost[tile + i] = __float_as_int(p.post_w[base * S + done + i]);
      }
      __syncthreads();
      if (g < p.num_gauss)
        for (int i = 0; i < n; i++)
          if (spost[i] == g) acc = __fadd_rn(acc, __int_as_float(spost[tile + i]));
      done += n;
    }
    if (g < p.num_gauss) p.gw[(size_t)j * p.num_gauss + g] = acc;
  }
}

// linear[u][r] = prior + sum_g sum_d sigma_inv_m[g][d][r] * wf[u][g][d]
// quad[u][k]   = prior + sum_g gw[u][g] * U[g][k]
// The (g, d) sum is split over CTAs: CTA (c, group) adds kLinChunk terms for kUT utterances, reading
// its slice of the extractor table once, and writes the partial to linear_part[c][u][r]; the solve
// kernel folds the partials in a fixed order (deterministic, unlike atomics).
constexpr int kUT = 8;
constexpr int kLinChunk = 512;
// (Round 2, measured and dropped: the per-solve factors laid out [term][solve] and read as 16-byte broadcasts -- 155 -> 210 us
//  and 187 -> 211 us; sixteen solves per CTA -- 268 / 258 us, too few CTAs.)
__global__ void __launch_bounds__(128) ivec_linear_kernel(IvecParams p) {
  const int r = threadIdx.x;
  const int u0 = blockIdx.y * kUT;
  const int R = p.ivector_dim, GD = p.num_gauss * p.ldim;
  __shared__ double swf[kUT][kLinChunk];
  const int k0 = blockIdx.x * kLinChunk;
  const int n = min(kLinChunk, GD - k0);
  for (int i = threadIdx.x; i < kUT * n; i += blockDim.x) {
    int j = i / n, k = i - j * n;
    swf[j][k] = (u0 + j < p.v_n) ? p.wf[(size_t)(u0 + j) * GD + k0 + k] : 0.0;
  }
  __syncthreads();
  for (int rr = r; rr < R; rr += blockDim.x) {
    double acc[kUT];
#pragma unroll
    for (int j = 0; j < kUT; j++) acc[j] = 0.0;
    const double *m = p.sigma_inv_m + (size_t)k0 * R + rr;
#pragma unroll 4
    for (int k = 0; k < n; k++) {
      const double mv = m[(size_t)k * R];
#pragma unroll
      for (int j = 0; j < kUT; j++) acc[j] = fma(mv, swf[j][k], acc[j]);
    }
    for (int j = 0; j < kUT; j++)
      if (u0 + j < p.v_n) p.linear_part[((size_t)blockIdx.x * p.v_n + u0 + j) * R + rr] = acc[j];
  }
}

__global__ void __launch_bounds__(128) ivec_quad_kernel(IvecParams p) {
  const int R = p.ivector_dim, P = R * (R + 1) / 2, G = p.num_gauss;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int u0 = blockIdx.y * kUT;
  extern __shared__ double sgw[];  // [kUT][G]
  for (int i = threadIdx.x; i < kUT * G; i += blockDim.x) {
    int j = i / G, g = i - j * G;
    sgw[i] = (u0 + j < p.v_n) ? (double)p.gw[(size_t)(u0 + j) * G + g] : 0.0;
  }
  __syncthreads();
  if (k >= P) return;
  double acc[kUT];
#pragma unroll
  for (int j = 0; j < kUT; j++) acc[j] = 0.0;
  for (int g = 0; g < G; g++) {
    double uv = p.u[(size_t)g * P + k];
#pragma unroll
    for (int j = 0; j < kUT; j++) acc[j] = fma(sgw[j * G + g], uv, acc[j]);
  }
  for (int j = 0; j < kUT; j++)
    if (u0 + j < p.v_n) p.quad[(size_t)(u0 + j) * P + k] = acc[j];
}

// ------------------------------------------------------------------------------- CG solve
// One CTA per utterance: adds the prior terms (ivector-extractor.cc:786-798, 652-667), then
// LinearCgd in double with the reference's residual-recompute rule.
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
  return s;
}

__global__ void __launch_bounds__(128) ivec_cg_kernel(IvecParams p) {
  const int utt = blockIdx.x, R = p.ivector_dim, P = R * (R + 1) / 2;
  extern __shared__ double sh[];
  double *A = sh;            // [R*R] dense symmetric
  double *b = A + (size_t)R * R, *x = b + R, *r = x + R, *pv = r + R, *Ap = pv + R, *red = Ap + R;
  const int tid = threadIdx.x;
  // current_ivector_ of the reference: starts at the prior mean and is carried from solve to solve
  for (int i = tid; i < R; i += blockDim.x) x[i] = i == 0 ? p.prior_offset : 0.0;
  __syncthreads();
  for (int u = p.v_begin[utt]; u < p.v_begin[utt + 1]; u++) {  // the utterance's solves, in order
    // total weight (double sum of the float per-Gaussian totals)
    double tw = 0.0;
    for (int g = tid; g < p.num_gauss; g += blockDim.x) tw += (double)p.gw[(size_t)u * p.num_gauss + g];
    tw = block_sum(tw, red);
    double prior_scale_change = 0.0;
    if (p.max_count > 0.f) {
      double mc = (double)p.max_count;
      double old_scale = fmax(0.0, mc) / mc, new_scale = fmax(tw, mc) / mc;
      prior_scale_change = new_scale - old_scale;
    }
    const double *q = p.quad + (size_t)u * P;
    for (int i = tid; i < R * R; i += blockDim.x) {
      int a = i / R, c = i - a * R;
      int hi = a > c ? a : c, lo = a > c ? c : a;
      double v = q[(size_t)hi * (hi + 1) / 2 + lo];
      if (a == c) v += 1.0 + prior_scale_change;
      A[i] = v;
    }
    for (int i = tid; i < R; i += blockDim.x) {
      double v = 0.0;  // fold the split-K partials of ivec_linear_kernel in chunk order
      for (int c = 0; c < p.linear_chunks; c++) v += p.linear_part[((size_t)c * p.v_n + u) * R + i];
      if (i == 0) v += p.prior_offset + p.prior_offset * prior_scale_change;
      b[i] = v;
    }
    __syncthreads();
    const bool have_data = tw > 0.0;
    auto matvec = [&](const double *v, double *out) {  // out = A v
      for (int i = tid; i < R; i += blockDim.x) {
        double acc = 0.0;
        const double *row = A + (size_t)i * R;
        for (int k = 0; k < R; k++) acc = fma(row[k], v[k], acc);
        out[i] = acc;
      }
      __syncthreads();
    };
    auto dot = [&](const double *a, const double *c) {
      double v = 0.0;
      for (int i = tid; i < R; i += blockDim.x) v += a[i] * c[i];
      return block_sum(v, red);
    };
    if (have_data) {
      // GetIvector (ivector-extractor.cc:732-756): warm start, "better initial guess" if x[0] == 0
      if (tid == 0 && x[0] == 0.0) x[0] = p.prior_offset;
      __syncthreads();
      matvec(x, Ap);
      for (int i = tid; i < R; i += blockDim.x) {
        pv[i] = b[i] - Ap[i];
        r[i] = -pv[i];
      }
      __syncthreads();
      double r_cur = dot(r, r), r_recompute = r_cur;
      const double max_err_sq = DBL_MIN, rf = 0.01 * 0.01, inv_rf = 1.0 / rf;
      for (int k = 0; k < R + 5 && k != p.num_cg_iters; k++) {
        matvec(pv, Ap);
        double alpha = -dot(pv, r) / dot(pv, Ap);
        for (int i = tid; i < R; i += blockDim.x) {
          x[i] += alpha * pv[i];
          r[i] += alpha * Ap[i];
        }
        __syncthreads();
        double r_next = dot(r, r);
        if (r_next < rf * r_recompute || r_next > inv_rf * r_recompute) {
          matvec(x, Ap);
          for (int i = tid; i < R; i += blockDim.x) r[i] = Ap[i] - b[i];
          __syncthreads();
          r_next = dot(r, r);
          r_recompute = r_next;
        }
        if (r_next <= max_err_sq) break;
        double beta = r_next / r_cur;
        for (int i = tid; i < R; i += blockDim.x) pv[i] = beta * pv[i] - r[i];
        __syncthreads();
        r_cur = r_next;
      }
    } else {
      for (int i = tid; i < R; i += blockDim.x) x[i] = i == 0 ? p.prior_offset : 0.0;
      __syncthreads();
    }
    // nnet input: float copy, prior offset removed from the first element
    // (online-ivector-feature.cc:344-347)
    for (int i = tid; i < R; i += blockDim.x) {
      float f = (float)x[i];
      if (i == 0) f = (float)((double)f - p.prior_offset);
      p.ivector[(size_t)u * p.ivector_ld + i] = f;
    }
    __syncthreads();
  }
}

void LaunchIvector(const IvecParams &p, cudaStream_t stream) {
  if (p.n_utts == 0) return;
  const int G = p.num_gauss, D = p.ldim, R = p.ivector_dim, P = R * (R + 1) / 2;
  if (p.total_frames > 0) {
    const int K = p.dim * (p.left + 1 + p.right);
    if ((p.dim & 3) == 0 && (p.ldim & 3) == 0 && (K & 3) == 0) {
      dim3 g1((p.max_frames + kLdaFrames4 - 1) / kLdaFrames4, p.n_utts);
      size_t sm1 = (size_t)2 * (kLdaFrames4 + p.left + p.right) * p.dim * sizeof(float);
      splice_lda4_kernel<<<g1, 256, sm1, stream>>>(p);
      RS_CHECK_LAUNCH("splice_lda4_kernel");
    } else {
      dim3 g1((p.max_frames + kLdaFrames - 1) / kLdaFrames, p.n_utts);
      size_t sm1 = (size_t)2 * (kLdaFrames + p.left + p.right) * p.dim * sizeof(float);
      splice_lda_kernel<<<g1, 256, sm1, stream>>>(p);
      RS_CHECK_LAUNCH("splice_lda_kernel");
    }
    size_t sm2 = (size_t)kUbmFrames * (2 * D + G) * sizeof(float);
    EnsureDynSmem(ubm_post_kernel, sm2);
    ubm_post_kernel<<<(p.total_frames + kUbmFrames - 1) / kUbmFrames, 256, sm2, stream>>>(p);
    RS_CHECK_LAUNCH("ubm_post_kernel");
  }
  cudaMemsetAsync(p.wf, 0, (size_t)p.v_n * G * D * sizeof(double), stream);
  if (p.total_frames > 0) {
    int per_utt_blocks = (p.max_frames * p.num_gselect * D + 255) / 256;
    if (per_utt_blocks > 64) per_utt_blocks = 64;
    if (per_utt_blocks < 1) per_utt_blocks = 1;
    ivec_acc_kernel<<<dim3(per_utt_blocks, p.v_n), 256, 0, stream>>>(p);
    RS_CHECK_LAUNCH("ivec_acc_kernel");
  }
  if (p.v_n > p.n_utts) {  // online schedule: several cumulative solves per utterance
    ivec_prefix_kernel<<<dim3((G * D + 255) / 256, p.n_utts), 256, 0, stream>>>(p);
    RS_CHECK_LAUNCH("ivec_prefix_kernel");
  }
  ivec_gw_kernel<<<dim3((G + 255) / 256, p.n_utts), 256, 2 * 256 * sizeof(int), stream>>>(p);
  RS_CHECK_LAUNCH("ivec_gw_kernel");
  int groups = (p.v_n + kUT - 1) / kUT;
  const size_t sm_quad = (size_t)kUT * G * sizeof(double);
  EnsureDynSmem(ivec_quad_kernel, sm_quad);
  ivec_linear_kernel<<<dim3(p.linear_chunks, groups), 128, 0, stream>>>(p);
  RS_CHECK_LAUNCH("ivec_linear_kernel");
  ivec_quad_kernel<<<dim3((P + 127) / 128, groups), 128, sm_quad, stream>>>(p);
  RS_CHECK_LAUNCH("ivec_quad_kernel");
  size_t sm3 = ((size_t)R * R + 5 * R + 8) * sizeof(double);
  EnsureDynSmem(ivec_cg_kernel, sm3);
  ivec_cg_kernel<<<p.n_utts, 128, sm3, stream>>>(p);
  RS_CHECK_LAUNCH("ivec_cg_kernel");
}

}  // namespace rs
