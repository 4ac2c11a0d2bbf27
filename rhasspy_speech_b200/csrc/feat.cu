// Stage (i): MFCC feature extraction, one warp per frame.
//
// Replaces OnlineGenericBaseFeature<MfccComputer>::ComputeFeatures and everything below it:
//   kaldi/src/feat/feature-window.cc:137-224  (ExtractWindow / ProcessWindow: DC removal,
//                                              pre-emphasis, povey window, zero padding)
//   kaldi/src/matrix/srfft.cc:207-440         (split-radix complex FFT + real-FFT unpacking)
//   kaldi/src/feat/feature-functions.cc:29-51 (power spectrum)
//   kaldi/src/feat/mel-computations.cc:226-251, feature-mfcc.cc:28-80 (mel, log, DCT, lifter)
//
// The FFT executes the reference's split-radix butterfly network -- same operands, same
// operation order, no FMA contraction -- but level by level across the lanes of a warp instead
// of recursively, so the packed spectrum is bit-identical to the CPU path.  Twiddle tables are
// built on the host with the reference's float expressions (feat_tables.cc).
#include "engine.h"

namespace rs {

// exact float helpers: never contracted into FMA
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

constexpr int kWarpsPerCta = 8;
constexpr int kFramesPerWarp = 2;

// One warp computes kFramesPerWarp consecutive frames of one utterance side by side: block offsets, twiddle factors,
// the bit-reversal permutation, window, mel and DCT tables are looked up once and applied to every frame of the warp
// (a third of the one-frame kernel's instructions was index arithmetic and table loads of the level-by-level
// split-radix replay).  Per frame the operations and their order are unchanged.
// smem per warp and frame: xr[N_], xi[N_] (N_ = padded/2) + mel[num_bins].
__global__ void __launch_bounds__(kWarpsPerCta * 32)
mfcc_kernel(FeatParams p) {
  constexpr int F = kFramesPerWarp;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.y;
  const int T = p.num_frames[u];
  const int t0 = (blockIdx.x * kWarpsPerCta + warp) * F;
  if (t0 >= T) return;  // whole warp exits together; only __syncwarp is used below
  const int N = p.padded, NH = N >> 1;
  const int fstride = N + p.num_bins + 8;  // floats per frame
  float *xr = smem + (size_t)warp * F * fstride;
  float *xi = xr + NH;
  float *melv = xr + N;
  const int L = p.length;
  // a warp whose last frames lie beyond the utterance computes its last valid frame again and does not store it
  int tf[F];
  const int16_t *pcm[F];
#pragma unroll
  for (int f = 0; f < F; f++) {
    tf[f] = t0 + f < T ? t0 + f : T - 1;
    pcm[f] = p.pcm + p.pcm_offset[u] + (size_t)tf[f] * p.shift;
  }

  // --- window: samples -> float, remove DC (the int16 sum is exact in float), pre-emphasis, window
  // Dither (feature-window.cc:90-98) draws from the C library RNG in the reference; here a counter
  // hash drives a Box-Muller draw: statistically equivalent, not bit-comparable (parity runs use
  // --dither=0, as the reference's own feature tests do, online-feature-test.cc:155).
  auto sample = [&](int f, int i) -> float {
    float x = (float)pcm[f][i];
    if (p.dither != 0.f) {
      const uint32_t dbase = (uint32_t)(p.pcm_offset[u] + (size_t)tf[f] * p.shift);
      uint32_t h1 = hash_u32((dbase + (uint32_t)i) * 2u + 0x9e3779b9u * (p.seed + 1u));
      uint32_t h2 = hash_u32(h1 ^ 0x85ebca6bu);
      float u1 = ((h1 >> 8) + 1.0f) * (1.0f / 16777217.0f), u2 = (h2 >> 8) * (1.0f / 16777216.0f);
      x += sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2) * p.dither;
    }
    return x;
  };
  float dc[F];
  float log_energy[F];
#pragma unroll
  for (int f = 0; f < F; f++) {
    float *xrf = xr + f * fstride, *xif = xi + f * fstride;
    // every lane keeps its samples (i = lane + 32 j) in registers: one load and one conversion per sample
    // (frames of up to 512 samples; the buffers below already assume padded <= 512)
    float sv[16];
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int i = lane + 32 * j;
      sv[j] = i < L ? sample(f, i) : 0.f;
      if (i < L) part += sv[j];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    dc[f] = 0.f;
    if (p.remove_dc) dc[f] = -part / (float)L;
    // interleaved complex input: sample 2j -> xr[j], sample 2j+1 -> xi[j]  (srfft.cc:147-156)
    const float pre = p.preemph;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int i = lane + 32 * j;
      // the previous sample sits in the lane below, or for lane 0 in lane 31 one row up (sample 0 is its own predecessor)
      float prev = __shfl_up_sync(0xffffffffu, sv[j], 1);
      const float wrap = __shfl_sync(0xffffffffu, sv[j > 0 ? j - 1 : 0], 31);
      if (lane == 0) prev = j > 0 ? wrap : sv[0];
      if (i < N) {
        float v = 0.f;
        if (i < L) {
          float x = fadd(sv[j], dc[f]);
          const float xm = fadd(prev, dc[f]);
          if (pre != 0.f) x = fsub(x, fmul(pre, xm));
          v = fmul(x, p.window[i]);
        }
        if (i & 1) xif[i >> 1] = v; else xrf[i >> 1] = v;
      }
    }
  }
  __syncwarp();
  // raw log-energy is only needed with --use-energy=true
#pragma unroll
  for (int f = 0; f < F; f++) {
    log_energy[f] = 0.f;
    if (p.use_energy) {
      const float *xrf = xr + f * fstride, *xif = xi + f * fstride;
      // energy of the frame after DC removal (raw_energy) or after windowing (feature-mfcc.cc:38-40)
      float e = 0.f;
      for (int i = lane; i < L; i += 32) {
        float x;
        if (p.raw_energy) {
          x = fadd(sample(f, i), dc[f]);
        } else {
          x = (i & 1) ? xif[i >> 1] : xrf[i >> 1];
        }
        e += x * x;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
      log_energy[f] = logf(fmaxf(e, 1.1920928955078125e-07f));
      if (p.energy_floor > 0.f) log_energy[f] = fmaxf(log_energy[f], logf(p.energy_floor));
    }
  }

  // --- split-radix complex FFT of size NH, level by level (srfft.cc:207-345)
  const int logn = p.logn;  // log2(NH)
  for (int lv = logn; lv >= 3; lv--) {
    const int m = 1 << lv, m2 = m >> 1, m4 = m >> 2, m8 = m >> 3;
    const int nb = p.level_count[lv];
    const uint16_t *offs = p.level_offsets + p.level_start[lv];
    // step 1
    for (int i = lane; i < nb * m2; i += 32) {
      const int o = offs[i >> (lv - 1)] + (i & (m2 - 1));
#pragma unroll
      for (int f = 0; f < F; f++) {
        float *xrf = xr + f * fstride, *xif = xi + f * fstride;
        float a = xrf[o], b = xrf[o + m2];
        xrf[o] = fadd(a, b);
        xrf[o + m2] = fsub(a, b);
        a = xif[o];
        b = xif[o + m2];
        xif[o] = fadd(a, b);
        xif[o + m2] = fsub(a, b);
      }
    }
    __syncwarp();
    // steps 2, 3 and 4 touch the same four values for a given n
    const float *tab = p.twiddle + p.twiddle_start[lv];
    const int nel = m4 - 2;
    for (int i = lane; i < nb * m4; i += 32) {
      const int n = i & (m4 - 1);
      const int o = offs[i >> (lv - 2)] + m2 + n;
      float cn = 0.f, spcn = 0.f, smcn = 0.f, c3n = 0.f, spc3n = 0.f, smc3n = 0.f;
      const bool general = n != 0 && n != m8;
      if (general) {
        const int k = n < m8 ? n - 1 : n - 2;
        cn = tab[k]; spcn = tab[nel + k]; smcn = tab[2 * nel + k];
        c3n = tab[3 * nel + k]; spc3n = tab[4 * nel + k]; smc3n = tab[5 * nel + k];
      }
#pragma unroll
      for (int f = 0; f < F; f++) {
        float *xrf = xr + f * fstride, *xif = xi + f * fstride;
        float r1 = xrf[o], r2 = xrf[o + m4], i1 = xif[o], i2 = xif[o + m4];
        float t1 = fadd(r1, i2), t2 = fadd(i1, r2);
        i1 = fsub(i1, r2);
        r2 = fsub(r1, i2);
        r1 = t1;
        i2 = t2;
        if (n == 0) {
          // no twiddle
        } else if (n == m8) {
          const float sq = 0.70710678118654752440f;
          t1 = fmul(sq, fadd(r1, i1));
          i1 = fmul(sq, fsub(i1, r1));
          r1 = t1;
          t2 = fmul(sq, fsub(i2, r2));
          i2 = fmul(-sq, fadd(r2, i2));
          r2 = t2;
        } else {
          t2 = fmul(cn, fadd(r1, i1));
          t1 = fadd(fmul(spcn, r1), t2);
          r1 = fadd(fmul(smcn, i1), t2);
          i1 = t1;
          t2 = fmul(c3n, fadd(r2, i2));
          t1 = fadd(fmul(spc3n, r2), t2);
          r2 = fadd(fmul(smc3n, i2), t2);
          i2 = t1;
        }
        xrf[o] = r1;
        xrf[o + m4] = r2;
        xif[o] = i1;
        xif[o + m4] = i2;
      }
    }
    __syncwarp();
  }
  {  // length-4 blocks (srfft.cc:223-266)
    const int nb = p.level_count[2];
    const uint16_t *offs = p.level_offsets + p.level_start[2];
    for (int i = lane; i < nb; i += 32) {
      const int o = offs[i];
#pragma unroll
      for (int f = 0; f < F; f++) {
        float *xrf = xr + f * fstride, *xif = xi + f * fstride;
        float r0 = xrf[o], r1 = xrf[o + 1], r2 = xrf[o + 2], r3 = xrf[o + 3];
        float i0 = xif[o], i1 = xif[o + 1], i2 = xif[o + 2], i3 = xif[o + 3];
        float t;
        t = fadd(r0, r2); r2 = fsub(r0, r2); r0 = t;
        t = fadd(i0, i2); i2 = fsub(i0, i2); i0 = t;
        t = fadd(r1, r3); r3 = fsub(r1, r3); r1 = t;
        t = fadd(i1, i3); i3 = fsub(i1, i3); i1 = t;
        t = fadd(r0, r1); r1 = fsub(r0, r1); r0 = t;
        t = fadd(i0, i1); i1 = fsub(i0, i1); i0 = t;
        float t1 = fadd(r2, i3), t2 = fadd(i2, r3);
        i2 = fsub(i2, r3);
        r3 = fsub(r2, i3);
        r2 = t1;
        i3 = t2;
        xrf[o] = r0; xrf[o + 1] = r1; xrf[o + 2] = r2; xrf[o + 3] = r3;
        xif[o] = i0; xif[o + 1] = i1; xif[o + 2] = i2; xif[o + 3] = i3;
      }
    }
  }
  {  // length-2 blocks (srfft.cc:268-277)
    const int nb = p.level_count[1];
    const uint16_t *offs = p.level_offsets + p.level_start[1];
    for (int i = lane; i < nb; i += 32) {
      const int o = offs[i];
#pragma unroll
      for (int f = 0; f < F; f++) {
        float *xrf = xr + f * fstride, *xif = xi + f * fstride;
        float a = xrf[o], b = xrf[o + 1];
        xrf[o] = fadd(a, b);
        xrf[o + 1] = fsub(a, b);
        a = xif[o];
        b = xif[o + 1];
        xif[o] = fadd(a, b);
        xif[o + 1] = fsub(a, b);
      }
    }
  }
  __syncwarp();

  // --- real-FFT unpacking (srfft.cc:356-420) fused with the power spectrum; B_k = complex FFT
  // output k, read through the bit-reversal permutation.  Results go to a second view of smem:
  // power[k] for k in [0, NH] is written after all reads of this pass are done.
  float pw[F][9];  // up to NH/32 + 1 values per lane (NH <= 256)
  float pw2[F][9];
  {
    int npw = 0;
    for (int k = lane; k <= NH / 2; k += 32, npw++) {
      if (k == 0) {
        const int j = p.perm[0];
#pragma unroll
        for (int f = 0; f < F; f++) {
          const float d0 = xr[f * fstride + j], d1 = xi[f * fstride + j];
          const float z = fadd(d0, d1);
          pw[f][npw] = fmul(z, z);
        }
      } else {
        const int ja = p.perm[k], jb = p.perm[NH - k];
        const float kre = p.kn[2 * (k - 1)], kim = p.kn[2 * (k - 1) + 1];
#pragma unroll
        for (int f = 0; f < F; f++) {
          const float *xrf = xr + f * fstride, *xif = xi + f * fstride;
          float a_re = xrf[ja], a_im = xif[ja], b_re = xrf[jb], b_im = xif[jb];
          float ck_re = fmul(0.5f, fadd(a_re, b_re));
          float ck_im = fmul(0.5f, fsub(a_im, b_im));
          float dk_re = fmul(0.5f, fadd(a_im, b_im));
          float dk_im = fmul(-0.5f, fsub(a_re, b_re));
          float re = fadd(ck_re, fsub(fmul(kre, dk_re), fmul(kim, dk_im)));
          float im = fadd(ck_im, fadd(fmul(kre, dk_im), fmul(kim, dk_re)));
          pw[f][npw] = fadd(fmul(re, re), fmul(im, im));
        }
      }
    }
    int npw2 = 0;
    for (int k = lane; k < NH / 2; k += 32, npw2++) {
      // index NH - k (k >= 1), plus the Nyquist bin for k == 0
      if (k == 0) {
        const int j = p.perm[0];
#pragma unroll
        for (int f = 0; f < F; f++) {
          const float d0 = xr[f * fstride + j], d1 = xi[f * fstride + j];
          const float z = fsub(d0, d1);
          pw2[f][npw2] = fmul(z, z);
        }
      } else {
        const int ja = p.perm[k], jb = p.perm[NH - k];
        const float kre = p.kn[2 * (k - 1)], kim = p.kn[2 * (k - 1) + 1];
#pragma unroll
        for (int f = 0; f < F; f++) {
          const float *xrf = xr + f * fstride, *xif = xi + f * fstride;
          float a_re = xrf[ja], a_im = xif[ja], b_re = xrf[jb], b_im = xif[jb];
          float ck_re = fmul(0.5f, fadd(a_re, b_re));
          float ck_im = fmul(0.5f, fsub(a_im, b_im));
          float dk_re = fmul(0.5f, fadd(a_im, b_im));
          float dk_im = fmul(-0.5f, fsub(a_re, b_re));
          float ndk_im = -dk_im, nkre = -kre;
          float re = fadd(ck_re, fsub(fmul(nkre, dk_re), fmul(kim, ndk_im)));
          float im = fadd(-ck_im, fadd(fmul(nkre, ndk_im), fmul(kim, dk_re)));
          pw2[f][npw2] = fadd(fmul(re, re), fmul(im, im));
        }
      }
    }
  }
  __syncwarp();
  // power spectrum into xr[0 .. NH] (xr has NH entries, xi follows contiguously)
  {
    int npw = 0;
    for (int k = lane; k <= NH / 2; k += 32, npw++)
#pragma unroll
      for (int f = 0; f < F; f++) xr[f * fstride + k] = pw[f][npw];
    int npw2 = 0;
    for (int k = lane; k < NH / 2; k += 32, npw2++)
#pragma unroll
      for (int f = 0; f < F; f++) xr[f * fstride + NH - k] = pw2[f][npw2];
  }
  __syncwarp();

  // --- mel filterbank + log (mel-computations.cc:226-251, feature-mfcc.cc:57-58)
  // (weights read tap-major: the lanes of a warp -- one bin each -- read consecutive addresses; same taps, same order)
  for (int b = lane; b < p.num_bins; b += 32) {
    const int off = p.mel_offset[b], len = p.mel_len[b];
    const float *w = p.mel_weights_t + b;
    float e[F];
#pragma unroll
    for (int f = 0; f < F; f++) e[f] = 0.f;
    for (int i = 0; i < len; i++) {
      const float wi = w[(size_t)i * p.num_bins];
#pragma unroll
      for (int f = 0; f < F; f++) e[f] = fmaf(wi, xr[f * fstride + off + i], e[f]);
    }
#pragma unroll
    for (int f = 0; f < F; f++) melv[f * fstride + b] = logf(fmaxf(e[f], 1.1920928955078125e-07f));
  }
  __syncwarp();
  // --- DCT + lifter (feature-mfcc.cc:61-66)
  for (int c = lane; c < p.num_ceps; c += 32) {
    const float *col = p.dct_t + c;  // transposed: the lanes (one cepstrum each) read consecutive addresses
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; f++) acc[f] = 0.f;
    for (int b = 0; b < p.num_bins; b++) {
      const float cb = col[(size_t)b * p.num_ceps];
#pragma unroll
      for (int f = 0; f < F; f++) acc[f] = fmaf(cb, melv[f * fstride + b], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < F; f++) {
      float a = acc[f];
      if (p.lifter) a = fmul(a, p.lifter[c]);
      if (p.use_energy && c == 0) a = log_energy[f];
      if (t0 + f < T) p.mfcc[((size_t)p.frame_offset[u] + t0 + f) * p.num_ceps + c] = a;
    }
  }
}

void LaunchMfcc(const FeatParams &p, int n_utts, int max_frames, cudaStream_t stream) {
  if (n_utts == 0 || max_frames == 0) return;
  const int per_cta = kWarpsPerCta * kFramesPerWarp;
  dim3 grid((max_frames + per_cta - 1) / per_cta, n_utts);
  size_t smem = (size_t)per_cta * (p.padded + p.num_bins + 8) * sizeof(float);
  mfcc_kernel<<<grid, kWarpsPerCta * 32, smem, stream>>>(p);
}

}  // namespace rs
