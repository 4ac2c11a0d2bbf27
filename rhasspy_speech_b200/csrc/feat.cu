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

// One warp computes one frame.  smem per warp: xr[N_], xi[N_] (N_ = padded/2) + mel[num_bins].
__global__ void __launch_bounds__(kWarpsPerCta * 32)
mfcc_kernel(FeatParams p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.y;
  const int T = p.num_frames[u];
  const int t = blockIdx.x * kWarpsPerCta + warp;
  if (t >= T) return;  // whole warp exits together; only __syncwarp is used below
  const int N = p.padded, NH = N >> 1;
  float *xr = smem + (size_t)warp * (N + p.num_bins + 8);
  float *xi = xr + NH;
  float *melv = xr + N;
  const int16_t *pcm = p.pcm + p.pcm_offset[u] + (size_t)t * p.shift;
  const int L = p.length;

  // --- window: samples -> float, remove DC (the int16 sum is exact in float), pre-emphasis, window
  // Dither (feature-window.cc:90-98) draws from the C library RNG in the reference; here a counter
  // hash drives a Box-Muller draw: statistically equivalent, not bit-comparable (parity runs use
  // --dither=0, as the reference's own feature tests do, online-feature-test.cc:155).
  const uint32_t dbase = (uint32_t)(p.pcm_offset[u] + (size_t)t * p.shift);
  auto sample = [&](int i) -> float {
    float x = (float)pcm[i];
    if (p.dither != 0.f) {
      uint32_t h1 = hash_u32((dbase + (uint32_t)i) * 2u + 0x9e3779b9u * (p.seed + 1u));
      uint32_t h2 = hash_u32(h1 ^ 0x85ebca6bu);
      float u1 = ((h1 >> 8) + 1.0f) * (1.0f / 16777217.0f), u2 = (h2 >> 8) * (1.0f / 16777216.0f);
      x += sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2) * p.dither;
    }
    return x;
  };
  // every lane keeps its samples (i = lane + 32 j) in registers: one load and one conversion per sample
  // (frames of up to 512 samples; the buffers below already assume padded <= 512)
  float sv[16];
  float part = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int i = lane + 32 * j;
    sv[j] = i < L ? sample(i) : 0.f;
    if (i < L) part += sv[j];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  float dc = 0.f;
  if (p.remove_dc) dc = -part / (float)L;
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
        float x = fadd(sv[j], dc);
        const float xm = fadd(prev, dc);
        if (pre != 0.f) x = fsub(x, fmul(pre, xm));
        v = fmul(x, p.window[i]);
      }
      if (i & 1) xi[i >> 1] = v; else xr[i >> 1] = v;
    }
  }
  __syncwarp();
  // raw log-energy is only needed with --use-energy=true
  float log_energy = 0.f;
  if (p.use_energy) {
    // energy of the frame after DC removal (raw_energy) or after windowing (feature-mfcc.cc:38-40)
    float e = 0.f;
    for (int i = lane; i < L; i += 32) {
      float x;
      if (p.raw_energy) {
        x = fadd(sample(i), dc);
      } else {
        x = (i & 1) ? xi[i >> 1] : xr[i >> 1];
      }
      e += x * x;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    log_energy = logf(fmaxf(e, 1.1920928955078125e-07f));
    if (p.energy_floor > 0.f) log_energy = fmaxf(log_energy, logf(p.energy_floor));
  }

  // --- split-radix complex FFT of size NH, level by level (srfft.cc:207-345)
  const int logn = p.logn;  // log2(NH)
  for (int lv = logn; lv >= 3; lv--) {
    const int m = 1 << lv, m2 = m >> 1, m4 = m >> 2, m8 = m >> 3;
    const int nb = p.level_count[lv];
    const uint16_t *offs = p.level_offsets + p.level_start[lv];
    // step 1
    for (int i = lane; i < nb * m2; i += 32) {
      int o = offs[i >> (lv - 1)] + (i & (m2 - 1));
      float a = xr[o], b = xr[o + m2];
      xr[o] = fadd(a, b);
      xr[o + m2] = fsub(a, b);
      a = xi[o];
      b = xi[o + m2];
      xi[o] = fadd(a, b);
      xi[o + m2] = fsub(a, b);
    }
    __syncwarp();
    // steps 2, 3 and 4 touch the same four values for a given n
    const float *tab = p.twiddle + p.twiddle_start[lv];
    const int nel = m4 - 2;
    for (int i = lane; i < nb * m4; i += 32) {
      int n = i & (m4 - 1);
      int o = offs[i >> (lv - 2)] + m2 + n;
      float r1 = xr[o], r2 = xr[o + m4], i1 = xi[o], i2 = xi[o + m4];
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
        int k = n < m8 ? n - 1 : n - 2;
        float cn = tab[k], spcn = tab[nel + k], smcn = tab[2 * nel + k];
        float c3n = tab[3 * nel + k], spc3n = tab[4 * nel + k], smc3n = tab[5 * nel + k];
        t2 = fmul(cn, fadd(r1, i1));
        t1 = fadd(fmul(spcn, r1), t2);
        r1 = fadd(fmul(smcn, i1), t2);
        i1 = t1;
        t2 = fmul(c3n, fadd(r2, i2));
        t1 = fadd(fmul(spc3n, r2), t2);
        r2 = fadd(fmul(smc3n, i2), t2);
        i2 = t1;
      }
      xr[o] = r1;
      xr[o + m4] = r2;
      xi[o] = i1;
      xi[o + m4] = i2;
    }
    __syncwarp();
  }
  {  // length-4 blocks (srfft.cc:223-266)
    const int nb = p.level_count[2];
    const uint16_t *offs = p.level_offsets + p.level_start[2];
    for (int i = lane; i < nb; i += 32) {
      int o = offs[i];
      float r0 = xr[o], r1 = xr[o + 1], r2 = xr[o + 2], r3 = xr[o + 3];
      float i0 = xi[o], i1 = xi[o + 1], i2 = xi[o + 2], i3 = xi[o + 3];
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
      xr[o] = r0; xr[o + 1] = r1; xr[o + 2] = r2; xr[o + 3] = r3;
      xi[o] = i0; xi[o + 1] = i1; xi[o + 2] = i2; xi[o + 3] = i3;
    }
  }
  {  // length-2 blocks (srfft.cc:268-277)
    const int nb = p.level_count[1];
    const uint16_t *offs = p.level_offsets + p.level_start[1];
    for (int i = lane; i < nb; i += 32) {
      int o = offs[i];
      float a = xr[o], b = xr[o + 1];
      xr[o] = fadd(a, b);
      xr[o + 1] = fsub(a, b);
      a = xi[o];
      b = xi[o + 1];
      xi[o] = fadd(a, b);
      xi[o + 1] = fsub(a, b);
    }
  }
  __syncwarp();

  // --- real-FFT unpacking (srfft.cc:356-420) fused with the power spectrum; B_k = complex FFT
  // output k, read through the bit-reversal permutation.  Results go to a second view of smem:
  // power[k] for k in [0, NH] is written after all reads of this pass are done.
  float pw[9];  // up to NH/32 + 1 values per lane (NH <= 256)
  int npw = 0;
  for (int k = lane; k <= NH / 2; k += 32) {
    float out;
    if (k == 0) {
      int j = p.perm[0];
      float d0 = xr[j], d1 = xi[j];
      float z = fadd(d0, d1);
      out = fmul(z, z);
    } else {
      int ja = p.perm[k], jb = p.perm[NH - k];
      float a_re = xr[ja], a_im = xi[ja], b_re = xr[jb], b_im = xi[jb];
      float kre = p.kn[2 * (k - 1)], kim = p.kn[2 * (k - 1) + 1];
      float ck_re = fmul(0.5f, fadd(a_re, b_re));
      float ck_im = fmul(0.5f, fsub(a_im, b_im));
      float dk_re = fmul(0.5f, fadd(a_im, b_im));
      float dk_im = fmul(-0.5f, fsub(a_re, b_re));
      float re = fadd(ck_re, fsub(fmul(kre, dk_re), fmul(kim, dk_im)));
      float im = fadd(ck_im, fadd(fmul(kre, dk_im), fmul(kim, dk_re)));
      out = fadd(fmul(re, re), fmul(im, im));
    }
    pw[npw++] = out;
  }
  float pw2[9];
  int npw2 = 0;
  for (int k = lane; k < NH / 2; k += 32) {
    // index NH - k (k >= 1), plus the Nyquist bin for k == 0
    float out;
    if (k == 0) {
      int j = p.perm[0];
      float d0 = xr[j], d1 = xi[j];
      float z = fsub(d0, d1);
      out = fmul(z, z);
    } else {
      int ja = p.perm[k], jb = p.perm[NH - k];
      float a_re = xr[ja], a_im = xi[ja], b_re = xr[jb], b_im = xi[jb];
      float kre = p.kn[2 * (k - 1)], kim = p.kn[2 * (k - 1) + 1];
      float ck_re = fmul(0.5f, fadd(a_re, b_re));
      float ck_im = fmul(0.5f, fsub(a_im, b_im));
      float dk_re = fmul(0.5f, fadd(a_im, b_im));
      float dk_im = fmul(-0.5f, fsub(a_re, b_re));
      float ndk_im = -dk_im, nkre = -kre;
      float re = fadd(ck_re, fsub(fmul(nkre, dk_re), fmul(kim, ndk_im)));
      float im = fadd(-ck_im, fadd(fmul(nkre, ndk_im), fmul(kim, dk_re)));
      out = fadd(fmul(re, re), fmul(im, im));
    }
    pw2[npw2++] = out;
  }
  __syncwarp();
  // power spectrum into xr[0 .. NH] (xr has NH entries, xi follows contiguously)
  npw = 0;
  for (int k = lane; k <= NH / 2; k += 32) xr[k] = pw[npw++];
  npw2 = 0;
  for (int k = lane; k < NH / 2; k += 32) xr[NH - k] = pw2[npw2++];
  __syncwarp();

  // --- mel filterbank + log (mel-computations.cc:226-251, feature-mfcc.cc:57-58)
  // (weights read tap-major: the lanes of a warp -- one bin each -- read consecutive addresses; same taps, same order)
  for (int b = lane; b < p.num_bins; b += 32) {
    int off = p.mel_offset[b], len = p.mel_len[b];
    const float *w = p.mel_weights_t + b;
    float e = 0.f;
    for (int i = 0; i < len; i++) e = fmaf(w[(size_t)i * p.num_bins], xr[off + i], e);
    melv[b] = logf(fmaxf(e, 1.1920928955078125e-07f));
  }
  __syncwarp();
  // --- DCT + lifter (feature-mfcc.cc:61-66)
  float *out = p.mfcc + ((size_t)p.frame_offset[u] + t) * p.num_ceps;
  for (int c = lane; c < p.num_ceps; c += 32) {
    const float *col = p.dct_t + c;  // transposed: the lanes (one cepstrum each) read consecutive addresses
    float acc = 0.f;
    for (int b = 0; b < p.num_bins; b++) acc = fmaf(col[(size_t)b * p.num_ceps], melv[b], acc);
    if (p.lifter) acc = fmul(acc, p.lifter[c]);
    if (p.use_energy && c == 0) acc = log_energy;
    out[c] = acc;
  }
}

void LaunchMfcc(const FeatParams &p, int n_utts, int max_frames, cudaStream_t stream) {
  if (n_utts == 0 || max_frames == 0) return;
  dim3 grid((max_frames + kWarpsPerCta - 1) / kWarpsPerCta, n_utts);
  size_t smem = (size_t)kWarpsPerCta * (p.padded + p.num_bins + 8) * sizeof(float);
  mfcc_kernel<<<grid, kWarpsPerCta * 32, smem, stream>>>(p);
}

}  // namespace rs
