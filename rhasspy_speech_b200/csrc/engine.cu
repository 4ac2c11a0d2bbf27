// Host engine + C ABI (include/rs_b200.h): uploads the model tables, lays a batch of utterances out
// on the global time axis, launches the three stages on one CUDA stream and returns word ids.
// There is no CPU fallback: every entry point fails loudly if CUDA is unavailable.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <map>
#include <memory>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "../../include/rs_b200.h"
#include "engine.h"
#include "kaldi_io.h"
#include "model.h"
#include "nnet_tc.h"
#include "nbest.h"
#include "strict_decode.h"

namespace rs {

#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) RS_FAIL("CUDA error: " << cudaGetErrorString(e_) << " at " << #expr << " (engine.cu:" << __LINE__ << ")"); \
  } while (0)

struct DevBuf {  // grow-only device allocation
  void *p = nullptr;
  size_t cap = 0;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  void *ensure(size_t bytes) {
    if (bytes > cap) {
      if (p) CUDA_OK(cudaFree(p));
      p = nullptr;
      size_t want = bytes + bytes / 8 + 256;
      CUDA_OK(cudaMalloc(&p, want));
      // zero on growth: padding columns of activation / feature buffers are multiplied by zero weights and must
      // be finite; fresh driver memory is zero, memory recycled from an earlier decoder of the process is not
      CUDA_OK(cudaMemset(p, 0, want));
      CUDA_OK(cudaDeviceSynchronize());
      cap = want;
    }
    return p;
  }
  template <typename T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};
// page-locked host ranges handed out by rs_host_alloc
static std::mutex g_pinned_mu;
static std::map<uintptr_t, size_t> g_pinned;
static bool HostRangeIsPinned(const void *p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  auto it = g_pinned.upper_bound(a);
  if (it == g_pinned.begin()) return false;
  --it;
  return a >= it->first && a + bytes <= it->first + it->second;
}

struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
  void *ensure(size_t bytes) {
    if (bytes > cap) {
      if (p) CUDA_OK(cudaFreeHost(p));
      p = nullptr;
      size_t want = bytes + bytes / 8 + 256;
      CUDA_OK(cudaMallocHost(&p, want));
      cap = want;
    }
    return p;
  }
};

template <typename T>
static T *Upload(const std::vector<T> &v, std::vector<void *> *owned) {
  void *d = nullptr;
  size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  CUDA_OK(cudaMalloc(&d, bytes));
  if (!v.empty()) CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  owned->push_back(d);
  return reinterpret_cast<T *>(d);
}

// --------------------------------------------------------------------------------------------
// feature tables, built with the reference's float expressions (host libm, no FMA contraction):
// feature-window.cc:109-135, srfft.cc:77-118 / 180-204 / 356-375, mel-computations.cc:33-142,
// matrix-functions.cc:592-608, mel-computations.cc:253-259.
struct FeatTables {
  std::vector<float> window, twiddle, kn, mel_weights, dct, lifter;
  std::vector<uint16_t> level_offsets, perm;
  int level_start[12] = {0}, level_count[12] = {0}, twiddle_start[12] = {0};
  std::vector<int> mel_offset, mel_len, mel_start;
  int logn = 0;
};

static void CollectBlocks(int offset, int logn, std::vector<std::vector<uint16_t>> *levels) {
  if (logn < 1) return;
  (*levels)[logn].push_back((uint16_t)offset);
  if (logn < 3) return;
  int m = 1 << logn;
  CollectBlocks(offset, logn - 1, levels);
  CollectBlocks(offset + m / 2, logn - 2, levels);
  CollectBlocks(offset + 3 * (m / 4), logn - 2, levels);
}

static void BuildFeatTables(const MfccOptions &o, FeatTables *t) {
  const int L = o.WindowSize(), N = o.PaddedWindowSize(), NH = N / 2;
  if (N > 1024 || N < 16) RS_FAIL("unsupported FFT size " << N);
  if (o.window_type != "povey" && o.window_type != "hamming" && o.window_type != "hanning" && o.window_type != "rectangular")
    RS_FAIL("unsupported window type " << o.window_type);
  t->window.resize(L);
  const double a = 6.283185307179586476925286766559005 / (L - 1);
  for (int i = 0; i < L; i++) {
    double x = (double)i;
    double w = 1.0;
    if (o.window_type == "povey") w = pow(0.5 - 0.5 * cos(a * x), 0.85);
    else if (o.window_type == "hamming") w = 0.54 - 0.46 * cos(a * x);
    else if (o.window_type == "hanning") w = 0.5 - 0.5 * cos(a * x);
    t->window[i] = (float)w;
  }
  int logn = 0;
  while ((1 << logn) < NH) logn++;
  t->logn = logn;
  std::vector<std::vector<uint16_t>> levels(12);
  CollectBlocks(0, logn, &levels);
  for (int lv = 0; lv < 12; lv++) {
    t->level_start[lv] = (int)t->level_offsets.size();
    t->level_count[lv] = (int)levels[lv].size();
    t->level_offsets.insert(t->level_offsets.end(), levels[lv].begin(), levels[lv].end());
  }
  const double M_2PI_ = 6.283185307179586476925286766559005;
  for (int lv = 4; lv <= logn; lv++) {
    int m = 1 << lv, m4 = m / 4, m8 = m / 8, nel = m4 - 2;
    t->twiddle_start[lv] = (int)t->twiddle.size();
    std::vector<float> tab(6 * nel);
    int k = 0;
    for (int n = 1; n < m4; n++) {
      if (n == m8) continue;
      float ang = n * M_2PI_ / m;
      float c = std::cos(ang), s = std::sin(ang);
      tab[k] = c;
      tab[nel + k] = -(s + c);
      tab[2 * nel + k] = s - c;
      ang = 3 * n * M_2PI_ / m;
      c = std::cos(ang);
      s = std::sin(ang);
      tab[3 * nel + k] = c;
      tab[4 * nel + k] = -(s + c);
      tab[5 * nel + k] = s - c;
      k++;
    }
    t->twiddle.insert(t->twiddle.end(), tab.begin(), tab.end());
  }
  {  // bit-reversal: replay the reference's swap sequence on an index vector
    int lg2 = logn >> 1;
    if (logn & 1) lg2++;
    std::vector<int> brseed(1 << lg2, 0);
    brseed[0] = 0;
    if (brseed.size() > 1) brseed[1] = 1;
    for (int j = 2; j <= lg2; j++) {
      int imax = 1 << (j - 1);
      for (int i = 0; i < imax; i++) {
        brseed[i] <<= 1;
        brseed[i + imax] = brseed[i] + 1;
      }
    }
    std::vector<int> perm(NH);
    for (int i = 0; i < NH; i++) perm[i] = i;
    lg2 = logn >> 1;
    int n = 1 << lg2;
    for (int off = 1; off < n; off++) {
      int fj = n * brseed[off], i = off, j = fj;
      std::swap(perm[i], perm[j]);
      int xp = i;
      for (int gno = 1; gno < brseed[off]; gno++) {
        xp += n;
        j = fj + brseed[gno];
        std::swap(perm[xp], perm[j]);
      }
    }
    t->perm.assign(perm.begin(), perm.end());
  }
  {  // real-FFT twiddles: kN *= rootN in float (ComplexMul)
    float x = (float)(M_2PI_ / N * -1);
    float rre = std::cos(x), rim = std::sin(x);
    float kre = 1.0f, kim = 0.0f;
    for (int k = 1; 2 * k <= NH; k++) {
      float tre = (kre * rre) - (kim * rim);
      kim = kre * rim + kim * rre;
      kre = tre;
      t->kn.push_back(kre);
      t->kn.push_back(kim);
    }
  }
  {  // mel banks
    float sample_freq = o.samp_freq, nyquist = 0.5f * sample_freq;
    float low_freq = o.low_freq, high_freq = o.high_freq > 0.0f ? o.high_freq : nyquist + o.high_freq;
    if (low_freq < 0.0 || low_freq >= nyquist || high_freq <= 0.0 || high_freq > nyquist || high_freq <= low_freq)
      RS_FAIL("bad mel options: low-freq " << low_freq << " high-freq " << high_freq);
    float fft_bin_width = sample_freq / N;
    auto mel = [](float f) -> float { return 1127.0f * logf(1.0f + f / 700.0f); };
    float mel_low = mel(low_freq), mel_high = mel(high_freq);
    float delta = (mel_high - mel_low) / (o.num_bins + 1);
    for (int b = 0; b < o.num_bins; b++) {
      float left = mel_low + b * delta, center = mel_low + (b + 1) * delta, right = mel_low + (b + 2) * delta;
      std::vector<float> w(NH, 0.f);
      int first = -1, last = -1;
      for (int i = 0; i < NH; i++) {
        float freq = fft_bin_width * i;
        float m = mel(freq);
        if (m > left && m < right) {
          float weight;
          if (m <= center) weight = (m - left) / (center - left);
          else weight = (right - m) / (right - center);
          w[i] = weight;
          if (first == -1) first = i;
          last = i;
        }
      }
      if (first < 0) RS_FAIL("--num-mel-bins is too large for this window");
      t->mel_offset.push_back(first);
      t->mel_len.push_back(last + 1 - first);
      t->mel_start.push_back((int)t->mel_weights.size());
      t->mel_weights.insert(t->mel_weights.end(), w.begin() + first, w.begin() + last + 1);
    }
  }
  {
    int K = o.num_ceps, Nb = o.num_bins;
    t->dct.assign((size_t)K * Nb, 0.f);
    float normalizer = std::sqrt(1.0 / static_cast<float>(Nb));
    for (int j = 0; j < Nb; j++) t->dct[j] = normalizer;
    normalizer = std::sqrt(2.0 / static_cast<float>(Nb));
    for (int k = 1; k < K; k++)
      for (int n = 0; n < Nb; n++)
        t->dct[(size_t)k * Nb + n] = normalizer * std::cos(static_cast<double>(M_PI) / Nb * (n + 0.5) * k);
    if (o.cepstral_lifter != 0.0f) {
      float Q = o.cepstral_lifter;
      for (int i = 0; i < K; i++) t->lifter.push_back(1.0 + 0.5 * Q * sin(M_PI * i / Q));
    }
  }
}

// --------------------------------------------------------------------------------------------
// A few persistent host threads that pack the caller's PCM buffers into pinned memory item by item
// (thread creation per call cost more than the copies themselves).
class PackPool {
 public:
  explicit PackPool(int n_threads) {
    for (int i = 0; i < n_threads; i++) threads_.emplace_back([this] { Loop(); });
  }
  ~PackPool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
  }
  void Start(int n_items, std::function<void(int)> fn) {
    WaitAll();
    {
      std::lock_guard<std::mutex> l(mu_);
      job_ = std::move(fn);
      n_items_ = n_items;
      next_.store(0);
      done_.reset(new std::atomic<int>[std::max(n_items, 1)]);
      for (int i = 0; i < n_items; i++) done_[i].store(0);
      active_ = (int)threads_.size();
      gen_++;
    }
    cv_.notify_all();
  }
  bool Help() {  // run one item on the calling thread; false when none is left
    const int i = next_.fetch_add(1);
    if (i >= n_items_) return false;
    job_(i);
    done_[i].store(1, std::memory_order_release);
    return true;
  }
  void WaitItem(int i) {
    while (!done_[i].load(std::memory_order_acquire))
      if (!Help()) std::this_thread::yield();
  }
  void WaitAll() {  // every worker has left the current job (its captures may go out of scope)
    std::unique_lock<std::mutex> l(mu_);
    idle_cv_.wait(l, [this] { return active_ == 0; });
  }

 private:
  void Loop() {
    uint64_t seen = 0;
    while (true) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      while (Help()) {
      }
      {
        std::lock_guard<std::mutex> l(mu_);
        active_--;
      }
      idle_cv_.notify_all();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, idle_cv_;
  std::function<void(int)> job_;
  std::unique_ptr<std::atomic<int>[]> done_;
  std::atomic<int> next_{0};
  int n_items_ = 0, active_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// --------------------------------------------------------------------------------------------
struct ModelImpl {
  Model m;
  int device = 0;
  std::vector<void *> owned;
  std::string plan_text;
  // features
  FeatTables ft;
  FeatParams feat{};  // table pointers filled, batch pointers per call
  // ivector
  IvecParams ivec{};
  const double *d_global_cmvn = nullptr, *d_nnet_global_cmvn = nullptr;
  // nnet
  std::vector<const float *> d_matrices, d_vectors;
  // tensor-core path (nnet_tc.cu): per plan step, the packed + split weights and their tensor maps
  struct TcStep {
    bool ok = false;
    const __half *w_hi = nullptr, *w_lo = nullptr;
    int kp = 0, bn = 0;
    std::vector<int> k0;
    CUtensorMap map_hi, map_lo;
  };
  bool use_tc = true;
  std::vector<TcStep> tc_steps;
  std::vector<char> split;  // per plan buffer: stored as two TF32 planes
  int num_sms = 0;
  uint64_t flops_per_axis_unit = 0;  // sum over gemm steps of 2*n*ktot/step (per time unit of the axis)
  uint64_t bytes_per_axis_unit = 0;  // algorithmic HBM bytes of the same steps: each input once, bypass, output (x1000)
  ~ModelImpl() {
    for (void *p : owned) cudaFree(p);
  }
};

struct GraphImpl {
  Graph g;
  int device = 0;
  std::vector<void *> owned;
  DevGraph dev{};
  ~GraphImpl() {
    for (void *p : owned) cudaFree(p);
  }
};

struct StreamImpl;

constexpr int kMaxStagingItems = 64;

struct DecoderImpl {
  ModelImpl *model = nullptr;
  GraphImpl *graph = nullptr;
  rs_decoder_opts opts{};
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // audio H2D copies of a staged batch (the feature kernels of item k run under the copy of item k + 1)
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t item_ev[kMaxStagingItems] = {};
  bool staging_overlap = true;  // rs_decoder_set_staging_overlap
  int *d_item_flag = nullptr;   // written behind the first item's copy; staging_spin_kernel watches it
  int *h_item_seq = nullptr;    // pinned source word of that flag copy
  int staging_seq = 0;
  double host_last_sync_ms = 0.0;  // RS_B200_HOST_PROFILE
  int n_lanes = 0;
  std::vector<void *> owned;
  DevBuf d_pcm, d_desc, d_mfcc, d_mfcc_norm, d_xraw, d_xnorm, d_post_idx, d_post_w, d_wf, d_gw, d_linear, d_quad;
  DevBuf d_out;  // decode outputs
  DevBuf d_small_arena, d_small_off;  // decode_small.cu: batch traceback arena and its per-utterance offsets
  PinBuf h_small;
  // n-best tail (rs_decoder_set_nbest): lattice recorded by decode_kernel<true>, pruned + compacted on the device
  int nbest = 1;
  float nbest_scale = 1.0f;
  size_t lattice_mb = 0;  // current lattice budget (grows when a batch did not fit)
  DevBuf d_lat_tok, d_lat_extra, d_lat_newid, d_lat_link, d_lat_surv, d_lat_tb, d_lat_pos, d_lat_off, d_lat_hdr, d_lat_arcs;
  PinBuf h_lat;
  std::vector<LatticeHeader> lat_hdr;  // of the last n-best call (rs_debug_fetch item 5)
  std::map<int, std::vector<LatticeArc>> strict_lat;  // lattices of the utterances the strict-order decoder re-decoded
  std::vector<DevBuf> slots;
  DevBuf d_tid_pdf;
  DevBuf d_loglikes_ext;
  PinBuf h_in, h_out;
  PinBuf h_ll;  // log-likelihoods of the utterances the strict-order host decoder takes over
  LaneWorkspace *d_lanes = nullptr;   // general decode kernel only; see EnsureLaneWorkspace
  int lanes_allocated = 0;
  std::vector<void *> lane_owned;
  int *d_next_utt = nullptr;
  int *d_range_flag = nullptr;  // set by a split store that had to saturate (fp16 planes)
  int *h_range_flag = nullptr;  // pinned
  std::unique_ptr<PackPool> pool;
  DevBuf d_earc_buf;                // graph arcs with ilabel mapped to pdf for this model
  const int4 *d_earc = nullptr;
  std::vector<int32_t> h_epdf;      // the same map on the host, for the strict-order decoder (strict_decode.cc)
  StrictArcs strict_arcs;           // its arc records, built when the first utterance needs them
  bool strict_arcs_ready = false;
  rs_timings last{};
  // layout of the last batch (for rs_debug_fetch)
  struct Batch {
    int n = 0, total_frames = 0, axis_len = 0;
    std::vector<int> num_frames, frame_offset, origin, n_out, ll_row0;
    std::vector<int> v_begin;  // first iVector solve (row of the per-utterance buffers) of each utterance
    bool from_loglikes = false;
    const float *loglikes = nullptr;
    int ll_ld = 0;
  } batch;
  ~DecoderImpl() {
    for (void *p : owned) cudaFree(p);
    for (void *p : lane_owned) cudaFree(p);
    for (auto &e : ev)
      if (e) cudaEventDestroy(e);
    for (auto &e : item_ev)
      if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (stream) cudaStreamDestroy(stream);
    if (h_range_flag) cudaFreeHost(h_range_flag);
    if (h_item_seq) cudaFreeHost(h_item_seq);
    if (d_item_flag) cudaFree(d_item_flag);
  }
};

struct StreamImpl {
  DecoderImpl *dec;
  std::vector<int16_t> pcm;
};

// One warp that keeps the decoder's stream busy until the first staged item has landed (or max_clk has passed).
// Measured: a kernel that arrives on an idle stream while the copy stream still has copies queued is not started before
// the last of them is done; with this warp in front, the stream is running when the first MFCC kernel arrives and the
// kernels do run under the copies.  Nothing depends on what it sees (the host launches every MFCC kernel only after its
// item's copy event), it holds one warp slot, and it leaves after max_clk whatever happens.
__global__ void staging_spin_kernel(const int *flag, int value, long long max_clk) {
  const long long t0 = clock64();
  int v;
  do {
    asm volatile("ld.relaxed.sys.global.b32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v == value) break;
    __nanosleep(128);
  } while (clock64() - t0 < max_clk);
}

static int RoundUp(int x, int m) { return (x + m - 1) / m * m; }

static void UploadModel(ModelImpl *mi) {
  Model &m = mi->m;
  auto &own = mi->owned;
  // --- features
  BuildFeatTables(m.mfcc, &mi->ft);
  FeatTables &t = mi->ft;
  FeatParams &f = mi->feat;
  f.shift = m.mfcc.WindowShift();
  f.length = m.mfcc.WindowSize();
  f.padded = m.mfcc.PaddedWindowSize();
  if (f.padded > 512) RS_FAIL("frame lengths beyond 512 samples (padded) are not supported by the MFCC kernel: " << f.padded);
  f.logn = t.logn;
  f.preemph = m.mfcc.preemph_coeff;
  f.dither = m.mfcc.dither;
  f.energy_floor = m.mfcc.energy_floor;
  f.remove_dc = m.mfcc.remove_dc_offset;
  f.use_energy = m.mfcc.use_energy;
  f.raw_energy = m.mfcc.raw_energy;
  f.num_bins = m.mfcc.num_bins;
  f.num_ceps = m.mfcc.num_ceps;
  f.window = Upload(t.window, &own);
  f.level_offsets = Upload(t.level_offsets, &own);
  memcpy(f.level_start, t.level_start, sizeof(f.level_start));
  memcpy(f.level_count, t.level_count, sizeof(f.level_count));
  f.twiddle = Upload(t.twiddle, &own);
  memcpy(f.twiddle_start, t.twiddle_start, sizeof(f.twiddle_start));
  f.perm = Upload(t.perm, &own);
  f.kn = Upload(t.kn, &own);
  f.mel_offset = Upload(t.mel_offset, &own);
  f.mel_len = Upload(t.mel_len, &own);
  f.mel_start = Upload(t.mel_start, &own);
  f.mel_weights = Upload(t.mel_weights, &own);
  {
    // lane-friendly copies: tap-major mel weights [max_len][num_bins] and the DCT matrix transposed [num_bins][num_ceps],
    // so that the 32 lanes of a warp (one mel bin / one cepstrum each) read consecutive addresses
    const int Nb = (int)t.mel_len.size(), K = (int)(t.dct.size() / (size_t)Nb);
    int max_len = 0;
    for (int b = 0; b < Nb; b++) max_len = std::max(max_len, t.mel_len[b]);
    std::vector<float> mel_t((size_t)max_len * Nb, 0.f), dct_t((size_t)Nb * K);
    for (int b = 0; b < Nb; b++)
      for (int i = 0; i < t.mel_len[b]; i++) mel_t[(size_t)i * Nb + b] = t.mel_weights[t.mel_start[b] + i];
    for (int k = 0; k < K; k++)
      for (int b = 0; b < Nb; b++) dct_t[(size_t)b * K + k] = t.dct[(size_t)k * Nb + b];
    f.mel_weights_t = Upload(mel_t, &own);
    f.mel_max_len = max_len;
    f.dct_t = Upload(dct_t, &own);
  }
  f.dct = Upload(t.dct, &own);
  f.lifter = t.lifter.empty() ? nullptr : Upload(t.lifter, &own);
  // --- ivector
  if (m.has_ivector) {
    IvecParams &iv = mi->ivec;
    const int D = m.mfcc.num_ceps, ns = m.ivec.splice_left + 1 + m.ivec.splice_right, K = D * ns;
    iv.dim = D;
    iv.left = m.ivec.splice_left;
    iv.right = m.ivec.splice_right;
    iv.ldim = m.lda.rows;
    std::vector<float> lda_t((size_t)K * iv.ldim), lda_bias;
    for (int j = 0; j < iv.ldim; j++)
      for (int k = 0; k < K; k++) lda_t[(size_t)k * iv.ldim + j] = m.lda(j, k);
    if (m.lda.cols == K + 1)
      for (int j = 0; j < iv.ldim; j++) lda_bias.push_back(m.lda(j, K));
    iv.lda_t = Upload(lda_t, &own);
    iv.lda_bias = lda_bias.empty() ? nullptr : Upload(lda_bias, &own);
    const int G = m.ubm.num_gauss;
    if (G > 32 * 64) RS_FAIL("UBMs with more than 2048 Gaussians are not supported");
    if (m.ivec.num_gselect > 8) RS_FAIL("--num-gselect > 8 is not supported");
    iv.num_gauss = G;
    iv.gconsts = Upload(m.ubm.gconsts, &own);
    std::vector<float> mt((size_t)iv.ldim * G), vt((size_t)iv.ldim * G);
    for (int g = 0; g < G; g++)
      for (int d = 0; d < iv.ldim; d++) {
        mt[(size_t)d * G + g] = m.ubm.means_invvars(g, d);
        vt[(size_t)d * G + g] = m.ubm.inv_vars(g, d);
      }
    iv.means_invvars_t = Upload(mt, &own);
    iv.inv_vars_t = Upload(vt, &own);
    iv.num_gselect = m.ivec.num_gselect;
    iv.min_post = m.ivec.min_post;
    iv.posterior_scale = m.ivec.posterior_scale;
    iv.ivector_dim = m.ie.ivector_dim;
    iv.sigma_inv_m = Upload(m.ie.sigma_inv_m, &own);
    iv.u = Upload(m.ie.u, &own);
    iv.prior_offset = m.ie.prior_offset;
    iv.max_count = m.ivec.max_count;
    iv.num_cg_iters = m.ivec.num_cg_iters;
    iv.online_cmvn_iextractor = m.ivec.online_cmvn_iextractor;
    mi->d_global_cmvn = Upload(m.global_cmvn.d, &own);
  }
  if (m.nnet_cmvn) mi->d_nnet_global_cmvn = Upload(m.nnet_global_cmvn.d, &own);
  // --- nnet: fold "subtract log-prior, apply acoustic scale" (decodable-online-looped.cc:218-223) into
  // the epilogue of the step that writes the output buffer
  Plan &pl = m.plan;
  for (int si = (int)pl.steps.size() - 1; si >= 0; si--)
    if (pl.steps[si].out == pl.output_buffer) {
      if (!m.log_priors.empty()) {
        std::vector<float> neg(m.log_priors.size());
        for (size_t i = 0; i < neg.size(); i++) neg[i] = -m.log_priors[i];
        pl.vectors.push_back(neg);
        EpiOp op;
        op.type = EpiOp::kBias;
        op.vec0 = (int)pl.vectors.size() - 1;
        pl.steps[si].ops.push_back(op);
      }
      if (m.acoustic_scale != 1.0f) {
        EpiOp op;
        op.type = EpiOp::kScale;
        op.alpha = m.acoustic_scale;
        pl.steps[si].ops.push_back(op);
      }
      break;
    }
  for (const auto &mat : pl.matrices) mi->d_matrices.push_back(Upload(mat.d, &own));
  for (const auto &v : pl.vectors) {
    std::vector<float> padded(v);  // the epilogues read bias / scale vectors as float4
    padded.resize((v.size() + 3) / 4 * 4, 0.f);
    mi->d_vectors.push_back(Upload(padded, &own));
  }
  // --- tensor-core layers: every time-axis GEMM whose slabs are integer-strided row views
  {
    const char *env = getenv("RS_B200_GEMM");
    mi->use_tc = !(env && std::string(env) == "simt");
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, mi->device));
    mi->num_sms = prop.multiProcessorCount;
    if (mi->use_tc && prop.major != 10)
      RS_FAIL("this library is built for sm_100a (B200); device " << mi->device << " is sm_" << prop.major << prop.minor);
    mi->tc_steps.resize(pl.steps.size());
    mi->split.assign(pl.buffers.size(), 0);
    for (size_t si = 0; mi->use_tc && si < pl.steps.size(); si++) {
      const Step &st = pl.steps[si];
      if (st.type != Step::kGemm || (int)st.slabs.size() > kTcMaxSlabs) continue;
      const PlanBuffer &ob = pl.buffers[st.out];
      bool ok = !ob.per_utt;
      for (const Slab &sl : st.slabs) {
        const PlanBuffer &sb = pl.buffers[sl.src];
        if (sb.per_utt || ob.step % sb.step != 0 || sl.t_offset % sb.step != 0 || sl.k > sb.dim || sl.src == pl.output_buffer)
          ok = false;
      }
      if (!ok) continue;
      ModelImpl::TcStep &ts = mi->tc_steps[si];
      std::vector<std::pair<int, int>> cols;
      for (const Slab &sl : st.slabs) cols.push_back({sl.wcol, sl.k});
      std::vector<__half> hi, lo;
      TcPackWeights(pl.matrices[st.weight].d.data(), st.n, st.ktot, cols, &hi, &lo, &ts.k0, &ts.kp);
      ts.w_hi = Upload(hi, &own);
      ts.w_lo = Upload(lo, &own);
      ts.bn = TcTileN(st.n);
      TcEncodeMap(&ts.map_hi, ts.w_hi, st.n, ts.kp, ts.kp, ts.bn);
      TcEncodeMap(&ts.map_lo, ts.w_lo, st.n, ts.kp, ts.kp, ts.bn);
      ts.ok = true;
      for (const Slab &sl : st.slabs) mi->split[sl.src] = 1;
    }
  }
  for (const Step &st : pl.steps) {
    if ((int)st.slabs.size() > kMaxSlabs) RS_FAIL("layer " << st.name << " has too many input blocks");
    if ((int)st.ops.size() > kMaxOps) RS_FAIL("layer " << st.name << " has too many fused operations");
    if (st.type == Step::kGemm) {
      uint64_t k = 0;
      for (const auto &s : st.slabs) k += s.k;
      mi->flops_per_axis_unit += 2ull * st.n * k * 1000 / pl.buffers[st.out].step;  // x1000 fixed point
      // 4 bytes per element: a split buffer is two fp16 planes, a plain one fp32
      uint64_t b = 4ull * st.n * 1000 / pl.buffers[st.out].step;
      std::vector<int> seen;
      for (const auto &sl : st.slabs)
        if (!pl.buffers[sl.src].per_utt && std::find(seen.begin(), seen.end(), sl.src) == seen.end()) {
          seen.push_back(sl.src);
          b += 4ull * pl.buffers[sl.src].dim * 1000 / pl.buffers[sl.src].step;
        }
      for (const auto &op : st.ops)
        if (op.type == EpiOp::kAddScaled && op.buffer >= 0) b += 4ull * pl.buffers[op.buffer].dim * 1000 / pl.buffers[op.buffer].step;
      mi->bytes_per_axis_unit += b;
    }
  }
  mi->plan_text = DescribePlan(pl);
}

}  // namespace rs

using namespace rs;

static void SetErr(char *err, size_t errlen, const std::string &msg) {
  if (err && errlen) {
    size_t n = std::min(errlen - 1, msg.size());
    memcpy(err, msg.data(), n);
    err[n] = 0;
  }
}

#define API_GUARD_BEGIN try {
#define API_GUARD_END(ret)                          \
  }                                                 \
  catch (const std::exception &e) {                 \
    SetErr(err, errlen, e.what());                  \
    return ret;                                     \
  }

extern "C" {

int rs_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

void rs_decoder_opts_default(rs_decoder_opts *o) {
  o->beam = 24.0f;
  o->max_active = 7000;
  o->min_active = 200;
  o->lattice_beam = 8.0f;
  o->acoustic_scale = 1.0f;
  o->beam_delta = 0.5f;
  o->max_tokens_per_frame = 65536;
  o->max_tokens_per_utt = 4194304;
  o->max_words = 256;
  o->num_lanes = 0;
  o->dither_seed = 0;
  o->strict_fallback = 1;
}

rs_model *rs_model_load(const char *final_mdl, const char *online_conf, int device, char *err, size_t errlen) {
  API_GUARD_BEGIN
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    RS_FAIL("no CUDA device available: this library has no CPU path");
  if (device < 0 || device >= ndev) RS_FAIL("CUDA device " << device << " does not exist (" << ndev << " visible)");
  CUDA_OK(cudaSetDevice(device));
  std::unique_ptr<ModelImpl> mi(new ModelImpl());
  mi->device = device;
  LoadModel(final_mdl, online_conf, &mi->m);
  UploadModel(mi.get());
  return reinterpret_cast<rs_model *>(mi.release());
  API_GUARD_END(nullptr)
}

void rs_model_free(rs_model *m) { delete reinterpret_cast<ModelImpl *>(m); }

int rs_model_info(const rs_model *m_, int32_t *num_pdfs, int32_t *sf, int32_t *ivector_dim, int32_t *feat_dim,
                  int32_t *left_context, int32_t *right_context) {
  const ModelImpl *mi = reinterpret_cast<const ModelImpl *>(m_);
  if (!mi) return 1;
  if (num_pdfs) *num_pdfs = mi->m.trans.num_pdfs;
  if (sf) *sf = mi->m.frame_subsampling_factor;
  if (ivector_dim) *ivector_dim = mi->m.has_ivector ? mi->m.ie.ivector_dim : 0;
  if (feat_dim) *feat_dim = mi->m.mfcc.num_ceps;
  if (left_context) *left_context = mi->m.plan.left_context;
  if (right_context) *right_context = mi->m.plan.right_context;
  return 0;
}

const char *rs_model_plan(const rs_model *m_) {
  const ModelImpl *mi = reinterpret_cast<const ModelImpl *>(m_);
  return mi ? mi->plan_text.c_str() : "";
}

rs_graph *rs_graph_load(const char *hclg_fst, const char *words_txt, int device, char *err, size_t errlen) {
  API_GUARD_BEGIN
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    RS_FAIL("no CUDA device available: this library has no CPU path");
  if (device < 0 || device >= ndev) RS_FAIL("CUDA device " << device << " does not exist");
  CUDA_OK(cudaSetDevice(device));
  std::unique_ptr<GraphImpl> gi(new GraphImpl());
  gi->device = device;
  LoadGraph(hclg_fst, words_txt ? words_txt : "", &gi->g);
  Graph &g = gi->g;
  DevGraph &d = gi->dev;
  d.num_states = g.num_states;
  d.start = (int)g.start;
  d.num_earcs = (unsigned)g.e_next.size();
  d.num_parcs = (unsigned)g.p_next.size();
  d.e_begin = Upload(g.e_begin, &gi->owned);
  d.p_begin = Upload(g.p_begin, &gi->owned);
  d.e_src = Upload(g.e_src, &gi->owned);
  d.p_src = Upload(g.p_src, &gi->owned);
  d.final_cost = Upload(g.final_cost, &gi->owned);
  std::vector<int4> parc(g.p_next.size());
  for (size_t i = 0; i < parc.size(); i++) {
    int wbits;
    memcpy(&wbits, &g.p_weight[i], 4);
    parc[i] = make_int4(g.p_next[i], 0, wbits, g.p_olabel[i]);
  }
  d.parc = Upload(parc, &gi->owned);
  d.eps_flat = 1;
  for (size_t i = 0; i < g.p_next.size(); i++)
    if (g.p_begin[g.p_next[i] + 1] > g.p_begin[g.p_next[i]]) d.eps_flat = 0;
  d.earc = nullptr;  // per decoder: ilabels are mapped through the model's tid -> pdf table
  return reinterpret_cast<rs_graph *>(gi.release());
  API_GUARD_END(nullptr)
}

void rs_graph_free(rs_graph *g) { delete reinterpret_cast<GraphImpl *>(g); }

int rs_graph_info(const rs_graph *g_, int32_t *num_states, int64_t *num_arcs, int32_t *num_words) {
  const GraphImpl *gi = reinterpret_cast<const GraphImpl *>(g_);
  if (!gi) return 1;
  if (num_states) *num_states = gi->g.num_states;
  if (num_arcs) *num_arcs = (int64_t)gi->g.e_next.size() + (int64_t)gi->g.p_next.size();
  if (num_words) *num_words = (int32_t)gi->g.words.size();
  return 0;
}

const char *rs_graph_word(const rs_graph *g_, int32_t id) {
  const GraphImpl *gi = reinterpret_cast<const GraphImpl *>(g_);
  if (!gi || id < 0 || (size_t)id >= gi->g.words.size() || gi->g.words[id].empty()) return nullptr;
  return gi->g.words[id].c_str();
}

// arcs with ilabel -> pdf (TransitionIdToPdfFast, decodable-online-looped.cc:249-256) for (model, graph)
static void BindGraph(DecoderImpl *d, GraphImpl *gi) {
  const Graph &g = gi->g;
  const auto &t2p = d->model->m.trans.tid2pdf;
  std::vector<int4> earc(g.e_next.size());
  std::vector<int32_t> epdf(g.e_next.size());
  for (size_t i = 0; i < earc.size(); i++) {
    int il = g.e_ilabel[i];
    if (il <= 0 || il >= (int)t2p.size())
      RS_FAIL("HCLG has input label " << il << " but the model has only " << t2p.size() - 1 << " transition-ids");
    int wbits;
    memcpy(&wbits, &g.e_weight[i], 4);
    earc[i] = make_int4(g.e_next[i], t2p[il], wbits, g.e_olabel[i]);
    epdf[i] = t2p[il];
  }
  // a fresh allocation, so that a failed bind leaves the decoder on its previous graph
  DevBuf fresh;
  fresh.ensure(std::max<size_t>(earc.size() * sizeof(int4), 16));
  if (!earc.empty()) CUDA_OK(cudaMemcpy(fresh.p, earc.data(), earc.size() * sizeof(int4), cudaMemcpyHostToDevice));
  std::swap(d->d_earc_buf.p, fresh.p);
  std::swap(d->d_earc_buf.cap, fresh.cap);
  d->d_earc = d->d_earc_buf.as<int4>();
  d->h_epdf.swap(epdf);
  d->strict_arcs_ready = false;
  d->graph = gi;
}


// Workspace of the general decode kernel (decode.cu): per-lane state tables and a traceback arena.  Allocated on the
// first batch that needs it -- small graphs (decode_small.cu) never do -- and for the lanes that batch can occupy; a later,
// larger batch grows it.  (The arena is max_tokens_per_utt records per lane: 32 MB at the default.)
static void EnsureLaneWorkspace(DecoderImpl *d, int lanes_needed) {
  if (d->d_lanes && d->lanes_allocated >= lanes_needed) return;
  const rs_decoder_opts &o = d->opts;
  CUDA_OK(cudaStreamSynchronize(d->stream));
  for (void *p : d->lane_owned) cudaFree(p);
  d->lane_owned.clear();
  d->d_lanes = nullptr;
  int hash_size = 1;
  while (hash_size < 2 * o.max_tokens_per_frame) hash_size <<= 1;
  const size_t H = hash_size, C = o.max_tokens_per_frame;
  const int L = std::min(d->n_lanes, std::max(lanes_needed, 8));
  std::vector<LaneWorkspace> lanes(L);
  auto dalloc = [&](size_t bytes, int fill) {
    void *p = nullptr;
    CUDA_OK(cudaMalloc(&p, bytes));
    CUDA_OK(cudaMemset(p, fill, bytes));
    d->lane_owned.push_back(p);
    return p;
  };
  // one allocation per array kind, sliced per lane
  int *hkey = (int *)dalloc(sizeof(int) * H * 2 * L, 0xff);
  unsigned long long *hval = (unsigned long long *)dalloc(sizeof(unsigned long long) * H * 2 * L, 0xff);
  int *hidx = (int *)dalloc(sizeof(int) * H * 2 * L, 0xff);
  int *inq = (int *)dalloc(sizeof(int) * H * L, 0);
  int *ins = (int *)dalloc(sizeof(int) * C * 2 * L, 0);
  int *tstate = (int *)dalloc(sizeof(int) * C * 2 * L, 0);
  float *tcost = (float *)dalloc(sizeof(float) * C * 2 * L, 0);
  int *tslot = (int *)dalloc(sizeof(int) * C * L, 0);
  unsigned *pfx = (unsigned *)dalloc(sizeof(unsigned) * (C + 1) * L, 0);
  int *frontier = (int *)dalloc(sizeof(int) * C * 2 * L, 0);
  int2 *arena = (int2 *)dalloc(sizeof(int2) * (size_t)o.max_tokens_per_utt * L, 0);
  for (int l = 0; l < L; l++) {
    LaneWorkspace &w = lanes[l];
    for (int k = 0; k < 2; k++) {
      w.hkey[k] = hkey + ((size_t)l * 2 + k) * H;
      w.hval[k] = hval + ((size_t)l * 2 + k) * H;
      w.hidx[k] = hidx + ((size_t)l * 2 + k) * H;
      w.ins_list[k] = ins + ((size_t)l * 2 + k) * C;
      w.tok_state[k] = tstate + ((size_t)l * 2 + k) * C;
      w.tok_cost[k] = tcost + ((size_t)l * 2 + k) * C;
      w.frontier[k] = frontier + ((size_t)l * 2 + k) * C;
    }
    w.inq = inq + (size_t)l * H;
    w.tok_slot = tslot + (size_t)l * C;
    w.pfx = pfx + (size_t)l * (C + 1);
    w.arena = arena + (size_t)l * o.max_tokens_per_utt;
  }
  d->d_lanes = Upload(lanes, &d->lane_owned);
  d->lanes_allocated = L;
  CUDA_OK(cudaDeviceSynchronize());
}

rs_decoder *rs_decoder_create(rs_model *m_, rs_graph *g_, const rs_decoder_opts *opts, char *err, size_t errlen) {
  API_GUARD_BEGIN
  ModelImpl *mi = reinterpret_cast<ModelImpl *>(m_);
  GraphImpl *gi = reinterpret_cast<GraphImpl *>(g_);
  if (!mi || !gi) RS_FAIL("rs_decoder_create: model and graph are required");
  if (mi->device != gi->device) RS_FAIL("model and graph live on different devices");
  CUDA_OK(cudaSetDevice(mi->device));
  std::unique_ptr<DecoderImpl> d(new DecoderImpl());
  d->model = mi;
  d->graph = gi;
  if (opts) d->opts = *opts; else rs_decoder_opts_default(&d->opts);
  rs_decoder_opts &o = d->opts;
  if (o.beam <= 0 || o.max_active <= 1 || o.min_active < 0 || o.min_active >= o.max_active)
    RS_FAIL("bad decoder options (beam/max-active/min-active)");
  if (o.acoustic_scale != 1.0f && o.acoustic_scale != mi->m.acoustic_scale) {
    // the scale is folded into the model's output epilogue at load time
    RS_FAIL("acoustic_scale other than 1.0 is not supported (rhasspy always decodes with --acoustic-scale=1.0)");
  }
  if (o.max_tokens_per_frame < 1024) o.max_tokens_per_frame = 1024;
  if (o.max_words < 1) o.max_words = 1;
  CUDA_OK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
  for (auto &e : d->ev) CUDA_OK(cudaEventCreate(&e));
  for (auto &e : d->item_ev) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUDA_OK(cudaMalloc(&d->d_item_flag, sizeof(int)));
  CUDA_OK(cudaMemset(d->d_item_flag, 0, sizeof(int)));
  CUDA_OK(cudaHostAlloc(&d->h_item_seq, sizeof(int), cudaHostAllocDefault));

  BindGraph(d.get(), gi);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, mi->device));
  d->n_lanes = o.num_lanes > 0 ? o.num_lanes : 2 * prop.multiProcessorCount;
  auto dalloc = [&](size_t bytes, int fill) {
    void *p = nullptr;
    CUDA_OK(cudaMalloc(&p, bytes));
    CUDA_OK(cudaMemset(p, fill, bytes));
    d->owned.push_back(p);
    return p;
  };
  d->d_next_utt = (int *)dalloc(sizeof(int), 0);
  d->d_range_flag = (int *)dalloc(sizeof(int), 0);
  CUDA_OK(cudaMallocHost(&d->h_range_flag, sizeof(int)));
  *d->h_range_flag = 0;
  d->slots.resize(mi->m.plan.num_slots);
  CUDA_OK(cudaDeviceSynchronize());
  return reinterpret_cast<rs_decoder *>(d.release());
  API_GUARD_END(nullptr)
}

void rs_decoder_free(rs_decoder *d) { delete reinterpret_cast<DecoderImpl *>(d); }

void *rs_host_alloc(size_t bytes, char *err, size_t errlen) {
  API_GUARD_BEGIN
  void *p = nullptr;
  CUDA_OK(cudaMallocHost(&p, std::max<size_t>(bytes, 16)));
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  g_pinned[reinterpret_cast<uintptr_t>(p)] = std::max<size_t>(bytes, 16);
  return p;
  API_GUARD_END(nullptr)
}

void rs_host_free(void *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned.erase(reinterpret_cast<uintptr_t>(p));
  }
  cudaFreeHost(p);
}

int rs_decoder_set_graph(rs_decoder *d_, rs_graph *g_, char *err, size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  GraphImpl *gi = reinterpret_cast<GraphImpl *>(g_);
  if (!d || !gi) RS_FAIL("rs_decoder_set_graph: decoder and graph are required");
  if (d->model->device != gi->device) RS_FAIL("model and graph live on different devices");
  CUDA_OK(cudaSetDevice(gi->device));
  CUDA_OK(cudaStreamSynchronize(d->stream));
  BindGraph(d, gi);
  d->lat_hdr.clear();
  return 0;
  API_GUARD_END(1)
}

int rs_decoder_set_nbest(rs_decoder *d_, int32_t nbest, float acoustic_scale, char *err, size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  if (!d) RS_FAIL("rs_decoder_set_nbest: null decoder");
  if (nbest < 1 || !(acoustic_scale == acoustic_scale)) RS_FAIL("rs_decoder_set_nbest: n must be >= 1");
  d->nbest = nbest;
  d->nbest_scale = acoustic_scale;
  return 0;
  API_GUARD_END(1)
}

int rs_decoder_set_staging_overlap(rs_decoder *d_, int32_t on, char *err, size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  if (!d) RS_FAIL("rs_decoder_set_staging_overlap: null decoder");
  d->staging_overlap = on != 0;
  return 0;
  API_GUARD_END(1)
}

int rs_debug_lattice_nbest(const int32_t *src, const int32_t *dst, const int32_t *olabel, const float *graph,
                           const float *acoustic, int32_t n_arcs, int32_t n_nodes, int32_t n, float acoustic_scale,
                           int32_t *word_offset, int32_t *word_ids, int32_t max_words, float *cost) {
  if (!src || !dst || !olabel || !graph || !acoustic || !word_offset || !word_ids || !cost || n < 1) return -1;
  try {
    std::vector<LatticeArc> arcs(std::max(n_arcs, 0));
    for (int i = 0; i < n_arcs; i++) arcs[i] = LatticeArc{src[i], dst[i], olabel[i], graph[i], acoustic[i]};
    std::vector<NbestHyp> out;
    LatticeNbest(arcs.data(), n_arcs, n_nodes, n, acoustic_scale, &out);
    int w = 0;
    for (size_t h = 0; h < out.size(); h++) {
      word_offset[h] = w;
      if (w + (int)out[h].words.size() > max_words) return -2;
      for (int id : out[h].words) word_ids[w++] = id;
      cost[2 * h] = out[h].graph;
      cost[2 * h + 1] = out[h].acoustic;
    }
    word_offset[out.size()] = w;
    return (int)out.size();
  } catch (...) {
    return -3;
  }
}

}  // extern "C"

namespace rs {

static rs_result *NewResult(int n) {
  rs_result *r = new rs_result();
  r->n_utts = n;
  r->n_hyp = new int32_t[std::max(n, 1)]();
  r->word_offset = new int32_t[n + 1]();
  r->word_ids = nullptr;
  r->graph_cost = new float[std::max(n, 1)]();
  r->acoustic_cost = new float[std::max(n, 1)]();
  r->num_frames = new int32_t[std::max(n, 1)]();
  r->status = new int32_t[std::max(n, 1)]();
  r->hyp_offset = new int32_t[n + 1]();
  r->hyp_word_offset = n == 0 ? new int32_t[1]() : nullptr;
  r->hyp_word_ids = nullptr;
  r->hyp_graph_cost = nullptr;
  r->hyp_acoustic_cost = nullptr;
  return r;
}

// hypothesis-indexed arrays of a result from per-utterance lists (nb[u] empty: the best path alone, if any)
static void FillHypotheses(rs_result *r, const std::vector<std::vector<NbestHyp>> &nb) {
  const int n = r->n_utts;
  int total = 0, words = 0;
  for (int u = 0; u < n; u++) {
    if (!nb.empty() && !nb[u].empty()) {
      r->n_hyp[u] = (int)nb[u].size();
      for (const NbestHyp &h : nb[u]) words += (int)h.words.size();
    } else {
      words += r->n_hyp[u] ? r->word_offset[u + 1] - r->word_offset[u] : 0;
    }
    total += r->n_hyp[u];
  }
  r->hyp_word_offset = new int32_t[total + 1]();
  r->hyp_word_ids = new int32_t[std::max(words, 1)];
  r->hyp_graph_cost = new float[std::max(total, 1)]();
  r->hyp_acoustic_cost = new float[std::max(total, 1)]();
  int h = 0, w = 0;
  for (int u = 0; u < n; u++) {
    r->hyp_offset[u] = h;
    if (!nb.empty() && !nb[u].empty()) {
      for (const NbestHyp &hy : nb[u]) {
        r->hyp_word_offset[h] = w;
        for (int id : hy.words) r->hyp_word_ids[w++] = id;
        r->hyp_graph_cost[h] = hy.graph;
        r->hyp_acoustic_cost[h] = hy.acoustic;
        h++;
      }
    } else if (r->n_hyp[u]) {
      r->hyp_word_offset[h] = w;
      for (int i = r->word_offset[u]; i < r->word_offset[u + 1]; i++) r->hyp_word_ids[w++] = r->word_ids[i];
      r->hyp_graph_cost[h] = r->graph_cost[u];
      r->hyp_acoustic_cost[h] = r->acoustic_cost[u];
      h++;
    }
  }
  r->hyp_offset[n] = h;
  r->hyp_word_offset[h] = w;
}

// Runs stage (iii) for the batch laid out in d->batch and collects the results.
static rs_result *RunDecodeStage(DecoderImpl *d, const float *loglikes, int ld, int *d_ll_row0, int *d_n_out, int &launches) {
  const int n = d->batch.n;
  const rs_decoder_opts &o = d->opts;
  const int W = o.max_words;
  // outputs: words [n*W], n_words [n], status [n], cost [2n], counters [4n] (u64)
  size_t off_words = 0, off_nw = off_words + sizeof(int) * (size_t)n * W, off_status = off_nw + sizeof(int) * n,
         off_cost = off_status + sizeof(int) * n, off_cnt = RoundUp((int)(off_cost + sizeof(float) * 2 * n), 8),
         out_bytes = off_cnt + sizeof(unsigned long long) * 4 * n;
  char *dout = (char *)d->d_out.ensure(out_bytes);
  DecodeParams p{};
  p.g = d->graph->dev;
  p.g.earc = d->d_earc;
  p.cfg.beam = o.beam;
  p.cfg.beam_delta = o.beam_delta;
  p.cfg.max_active = o.max_active;
  p.cfg.min_active = o.min_active;
  p.cfg.tok_cap = o.max_tokens_per_frame;
  int hash_size = 1;
  while (hash_size < 2 * o.max_tokens_per_frame) hash_size <<= 1;
  p.cfg.hash_size = hash_size;
  p.cfg.arena_cap = o.max_tokens_per_utt;
  p.cfg.max_words = W;
  p.cfg.lattice_beam = o.lattice_beam;
  p.cfg.profile = getenv("RS_B200_DECODE_PROFILE") != nullptr;
  {
    // shared-memory tables when the graph is small enough for two lanes per SM (<= 1024 states)
    int slots = 64;
    while (slots < d->graph->g.num_states) slots <<= 1;
    const char *e = getenv("RS_B200_DECODE_SMEM");
    p.cfg.smem_slots = (slots <= 1024 && !(e && e[0] == '0')) ? slots : 0;
  }
  p.loglikes = loglikes;
  p.ld = ld;
  p.ll_row0 = d_ll_row0;
  p.n_frames = d_n_out;
  p.n_utts = n;
  p.lanes = d->d_lanes;
  p.next_utt = d->d_next_utt;
  p.words = (int *)(dout + off_words);
  p.n_words = (int *)(dout + off_nw);
  p.status = (int *)(dout + off_status);
  p.cost = (float *)(dout + off_cost);
  p.counters = (unsigned long long *)(dout + off_cnt);
  // ---- n-best tail: per-utterance lattice slices inside a byte budget (RS_B200_LATTICE_MB, default 8192).  When a
  // lattice of the batch does not fit, the stage is run again with four times the budget (the log-likelihoods are
  // still resident), up to RS_B200_LATTICE_MAX_MB (default 65536); the grown budget is kept for later calls.
  const bool lattice = d->nbest > 1 || d->nbest_scale != 1.0f;
  // small graphs: the kernel that reproduces the reference's token order (RS_B200_DECODE_EXACT=0 forces the general one)
  bool small = DecodeSmallSupports(p.g);
  if (const char *e = getenv("RS_B200_DECODE_EXACT")) small = small && e[0] != '0';
  LatticeHeader *d_hdr = nullptr;
  LatticeArc *d_arcs = nullptr;
  int arcs_cap = 0;
  const size_t hdr_bytes = lattice ? sizeof(LatticeHeader) * n + sizeof(int) : 0;
  char *hout = (char *)d->h_out.ensure(out_bytes + hdr_bytes);
  for (int attempt = 0;; attempt++) {
    CUDA_OK(cudaMemsetAsync(d->d_next_utt, 0, sizeof(int), d->stream));
    if (lattice) {
      int max_t = 1;
      for (int u = 0; u < n; u++) max_t = std::max(max_t, d->batch.n_out[u]);
      const char *e = getenv("RS_B200_LATTICE_MB");
      d->lattice_mb = std::max<size_t>(d->lattice_mb, (size_t)std::max(e ? atoi(e) : 8192, 16));
      const size_t budget = d->lattice_mb << 20;
      const int kLinksPerToken = 3;
      // tokens {state, cost}, extra cost, new id, 3 links, and a survivor list of a quarter of the links
      const size_t per_tok = sizeof(int2) + sizeof(float) + sizeof(int) + kLinksPerToken * sizeof(int4) * 5 / 4;
      const size_t tc = std::min<size_t>((size_t)o.max_tokens_per_utt, std::max<size_t>(budget / n / per_tok, 4096));
      LatticeBuf &L = p.lat;
      L.tok_cap = (int)tc;
      L.link_cap = (int)std::min<size_t>(tc * kLinksPerToken, 0x7fffffff);
      L.max_t = max_t;
      L.tok = (int2 *)d->d_lat_tok.ensure(sizeof(int2) * tc * n);
      L.extra = (float *)d->d_lat_extra.ensure(sizeof(float) * tc * n);
      L.newid = (int *)d->d_lat_newid.ensure(sizeof(int) * tc * n);
      L.link = (int4 *)d->d_lat_link.ensure(sizeof(int4) * (size_t)L.link_cap * n);
      L.surv_cap = L.link_cap / 4;
      L.surv = (int4 *)d->d_lat_surv.ensure(sizeof(int4) * (size_t)L.surv_cap * n);
      L.tok_base = (int *)d->d_lat_tb.ensure(sizeof(int) * (size_t)(max_t + 2) * n);
      L.link_pos = (int *)d->d_lat_pos.ensure(sizeof(int) * (size_t)(2 * max_t + 4) * n);
      L.cost_offset = (float *)d->d_lat_off.ensure(sizeof(float) * (size_t)(max_t + 1) * n);
      // headers [n] + cursor, then the compact arcs of the pruned lattices
      d_hdr = (LatticeHeader *)d->d_lat_hdr.ensure(sizeof(LatticeHeader) * n + sizeof(int));
      arcs_cap = (int)std::min<size_t>((size_t)L.link_cap * n / 4 + 65536, 64u << 20);
      d_arcs = (LatticeArc *)d->d_lat_arcs.ensure(sizeof(LatticeArc) * (size_t)arcs_cap);
    }
    if (small) {
      // per-utterance slices of one traceback arena: at most one token per state and time
      std::vector<long long> off(n + 1, 0);
      for (int u = 0; u < n; u++) off[u + 1] = off[u] + (long long)(d->batch.n_out[u] + 1) * d->graph->g.num_states;
      long long *h_off = (long long *)d->h_small.ensure(sizeof(long long) * (n + 1));
      memcpy(h_off, off.data(), sizeof(long long) * (n + 1));
      long long *d_off = (long long *)d->d_small_off.ensure(sizeof(long long) * (n + 1));
      CUDA_OK(cudaMemcpyAsync(d_off, h_off, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, d->stream));
      p.small_arena = (int2 *)d->d_small_arena.ensure(sizeof(int2) * (size_t)std::max<long long>(off[n], 1));
      p.small_arena_off = d_off;
      LaunchDecodeSmall(p, d->stream, lattice);
    } else {
      EnsureLaneWorkspace(d, std::min(d->n_lanes, std::max(n, 1)));
      p.lanes = d->d_lanes;
      LaunchDecode(p, std::min(d->lanes_allocated, std::max(n, 1)), d->stream, lattice);
    }
    launches += 1;
    if (lattice) {
      int *d_cursor = reinterpret_cast<int *>(d_hdr + n);
      CUDA_OK(cudaMemsetAsync(d_cursor, 0, sizeof(int), d->stream));
      LaunchLatticePrune(p, o.lattice_beam, d_hdr, d_arcs, arcs_cap, d_cursor, d->stream);
      launches += 1;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(d->ev[4], d->stream));
    CUDA_OK(cudaMemcpyAsync(hout, dout, out_bytes, cudaMemcpyDeviceToHost, d->stream));
    if (lattice) CUDA_OK(cudaMemcpyAsync(hout + out_bytes, d_hdr, hdr_bytes, cudaMemcpyDeviceToHost, d->stream));
    CUDA_OK(cudaStreamSynchronize(d->stream));
    if (!lattice) break;
    {
      const LatticeHeader *hh = reinterpret_cast<const LatticeHeader *>(hout + out_bytes);
      const int *decoded = reinterpret_cast<const int *>(hout + off_nw);
      bool overflow = false;
      for (int u = 0; u < n; u++) overflow |= !hh[u].ok && decoded[u] >= 0;
      const char *em = getenv("RS_B200_LATTICE_MAX_MB");
      const size_t max_mb = (size_t)std::max(em ? atoi(em) : 65536, 16);
      if (!overflow || d->lattice_mb >= max_mb || attempt >= 3) break;
      d->lattice_mb = std::min(d->lattice_mb * 4, max_mb);
    }
  }
  d->last.d2h_bytes = out_bytes + hdr_bytes;
  const LatticeHeader *hdr = reinterpret_cast<const LatticeHeader *>(hout + out_bytes);
  const LatticeArc *harcs = nullptr;
  if (lattice) {
    const int n_arcs = std::min(*reinterpret_cast<const int *>(hout + out_bytes + sizeof(LatticeHeader) * n), arcs_cap);
    harcs = (const LatticeArc *)d->h_lat.ensure(sizeof(LatticeArc) * (size_t)std::max(n_arcs, 1));
    if (n_arcs > 0)
      CUDA_OK(cudaMemcpyAsync((void *)harcs, d_arcs, sizeof(LatticeArc) * (size_t)n_arcs, cudaMemcpyDeviceToHost, d->stream));
    d->last.d2h_bytes += sizeof(LatticeArc) * (size_t)n_arcs;
  }
  CUDA_OK(cudaEventRecord(d->ev[5], d->stream));
  CUDA_OK(cudaStreamSynchronize(d->stream));
  d->host_last_sync_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  const int *words = (const int *)(hout + off_words), *nw = (const int *)(hout + off_nw), *status = (const int *)(hout + off_status);
  const float *cost = (const float *)(hout + off_cost);
  const unsigned long long *cnt = (const unsigned long long *)(hout + off_cnt);
  rs_result *r = NewResult(n);
  int total = 0;
  for (int u = 0; u < n; u++) total += std::max(nw[u], 0);
  r->word_ids = new int32_t[std::max(total, 1)];
  int pos = 0;
  d->last.tokens_expanded = d->last.arcs_visited = d->last.tokens_created = d->last.records_written = 0;
  d->last.frames_decoded = 0;
  for (int u = 0; u < n; u++) {
    r->word_offset[u] = pos;
    r->n_hyp[u] = nw[u] >= 0 ? 1 : 0;
    for (int i = 0; i < nw[u]; i++) r->word_ids[pos++] = words[(size_t)u * W + i];
    r->graph_cost[u] = cost[2 * u];
    r->acoustic_cost[u] = cost[2 * u + 1];
    r->num_frames[u] = d->batch.n_out[u];
    r->status[u] = status[u];
    d->last.tokens_expanded += cnt[4 * (size_t)u];
    d->last.arcs_visited += cnt[4 * (size_t)u + 1];
    d->last.tokens_created += cnt[4 * (size_t)u + 2];
    d->last.records_written += cnt[4 * (size_t)u + 3];
    d->last.frames_decoded += d->batch.n_out[u];
  }
  r->word_offset[n] = pos;
  // ---- order-sensitive utterances (status bit 4): decoded again by the strict-order host decoder from the
  // log-likelihoods that are still resident on the device; its result replaces the device result
  std::vector<int> strict_list;
  std::vector<StrictResult> strict_res;
  d->last.strict_utts = 0;
  d->last.strict_ms = 0.f;
  if (o.strict_fallback) {
    for (int u = 0; u < n; u++)
      if (d->batch.n_out[u] > 0 && !(status[u] & 7) &&
          (o.strict_fallback == 2 || (status[u] & 16) || (lattice && hdr[u].ok && hdr[u].pad[0]))) {
        r->status[u] |= 16 | ((lattice && hdr[u].ok && hdr[u].pad[0]) ? 4096 : 0);
        strict_list.push_back(u);
      }
  }
  if (!o.strict_fallback && lattice)
    for (int u = 0; u < n; u++)
      if (hdr[u].ok && hdr[u].pad[0]) r->status[u] |= 16 | 4096;
  if (!strict_list.empty()) {
    const auto t0 = std::chrono::steady_clock::now();
    const int ns = (int)strict_list.size();
    strict_res.resize(ns);
    // the log-likelihood rows of the flagged utterances, back to back in page-locked memory (into ordinary vectors the
    // 412 MB of a fully flagged batch of 256 were a synchronous, staged copy)
    std::vector<const float *> ll(ns);
    {
      std::vector<size_t> off(ns + 1, 0);
      for (int i = 0; i < ns; i++) off[i + 1] = off[i] + (size_t)d->batch.n_out[strict_list[i]] * ld;
      float *base = (float *)d->h_ll.ensure(std::max<size_t>(off[ns], 1) * sizeof(float));
      for (int i = 0; i < ns; i++) {
        const int u = strict_list[i];
        ll[i] = base + off[i];
        const size_t bytes = (off[i + 1] - off[i]) * sizeof(float);
        if (bytes)
          CUDA_OK(cudaMemcpyAsync(base + off[i], loglikes + (size_t)d->batch.ll_row0[u] * ld, bytes, cudaMemcpyDeviceToHost, d->stream));
        d->last.d2h_bytes += bytes;
      }
    }
    CUDA_OK(cudaStreamSynchronize(d->stream));
    StrictOptions so;
    so.beam = o.beam;
    so.beam_delta = o.beam_delta;
    so.lattice_beam = o.lattice_beam;
    so.max_active = o.max_active;
    so.min_active = o.min_active;
    so.max_words = W;
    if (!d->strict_arcs_ready) {
      BuildStrictArcs(d->graph->g, d->h_epdf.data(), &d->strict_arcs);
      d->strict_arcs_ready = true;
    }
    const int nthreads = std::max(1, std::min(ns, (int)std::thread::hardware_concurrency()));
    // longest utterances first: the threads draw from one counter, and the last draws should be the short ones
    std::vector<int> order(ns);
    for (int i = 0; i < ns; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return d->batch.n_out[strict_list[a]] > d->batch.n_out[strict_list[b]]; });
    std::atomic<int> next{0};
    std::string fail;
    std::mutex fail_mu;
    auto work = [&]() {
      try {
        for (int k = next++; k < ns; k = next++) {
          const int i = order[k];
          StrictDecode(d->graph->g, d->h_epdf.data(), d->strict_arcs, ll[i], ld, d->batch.n_out[strict_list[i]], so, lattice,
                       &strict_res[i]);
        }
      } catch (const std::exception &e) {
        std::lock_guard<std::mutex> lk(fail_mu);
        fail = e.what();
      }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (!fail.empty()) {
      rs_result_free(r);
      RS_FAIL("strict-order decoder: " << fail);
    }
    // splice the words of the re-decoded utterances into the result
    std::vector<int32_t> ids;
    const std::vector<int32_t> old(r->word_offset, r->word_offset + n + 1);
    int si = 0;
    for (int u = 0; u < n; u++) {
      r->word_offset[u] = (int)ids.size();
      if (si < ns && strict_list[si] == u) {
        const StrictResult &sr = strict_res[si++];
        r->n_hyp[u] = sr.decoded ? 1 : 0;
        r->graph_cost[u] = sr.decoded ? sr.graph : 0.f;
        r->acoustic_cost[u] = sr.decoded ? sr.acoustic : 0.f;
        r->status[u] = (r->status[u] & ~(4 | 8)) | 64 | (sr.decoded ? 0 : 4) | (sr.word_overflow ? 8 : 0);
        ids.insert(ids.end(), sr.words.begin(), sr.words.end());
      } else {
        ids.insert(ids.end(), r->word_ids + old[u], r->word_ids + old[u + 1]);
      }
    }
    r->word_offset[n] = (int)ids.size();
    delete[] r->word_ids;
    r->word_ids = new int32_t[std::max<size_t>(ids.size(), 1)];
    std::copy(ids.begin(), ids.end(), r->word_ids);
    d->last.strict_utts = ns;
    d->last.strict_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  }
  std::vector<std::vector<NbestHyp>> nb;
  if (lattice) {
    // host half of the tail: best-first search of each pruned lattice (nbest.cc), utterances across threads
    nb.resize(n);
    const int nthreads = std::max(1, std::min({n, 8, (int)std::thread::hardware_concurrency()}));
    std::atomic<int> next{0};
    std::vector<int> strict_of(n, -1);
    for (size_t i = 0; i < strict_list.size(); i++) strict_of[strict_list[i]] = (int)i;
    auto work = [&]() {
      for (int u = next++; u < n; u = next++) {
        if (r->n_hyp[u] == 0) continue;
        if (strict_of[u] >= 0) {
          const StrictResult &sr = strict_res[strict_of[u]];
          LatticeNbest(sr.lattice.data(), (int)sr.lattice.size(), sr.n_nodes, d->nbest, d->nbest_scale, &nb[u]);
          continue;
        }
        if (!hdr[u].ok) continue;
        LatticeNbest(harcs + hdr[u].arc_begin, hdr[u].n_arcs, hdr[u].n_nodes, d->nbest, d->nbest_scale, &nb[u]);
      }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; i++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    d->lat_hdr.assign(hdr, hdr + n);
    d->strict_lat.clear();
    for (size_t i = 0; i < strict_list.size(); i++) d->strict_lat[strict_list[i]].swap(strict_res[i].lattice);
    d->last.lattice_states = d->last.lattice_arcs = d->last.lattice_links_recorded = 0;
    for (int u = 0; u < n; u++) {
      if (r->n_hyp[u] && nb[u].empty()) r->status[u] |= 32;
      if (hdr[u].ok) {
        d->last.lattice_states += hdr[u].n_nodes;
        d->last.lattice_arcs += hdr[u].n_arcs;
        d->last.lattice_links_recorded += hdr[u].n_links;
      }
    }
  }
  if (!lattice) {
    d->lat_hdr.clear();
    d->strict_lat.clear();
  }
  if (lattice && d->nbest_scale != 1.0f) {
    // the per-utterance fields describe utt-1 of the list: under a ranking scale that need not be the
    // back-traced path of the search (which ran at scale 1)
    std::vector<int32_t> ids;
    const std::vector<int32_t> old(r->word_offset, r->word_offset + n + 1);
    for (int u = 0; u < n; u++) {
      r->word_offset[u] = (int)ids.size();
      if (!nb[u].empty()) {
        ids.insert(ids.end(), nb[u][0].words.begin(), nb[u][0].words.end());
        r->graph_cost[u] = nb[u][0].graph;
        r->acoustic_cost[u] = nb[u][0].acoustic;
      } else if (r->n_hyp[u]) {
        ids.insert(ids.end(), r->word_ids + old[u], r->word_ids + old[u + 1]);
      }
    }
    r->word_offset[n] = (int)ids.size();
    delete[] r->word_ids;
    r->word_ids = new int32_t[std::max<size_t>(ids.size(), 1)];
    std::copy(ids.begin(), ids.end(), r->word_ids);
  }
  FillHypotheses(r, nb);
  return r;
}

static void FinishTimings(DecoderImpl *d, int launches) {
  float ms;
  rs_timings &t = d->last;
  cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]);
  t.h2d_ms = ms;
  cudaEventElapsedTime(&ms, d->ev[1], d->ev[2]);
  t.feature_ms = ms;
  cudaEventElapsedTime(&ms, d->ev[2], d->ev[3]);
  t.nnet_ms = ms;
  cudaEventElapsedTime(&ms, d->ev[3], d->ev[4]);
  t.decode_ms = ms;
  cudaEventElapsedTime(&ms, d->ev[4], d->ev[5]);
  t.d2h_ms = ms;
  cudaEventElapsedTime(&ms, d->ev[0], d->ev[5]);
  t.total_ms = ms;
  t.kernel_launches = launches;
}

// Online decoding (online2-cli-nnet3-decode-faster.cc:139-160): audio arrives in reads of 1024 samples,
// after each read the decoder advances over the nnet chunks that have become ready
// (decodable-online-looped.cc:56-84) and every chunk is computed with the iVector estimated from the
// frames available at that moment (:186-194, online-ivector-feature.cc:248-279, 327-355).  The sequence is
// a pure function of the sample count.  Returns, for one utterance, the frame counts of the successive
// iVector solves and, per nnet chunk, the index of the solve it uses.
struct OnlineSchedule {
  std::vector<int> solve_frames;  // solve j covers frames [0, solve_frames[j])  (0 = no data: zero iVector)
  std::vector<int> chunk_solve;   // per chunk
};
static OnlineSchedule ComputeOnlineSchedule(int nsamp, int T, int length, int shift, int chunk, int right_context,
                                            int splice_right, int sf) {
  OnlineSchedule s;
  int chunks_done = 0, stats_frames = 0;
  auto advance = [&](int F, int iv_ready, int ready_chunks) {
    while (chunks_done < ready_chunks) {
      int frames = 0;
      if (iv_ready > 0) frames = std::min(F - 1, iv_ready - 1) + 1;
      if (s.solve_frames.empty() || frames > stats_frames) {  // GetFrame processes new frames, then one CG
        s.solve_frames.push_back(frames);
        stats_frames = frames;
      }
      s.chunk_solve.push_back((int)s.solve_frames.size() - 1);
      chunks_done++;
    }
  };
  for (long long got = 0; got < nsamp;) {
    got = std::min<long long>(got + 1024, nsamp);
    const int F = got < length ? 0 : 1 + (int)((got - length) / shift);
    if (F == 0) continue;
    advance(F, std::max(0, F - splice_right), std::max(0, F - right_context) / chunk);
  }
  if (T > 0) {  // InputFinished: every frame is ready, the tail is padded with the last frame
    const int out_per_chunk = chunk / sf, total_out = (T + sf - 1) / sf;
    advance(T, T, (total_out + out_per_chunk - 1) / out_per_chunk);
  }
  return s;
}

// `fill`, when given, writes utterance u's samples straight into the pinned staging area (rs_decode_wavs reads the files
// there: no intermediate copy); pcm[] is then unused.
using FillFn = std::function<void(int, int16_t *)>;
static rs_result *DecodePcm(DecoderImpl *d, const int16_t *const *pcm, const int32_t *nsamp, int n, bool online = false,
                            const FillFn *fill = nullptr) {
  ModelImpl *mi = d->model;
  const Model &m = mi->m;
  const Plan &pl = m.plan;
  CUDA_OK(cudaSetDevice(mi->device));
  if (n < 0) RS_FAIL("negative batch size");
  d->last = rs_timings{};
  if (n == 0) return NewResult(0);
  const bool host_prof = getenv("RS_B200_HOST_PROFILE") != nullptr;
  auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double hp0 = now_ms();
  for (int i = 0; i < n; i++)
    if (nsamp[i] < 0 || (nsamp[i] > 0 && !fill && !pcm[i])) RS_FAIL("utterance " << i << ": bad sample buffer");
  const int sf = m.frame_subsampling_factor, D = m.mfcc.num_ceps;
  const int shift = m.mfcc.WindowShift(), length = m.mfcc.WindowSize();
  auto &B = d->batch;
  B = DecoderImpl::Batch();
  B.n = n;
  B.num_frames.resize(n);
  B.frame_offset.resize(n);
  B.origin.resize(n);
  B.n_out.resize(n);
  B.ll_row0.resize(n);
  std::vector<int64_t> pcm_offset(n);
  int64_t total_samples = 0;
  int total_frames = 0, max_frames = 0;
  const int L = pl.left_context, R = pl.right_context, align = pl.align;
  int axis = RoundUp(L, align);
  double audio_s = 0.0;
  for (int u = 0; u < n; u++) {
    pcm_offset[u] = total_samples;
    total_samples += nsamp[u];
    // NumFrames with snip_edges (feature-window.cc:30-50)
    int T = nsamp[u] < length ? 0 : 1 + (nsamp[u] - length) / shift;
    B.num_frames[u] = T;
    B.frame_offset[u] = total_frames;
    total_frames += T;
    max_frames = std::max(max_frames, T);
    B.origin[u] = axis;
    B.n_out[u] = (T + sf - 1) / sf;  // decodable-online-looped.cc:71-73
    B.ll_row0[u] = axis / sf;
    axis += RoundUp(T + R + L, align);
    audio_s += nsamp[u] / (double)m.mfcc.samp_freq;
  }
  const int axis_len = axis + align;
  B.total_frames = total_frames;
  B.axis_len = axis_len;
  d->last.audio_seconds = audio_s;
  // ---- iVector solves: one per utterance (offline, online2-wav-nnet3-latgen-faster --online=false) or
  // one per nnet chunk that saw new frames (online, the stream binary)
  int chunk = m.frames_per_chunk;  // GetChunkSize (nnet-compile-looped.cc:81-94), modulus 1 for TDNN models
  while (chunk % sf) chunk++;
  std::vector<int> v_num_frames, v_frame_offset;
  std::vector<OnlineSchedule> sched(online && m.has_ivector ? n : 0);
  B.v_begin.assign(n + 1, 0);
  for (int u = 0; u < n; u++) {
    B.v_begin[u] = (int)v_num_frames.size();
    if (!sched.empty()) {
      sched[u] = ComputeOnlineSchedule(nsamp[u], B.num_frames[u], length, shift, chunk, R, m.ivec.splice_right, sf);
      if (sched[u].solve_frames.empty()) sched[u].solve_frames.push_back(0);
      for (int f : sched[u].solve_frames) {
        v_num_frames.push_back(f);
        v_frame_offset.push_back(B.frame_offset[u]);
      }
    } else {
      v_num_frames.push_back(B.num_frames[u]);
      v_frame_offset.push_back(B.frame_offset[u]);
    }
  }
  const int v_n = (int)v_num_frames.size();
  B.v_begin[n] = v_n;
  // ---- host staging: pcm + descriptor block, one H2D copy each
  const size_t pcm_bytes = sizeof(int16_t) * (size_t)std::max<int64_t>(total_samples, 1);
  // descriptor: pcm_offset[n] (i64) | num_frames | frame_offset | origin | n_out | ll_row0 | row_utt[axis_len]
  //             | v_num_frames[v_n] | v_frame_offset[v_n] | v_begin[n + 1]
  const size_t desc_ints = (size_t)2 * n + 5 * (size_t)n + axis_len + 2 * (size_t)v_n + n + 1;
  char *hin = (char *)d->h_in.ensure(pcm_bytes + 16 + desc_ints * sizeof(int));
  // Audio that already sits back to back in page-locked memory from rs_host_alloc goes to the device straight
  // from the caller's buffer; anything else is first packed into the decoder's own pinned staging area.
  const int16_t *direct_base = nullptr;
  {
    bool contiguous = total_samples > 0 && !fill;
    const int16_t *base = nullptr;
    for (int u = 0; u < n && contiguous; u++) {
      if (!nsamp[u]) continue;
      if (!base) base = pcm[u] - pcm_offset[u];
      contiguous = pcm[u] == base + pcm_offset[u];
    }
    if (contiguous && base && HostRangeIsPinned(base, sizeof(int16_t) * (size_t)total_samples)) direct_base = base;
  }
  int16_t *hpcm = direct_base ? const_cast<int16_t *>(direct_base) : (int16_t *)hin;
  // Staging is split into items of ~2 MB packed by a small persistent thread pool; each item goes to
  // the device as soon as it is complete, so the H2D copies overlap the packing of the later items.
  static const int item_shift = getenv("RS_B200_PACK_ITEM_SHIFT") ? atoi(getenv("RS_B200_PACK_ITEM_SHIFT")) : 21;
  int n_items = (int)std::min<size_t>(std::max<size_t>(pcm_bytes >> item_shift, 1), 64);
  // page-locked caller memory: nothing to pack, but the copy is still cut into a few items so that the MFCC kernel of
  // one item runs under the copy of the next (RS_B200_DIRECT_ITEMS, default 4 items of at least 2 MB: the MFCC
  // kernel of an eighth of batch 256 is a wave and a third of warps, and eight of those take longer than the copies)
  static const int direct_items = getenv("RS_B200_DIRECT_ITEMS") ? std::max(1, std::min(kMaxStagingItems, atoi(getenv("RS_B200_DIRECT_ITEMS")))) : 4;
  if (direct_base) n_items = d->staging_overlap ? std::min(n_items, direct_items) : 1;
  if (n_items > n) n_items = n;
  std::vector<int> range_begin(n_items + 1, n);
  {
    int u = 0;
    for (int w = 0; w < n_items; w++) {
      range_begin[w] = u;
      const int64_t target = total_samples * (w + 1) / n_items;
      while (u < n && pcm_offset[u] + nsamp[u] <= target) u++;
      if (w == n_items - 1) u = n;
    }
    range_begin[n_items] = n;
  }
  std::string pack_err;  // a failing file read must not escape a staging thread
  std::mutex pack_err_mu;
  auto pack = [&](int w) {
    if (direct_base) return;
    try {
      for (int u = range_begin[w]; u < range_begin[w + 1]; u++)
        if (nsamp[u]) {
          if (fill) (*fill)(u, hpcm + pcm_offset[u]);
          else memcpy(hpcm + pcm_offset[u], pcm[u], sizeof(int16_t) * (size_t)nsamp[u]);
        }
    } catch (const std::exception &e) {
      std::lock_guard<std::mutex> lk(pack_err_mu);
      if (pack_err.empty()) pack_err = e.what();
    }
  };
  const bool pooled = n_items > 1 && !direct_base;  // nothing to pack from page-locked caller memory: no worker wake-ups
  if (pooled && !d->pool) {
    // RS_B200_PACK_THREADS: worker threads of the staging pool; the caller's thread packs too.  Measured on the
    // 16-core B200 host: 3-4 workers are the optimum (1.3 ms for 33 MB), 6 and more are slower (1.8-2.9 ms)
    int nt = (int)std::min(3u, std::max(2u, std::thread::hardware_concurrency()) - 1);
    if (const char *e = getenv("RS_B200_PACK_THREADS")) nt = std::max(1, std::min(32, atoi(e)));
    d->pool.reset(new PackPool(nt));
  }
  struct PoolGuard {  // an error path must not leave workers running on this frame's captures
    PackPool *p;
    ~PoolGuard() {
      if (p) p->WaitAll();
    }
  } pool_guard{pooled ? d->pool.get() : nullptr};
  if (pooled) d->pool->Start(n_items, pack);
  size_t desc_off = (pcm_bytes + 15) & ~(size_t)15;
  int *hdesc = (int *)(hin + desc_off);
  memcpy(hdesc, pcm_offset.data(), sizeof(int64_t) * n);
  int *h_nf = hdesc + 2 * n, *h_fo = h_nf + n, *h_or = h_fo + n, *h_no = h_or + n, *h_r0 = h_no + n, *h_ru = h_r0 + n;
  for (int u = 0; u < n; u++) {
    h_nf[u] = B.num_frames[u];
    h_fo[u] = B.frame_offset[u];
    h_or[u] = B.origin[u];
    h_no[u] = B.n_out[u];
    h_r0[u] = B.ll_row0[u];
  }
  {  // row of the per-utterance buffers (iVector, its bias contribution) that axis time t uses
    int u = 0;
    const int lag = (R + chunk - 1) / chunk;  // iVector time k * chunk first appears in chunk max(0, k - lag)
    if (sched.empty()) {
      // one iVector per utterance: utterance u owns the axis rows [origin[u] - L, origin[u + 1] - L) (the first and
      // the last one also the margins), filled range by range instead of row by row (0.2 ms of every call at batch 256)
      int t0 = 0;
      for (u = 0; u < n; u++) {
        const int t1 = u + 1 < n ? std::max(t0, std::min(B.origin[u + 1] - L, axis_len)) : axis_len;
        std::fill(h_ru + t0, h_ru + t1, B.v_begin[u]);
        t0 = t1;
      }
    } else
    for (int t = 0; t < axis_len; t++) {
      while (u + 1 < n && t >= B.origin[u + 1] - L) u++;
      int row = B.v_begin[u];
      if (!sched.empty() && !sched[u].chunk_solve.empty()) {
        const int tl = t - B.origin[u];
        const int k = tl >= 0 ? tl / chunk : -((-tl + chunk - 1) / chunk);  // t - Mod(t, period), nnet-compile-looped.cc:188-208
        const int c = std::min(std::max(k - lag, 0), (int)sched[u].chunk_solve.size() - 1);
        row += sched[u].chunk_solve[c];
      }
      h_ru[t] = row;
    }
    int *h_v = h_ru + axis_len;
    memcpy(h_v, v_num_frames.data(), sizeof(int) * v_n);
    memcpy(h_v + v_n, v_frame_offset.data(), sizeof(int) * v_n);
    memcpy(h_v + 2 * v_n, B.v_begin.data(), sizeof(int) * (n + 1));
  }
  int16_t *dpcm = (int16_t *)d->d_pcm.ensure(pcm_bytes);
  int *ddesc = (int *)d->d_desc.ensure(desc_ints * sizeof(int));
  const double hp_ev0 = now_ms();
  CUDA_OK(cudaEventRecord(d->ev[0], d->stream));
  CUDA_OK(cudaMemcpyAsync(ddesc, hdesc, desc_ints * sizeof(int), cudaMemcpyHostToDevice, d->stream));
  const int64_t *d_pcm_off = (const int64_t *)ddesc;
  int *d_nf = ddesc + 2 * n, *d_fo = d_nf + n, *d_or = d_fo + n, *d_no = d_or + n, *d_r0 = d_no + n, *d_ru = d_r0 + n;
  int launches = 0;
  const size_t feat_elems = (size_t)std::max(total_frames, 1) * D;
  float *d_mfcc = (float *)d->d_mfcc.ensure(feat_elems * sizeof(float));
  FeatParams fp = mi->feat;
  fp.pcm = dpcm;
  fp.mfcc = d_mfcc;
  fp.seed = d->opts.dither_seed;
  // The audio copies of a staged batch go through a second stream, item by item; the MFCC kernel of an item waits
  // for that item's copy only, so it runs while the next item is still on the bus (and, for pageable input, while
  // the host packs the later ones).  RS_B200_OVERLAP_STAGING=0 keeps copies and kernels on one stream (the stage
  // times of rs_timings are then cleanly separated: with the overlap, h2d_ms ends when the last copy has landed and
  // contains the MFCC kernels of the earlier items).
  static const bool overlap_env = !(getenv("RS_B200_OVERLAP_STAGING") && getenv("RS_B200_OVERLAP_STAGING")[0] == '0');
  const bool overlap = overlap_env && d->staging_overlap && n_items > 1;
  std::vector<cudaEvent_t> prof_ev;  // RS_B200_HOST_PROFILE: a mark behind every copy and every MFCC kernel
  std::string prof_tag;
  // How the MFCC kernel of an item learns that its audio has landed: the HOST waits for the item's copy event and then
  // launches the kernel without any stream dependency.  Two device-side forms were measured and dropped: a cross-stream
  // cudaStreamWaitEvent (the kernels then start only when the copy stream has run dry -- copies + three quarter-batch
  // kernels 1.13 ms, against 0.62 + 0.45 one after the other), and kernels that poll a flag word written behind the
  // samples (0.90 ms, but a spinning grid that fills the SMs deadlocks as soon as its copy shares a hardware queue with
  // another decoder's kernels: eight engines per device in the pool).  RS_B200_STAGING_SYNC=event keeps the first form.
  static const bool host_sync_env = !(getenv("RS_B200_STAGING_SYNC") && !strcmp(getenv("RS_B200_STAGING_SYNC"), "event"));
  const bool host_sync = overlap && host_sync_env;
  static const bool spin_env = !(getenv("RS_B200_STAGING_SPIN") && getenv("RS_B200_STAGING_SPIN")[0] == '0');
  const bool spin = host_sync && spin_env;
  if (overlap) CUDA_OK(cudaStreamWaitEvent(d->copy_stream, d->ev[0], 0));
  if (spin) {
    d->staging_seq = d->staging_seq == 0x7fffffff ? 1 : d->staging_seq + 1;
    staging_spin_kernel<<<1, 32, 0, d->stream>>>(d->d_item_flag, d->staging_seq, 2000000ll /* ~1 ms */);
    launches++;
  }
  auto prof_mark = [&](cudaStream_t st) {
    if (!host_prof) return;
    cudaEvent_t e;
    CUDA_OK(cudaEventCreate(&e));
    CUDA_OK(cudaEventRecord(e, st));
    prof_ev.push_back(e);
    prof_tag.push_back(st == d->stream ? 'k' : 'c');
  };
  auto queue_copy = [&](int w) {
    const int u0 = range_begin[w], u1 = range_begin[w + 1];
    const int64_t s0 = u0 < n ? pcm_offset[u0] : total_samples;
    const int64_t s1 = u1 < n ? pcm_offset[u1] : total_samples;
    cudaStream_t cs = overlap ? d->copy_stream : d->stream;
    if (s1 > s0) CUDA_OK(cudaMemcpyAsync(dpcm + s0, hpcm + s0, sizeof(int16_t) * (size_t)(s1 - s0), cudaMemcpyHostToDevice, cs));
    if (overlap) {
      CUDA_OK(cudaEventRecord(d->item_ev[w], d->copy_stream));
      if (spin && w == 0) {
        *d->h_item_seq = d->staging_seq;
        CUDA_OK(cudaMemcpyAsync(d->d_item_flag, d->h_item_seq, sizeof(int), cudaMemcpyHostToDevice, d->copy_stream));
      }
      prof_mark(d->copy_stream);
    }
  };
  auto launch_item = [&](int w) {  // overlap only: the MFCC kernel of item w, once its copy is known to be done
    if (host_sync) CUDA_OK(cudaEventSynchronize(d->item_ev[w]));
    else CUDA_OK(cudaStreamWaitEvent(d->stream, d->item_ev[w], 0));
    if (w == n_items - 1) CUDA_OK(cudaEventRecord(d->ev[1], d->stream));
    const int u0 = range_begin[w], u1 = range_begin[w + 1];
    if (u1 > u0) {
      int mf = 0;
      for (int u = u0; u < u1; u++) mf = std::max(mf, B.num_frames[u]);
      fp.pcm_offset = d_pcm_off + u0;
      fp.num_frames = d_nf + u0;
      fp.frame_offset = d_fo + u0;
      LaunchMfcc(fp, u1 - u0, mf, d->stream);
      launches++;
    }
    prof_mark(d->stream);
  };
  if (overlap && !pooled) {
    // nothing to pack (page-locked caller memory): every copy is queued at once, the kernels follow them one by one
    for (int w = 0; w < n_items; w++) queue_copy(w);
    for (int w = 0; w < n_items; w++) launch_item(w);
  } else {
    for (int w = 0; w < n_items; w++) {
      if (pooled) d->pool->WaitItem(w);
      else pack(w);
      queue_copy(w);
      if (overlap && w > 0) launch_item(w - 1);  // while the copy of item w is on the bus
    }
    if (overlap) launch_item(n_items - 1);
  }
  if (!overlap) CUDA_OK(cudaEventRecord(d->ev[1], d->stream));
  {
    std::lock_guard<std::mutex> lk(pack_err_mu);
    if (!pack_err.empty()) {
      CUDA_OK(cudaStreamSynchronize(d->stream));
      RS_FAIL(pack_err);
    }
  }
  const double hp1 = now_ms();
  d->last.h2d_bytes = pcm_bytes + desc_ints * sizeof(int);
  // ---- stage (i)
  if (!overlap) {
    fp.pcm_offset = d_pcm_off;
    fp.num_frames = d_nf;
    fp.frame_offset = d_fo;
    LaunchMfcc(fp, n, max_frames, d->stream);
    launches++;
  }
  const float *nnet_feats = d_mfcc;
  float *d_mfcc_norm = nullptr;
  if (m.has_ivector || m.nnet_cmvn) d_mfcc_norm = (float *)d->d_mfcc_norm.ensure(feat_elems * sizeof(float));
  float *d_ivector = nullptr;
  int ivector_ld = 0;
  // slot storage for the plan
  auto slot_ptr = [&](int buffer) -> float * { return d->slots[pl.buffers[buffer].slot].as<float>(); };
  // row pitch in elements: fp32 buffers are padded to 16 bytes, split (2 x fp16 plane) buffers too
  auto buf_ld = [&](int buffer) { return RoundUp(pl.buffers[buffer].dim, mi->split[buffer] ? 8 : 4); };
  auto buf_rows = [&](int buffer) { return pl.buffers[buffer].per_utt ? v_n : axis_len / pl.buffers[buffer].step; };
  // split buffers hold two planes back to back: hi at slot_ptr, lo right behind it
  auto slot_lo = [&](int buffer) -> __half * {
    return mi->split[buffer] ? reinterpret_cast<__half *>(slot_ptr(buffer)) + (size_t)buf_rows(buffer) * buf_ld(buffer) : nullptr;
  };
  {
    std::vector<size_t> need(pl.num_slots, 0);
    for (size_t b = 0; b < pl.buffers.size(); b++) {
      size_t bytes = (size_t)buf_rows((int)b) * buf_ld((int)b) * 4;  // one fp32 plane or two fp16 planes
      need[pl.buffers[b].slot] = std::max(need[pl.buffers[b].slot], bytes);
    }
    for (int s = 0; s < pl.num_slots; s++) d->slots[s].ensure(std::max<size_t>(need[s], 16));
  }
  if (m.has_ivector) {
    CmvnParams c{};
    c.in = d_mfcc;
    c.out = d_mfcc_norm;
    c.num_frames = d_nf;
    c.frame_offset = d_fo;
    c.global_stats = mi->d_global_cmvn;
    c.dim = D;
    c.cmn_window = m.ivec.cmvn.cmn_window;
    c.global_frames = m.ivec.cmvn.global_frames;
    c.normalize_mean = m.ivec.cmvn.normalize_mean;
    c.normalize_variance = m.ivec.cmvn.normalize_variance;
    LaunchCmvn(c, n, d->stream);
    launches++;
    IvecParams iv = mi->ivec;
    const int G = iv.num_gauss, LD = iv.ldim, Rv = iv.ivector_dim, P = Rv * (Rv + 1) / 2;
    iv.mfcc = d_mfcc;
    iv.mfcc_norm = d_mfcc_norm;
    iv.num_frames = d_nf;
    iv.frame_offset = d_fo;
    iv.n_utts = n;
    iv.v_n = v_n;
    iv.v_num_frames = d_ru + axis_len;
    iv.v_frame_offset = d_ru + axis_len + v_n;
    iv.v_begin = d_ru + axis_len + 2 * v_n;
    iv.total_frames = total_frames;
    iv.max_frames = max_frames;
    const size_t tf = std::max(total_frames, 1);
    iv.x_raw = (float *)d->d_xraw.ensure(tf * LD * sizeof(float));
    iv.x_norm = (float *)d->d_xnorm.ensure(tf * LD * sizeof(float));
    iv.post_idx = (int *)d->d_post_idx.ensure(tf * iv.num_gselect * sizeof(int));
    iv.post_w = (float *)d->d_post_w.ensure(tf * iv.num_gselect * sizeof(float));
    iv.wf = (double *)d->d_wf.ensure((size_t)v_n * G * LD * sizeof(double));
    iv.gw = (float *)d->d_gw.ensure((size_t)v_n * G * sizeof(float));
    iv.linear_chunks = (G * LD + 511) / 512;
    iv.linear_part = (double *)d->d_linear.ensure((size_t)iv.linear_chunks * v_n * Rv * sizeof(double));
    iv.quad = (double *)d->d_quad.ensure((size_t)v_n * P * sizeof(double));
    d_ivector = slot_ptr(pl.ivector_buffer);
    ivector_ld = buf_ld(pl.ivector_buffer);
    iv.ivector = d_ivector;
    iv.ivector_ld = ivector_ld;
    LaunchIvector(iv, d->stream);
    launches += 8;
  }
  if (m.nnet_cmvn) {
    // --cmvn-config in online.conf: the nnet input is CMVN-normalised too
    // (online-nnet2-feature-pipeline.cc:90-147); it may use different options than the iVector one
    CmvnParams c{};
    c.in = d_mfcc;
    c.out = d_mfcc_norm;
    c.num_frames = d_nf;
    c.frame_offset = d_fo;
    c.global_stats = mi->d_nnet_global_cmvn;
    c.dim = D;
    c.cmn_window = m.nnet_cmvn_opts.cmn_window;
    c.global_frames = m.nnet_cmvn_opts.global_frames;
    c.normalize_mean = m.nnet_cmvn_opts.normalize_mean;
    c.normalize_variance = m.nnet_cmvn_opts.normalize_variance;
    LaunchCmvn(c, n, d->stream);
    launches++;
    nnet_feats = d_mfcc_norm;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(d->ev[2], d->stream));
  // ---- stage (ii)
  {
    float *in = slot_ptr(pl.input_buffer);
    const int ild = buf_ld(pl.input_buffer);
    CUDA_OK(cudaMemsetAsync(in, 0, (size_t)axis_len * ild * 4, d->stream));
    CUDA_OK(cudaMemsetAsync(d->d_range_flag, 0, sizeof(int), d->stream));
    AssembleParams a{};
    a.dst_lo = slot_lo(pl.input_buffer);
    a.range_flag = d->d_range_flag;
    a.feats = nnet_feats;
    a.num_frames = d_nf;
    a.frame_offset = d_fo;
    a.origin = d_or;
    a.dst = in;
    a.dim = D;
    a.ld = ild;
    a.left = L;
    a.right = R;
    a.axis_len = axis_len;
    LaunchAssembleInput(a, n, max_frames + L + R, d->stream);
    launches++;
  }
  for (size_t step_i = 0; step_i < pl.steps.size(); step_i++) {
    const Step &st = pl.steps[step_i];
    GemmParams g{};
    const PlanBuffer &ob = pl.buffers[st.out];
    g.out = slot_ptr(st.out);
    g.out_lo = slot_lo(st.out);
    g.range_flag = d->d_range_flag;
    g.out_ld = buf_ld(st.out);
    g.m = buf_rows(st.out);
    g.n = st.n;
    g.out_step = ob.per_utt ? 1 : ob.step;
    g.row_utt = d_ru;
    g.n_slabs = (int)st.slabs.size();
    for (int s = 0; s < g.n_slabs; s++) {
      const Slab &sl = st.slabs[s];
      const PlanBuffer &sb = pl.buffers[sl.src];
      GemmSlab &gs = g.slabs[s];
      gs.src = slot_ptr(sl.src);
      gs.src_lo = slot_lo(sl.src);
      gs.ld = buf_ld(sl.src);
      gs.rows = buf_rows(sl.src);
      gs.k = sl.k;
      gs.wcol = sl.wcol;
      if (sb.per_utt) {
        gs.num = 1;
        gs.den = 1;
        gs.shift = 0;
      } else {
        gs.num = ob.step;
        gs.den = sb.step;
        gs.shift = sl.t_offset;
      }
    }
    g.n_ops = (int)st.ops.size();
    for (int i = 0; i < g.n_ops; i++) {
      const EpiOp &op = st.ops[i];
      DevOp &dop = g.ops[i];
      dop.type = op.type;
      dop.v0 = op.vec0 >= 0 ? mi->d_vectors[op.vec0] : nullptr;
      dop.v1 = op.vec1 >= 0 ? mi->d_vectors[op.vec1] : nullptr;
      dop.alpha = op.alpha;
      dop.buf = nullptr;
      dop.num = dop.den = 1;
      if (op.buffer >= 0) {
        dop.buf = slot_ptr(op.buffer);
        dop.buf_lo = slot_lo(op.buffer);
        dop.buf_ld = buf_ld(op.buffer);
        dop.buf_rows = buf_rows(op.buffer);
        if (op.type == EpiOp::kAddScaled) {
          dop.num = ob.step;
          dop.den = pl.buffers[op.buffer].step;
        } else {
          dop.num = ob.step;  // kUttBias: axis time of an output row
        }
      }
    }
    switch (st.type) {
      case Step::kGemm:
      case Step::kUttGemm:
        if (mi->tc_steps[step_i].ok) {
          // tcgen05 path: each slab is a (row-strided, row-shifted) TMA view of its source planes
          const ModelImpl::TcStep &ts = mi->tc_steps[step_i];
          TcParams t{};
          t.w_hi = ts.map_hi;
          t.w_lo = ts.map_lo;
          t.n_slabs = g.n_slabs;
          t.bn = ts.bn;
          for (int s = 0; s < g.n_slabs; s++) {
            const Slab &sl = st.slabs[s];
            const PlanBuffer &sb = pl.buffers[sl.src];
            const int stride = ob.step / sb.step, sh = sl.t_offset / sb.step;
            const int rem = ((sh % stride) + stride) % stride;
            const int ld = buf_ld(sl.src), rows = buf_rows(sl.src);
            const long long view_rows = rows > rem ? (rows - 1 - rem) / stride + 1 : 0;
            TcEncodeMap(&t.a_hi[s], reinterpret_cast<const __half *>(slot_ptr(sl.src)) + (size_t)rem * ld, view_rows, sl.k,
                        (long long)stride * ld, kTcBM);
            TcEncodeMap(&t.a_lo[s], slot_lo(sl.src) + (size_t)rem * ld, view_rows, sl.k, (long long)stride * ld, kTcBM);
            t.slabs[s].kblocks = (sl.k + kTcBK - 1) / kTcBK;
            t.slabs[s].wk0 = ts.k0[s];
            t.slabs[s].yshift = (sh - rem) / stride;
          }
          t.out_hi = g.out;
          t.out_lo = g.out_lo;
          t.out_ld = g.out_ld;
          t.m = g.m;
          t.n = g.n;
          t.n_ops = g.n_ops;
          for (int i = 0; i < g.n_ops; i++) t.ops[i] = g.ops[i];
          t.row_utt = g.row_utt;
          t.range_flag = d->d_range_flag;
          TcConfigure(&t);
          LaunchGemmTc(t, mi->num_sms, d->stream);
          break;
        }
        g.w = mi->d_matrices[st.weight];
        g.ktot = st.ktot;
        LaunchGemm(g, d->stream);
        break;
      case Step::kElementwise:
        for (int s = 0; s < g.n_slabs; s++) g.slabs[s].wcol = 0;
        LaunchElementwise(g, st.term_scale.data(), st.col_offset, d->stream);
        break;
      case Step::kLogSoftmax:
        if (g.out_lo) RS_FAIL("internal: a log-softmax output feeding a tensor-core layer is not supported");
        LaunchLogSoftmax(g.slabs[0].src, g.slabs[0].src_lo, g.slabs[0].ld, g, d->stream);
        break;
    }
    launches++;
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(d->h_range_flag, d->d_range_flag, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  CUDA_OK(cudaEventRecord(d->ev[3], d->stream));
  d->last.nnet_flops = (uint64_t)((double)mi->flops_per_axis_unit / 1000.0 * axis_len);
  d->last.nnet_bytes = (uint64_t)((double)mi->bytes_per_axis_unit / 1000.0 * axis_len);
  // ---- stage (iii)
  B.loglikes = slot_ptr(pl.output_buffer);
  B.ll_ld = buf_ld(pl.output_buffer);
  const double hp2 = now_ms();
  rs_result *r = RunDecodeStage(d, B.loglikes, B.ll_ld, d_r0, d_no, launches);
  FinishTimings(d, launches);
  if (host_prof) {
    fprintf(stderr, "host ms: layout before the first event %.3f | pack+h2d issue %.3f | feature+nnet launches %.3f | decode launch+wait+result %.3f (after the last sync %.3f)\n",
            hp_ev0 - hp0, hp1 - hp_ev0, hp2 - hp1, now_ms() - hp2, now_ms() - d->host_last_sync_ms);
    if (!prof_ev.empty()) {
      fprintf(stderr, "staging marks in host order (ms after the call's first event; c = copy stream, k = kernel stream): ");
      for (size_t i = 0; i < prof_ev.size(); i++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, d->ev[0], prof_ev[i]);
        fprintf(stderr, "%c %.3f ", prof_tag[i], ms);
        cudaEventDestroy(prof_ev[i]);
      }
      fprintf(stderr, "\n");
    }
  }
  if (*d->h_range_flag) {
    rs_result_free(r);
    RS_FAIL("an activation exceeded the fp16 range (+-65504) of the tensor-core path; set RS_B200_GEMM=simt for this model");
  }
  return r;
}

static rs_result *DecodeLoglikes(DecoderImpl *d, const float *const *ll, const int32_t *nframes, int n) {
  ModelImpl *mi = d->model;
  CUDA_OK(cudaSetDevice(mi->device));
  d->last = rs_timings{};
  if (n <= 0) return NewResult(0);
  const int P = mi->m.trans.num_pdfs, ld = RoundUp(P, 4);
  auto &B = d->batch;
  B = DecoderImpl::Batch();
  B.n = n;
  B.from_loglikes = true;
  B.n_out.resize(n);
  B.ll_row0.resize(n);
  B.num_frames.assign(n, 0);
  B.frame_offset.assign(n, 0);
  B.origin.assign(n, 0);
  size_t rows = 0;
  for (int u = 0; u < n; u++) {
    if (nframes[u] < 0) RS_FAIL("negative frame count");
    B.ll_row0[u] = (int)rows;
    B.n_out[u] = nframes[u];
    rows += nframes[u];
  }
  const size_t ll_bytes = std::max<size_t>(rows, 1) * ld * sizeof(float);
  char *hin = (char *)d->h_in.ensure(ll_bytes + 2 * sizeof(int) * n);
  float *hll = (float *)hin;
  memset(hll, 0, ll_bytes);
  for (int u = 0; u < n; u++)
    for (int t = 0; t < nframes[u]; t++)
      memcpy(hll + ((size_t)B.ll_row0[u] + t) * ld, ll[u] + (size_t)t * P, sizeof(float) * P);
  int *hdesc = (int *)(hin + ll_bytes);
  for (int u = 0; u < n; u++) {
    hdesc[u] = B.ll_row0[u];
    hdesc[n + u] = B.n_out[u];
  }
  float *dll = (float *)d->d_loglikes_ext.ensure(ll_bytes);
  int *ddesc = (int *)d->d_desc.ensure(2 * sizeof(int) * n);
  CUDA_OK(cudaEventRecord(d->ev[0], d->stream));
  CUDA_OK(cudaMemcpyAsync(dll, hll, ll_bytes, cudaMemcpyHostToDevice, d->stream));
  CUDA_OK(cudaMemcpyAsync(ddesc, hdesc, 2 * sizeof(int) * n, cudaMemcpyHostToDevice, d->stream));
  for (int k = 1; k <= 3; k++) CUDA_OK(cudaEventRecord(d->ev[k], d->stream));
  d->last.h2d_bytes = ll_bytes;
  int launches = 0;
  B.loglikes = dll;
  B.ll_ld = ld;
  rs_result *r = RunDecodeStage(d, dll, ld, ddesc, ddesc + n, launches);
  FinishTimings(d, launches);
  return r;
}

// RIFF/WAVE PCM16 reader (kaldi/src/feat/wave-reader.cc:153-320): mono, 16-bit, no scaling.  Two steps so that a batch of
// files is read by several threads straight into the pinned staging area: WavOpen walks the chunk headers and validates
// the format, WavFile::Read copies the samples.
struct WavFile {
  int fd = -1;
  long long data_off = 0;
  int n_samples = 0;
  WavFile() = default;
  WavFile(const WavFile &) = delete;
  WavFile(WavFile &&o) noexcept : fd(o.fd), data_off(o.data_off), n_samples(o.n_samples) { o.fd = -1; }
  WavFile &operator=(WavFile &&o) noexcept {
    if (this != &o) {
      if (fd >= 0) close(fd);
      fd = o.fd;
      data_off = o.data_off;
      n_samples = o.n_samples;
      o.fd = -1;
    }
    return *this;
  }
  ~WavFile() {
    if (fd >= 0) close(fd);
  }
  void Read(const std::string &path, int16_t *dst) const {
    size_t want = sizeof(int16_t) * (size_t)n_samples, got = 0;
    while (got < want) {
      const ssize_t k = pread(fd, reinterpret_cast<char *>(dst) + got, want - got, data_off + (long long)got);
      if (k <= 0) RS_FAIL(path << ": read error");
      got += (size_t)k;
    }
  }
};

static WavFile WavOpen(const std::string &path, float expect_rate) {
  WavFile w;
  w.fd = open(path.c_str(), O_RDONLY);
  if (w.fd < 0) RS_FAIL("cannot open " << path);
  struct stat st;
  if (fstat(w.fd, &st) != 0) RS_FAIL("cannot stat " << path);
  const long long size = st.st_size;
  // the chunk headers of an ordinary file sit in its first few hundred bytes: one read serves the whole parse
  // (a chunk header beyond the buffer -- a long LIST chunk before the samples -- is read on its own)
  unsigned char head[1024];
  const long long have = std::max<long long>(0, pread(w.fd, head, sizeof(head), 0));
  auto peek = [&](long long at, void *dst, int len) -> bool {
    if (at + len <= have) {
      memcpy(dst, head + at, len);
      return true;
    }
    return pread(w.fd, dst, len, at) == len;
  };
  unsigned char hdr[12];
  if (size < 12 || !peek(0, hdr, 12) || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4))
    RS_FAIL(path << ": not a RIFF/WAVE file");
  long long p = 12;
  int channels = 0, bits = 0, fmt = 0;
  uint32_t rate = 0;
  bool have_fmt = false;
  while (p + 8 <= size) {
    unsigned char ch[8];
    if (!peek(p, ch, 8)) RS_FAIL(path << ": read error");
    uint32_t sz;
    memcpy(&sz, ch + 4, 4);
    p += 8;
    if (!memcmp(ch, "fmt ", 4)) {
      unsigned char f[16];
      if (p + 16 > size || !peek(p, f, 16)) RS_FAIL(path << ": truncated fmt chunk");
      uint16_t v16;
      memcpy(&v16, f, 2);
      fmt = v16;
      memcpy(&v16, f + 2, 2);
      channels = v16;
      memcpy(&rate, f + 4, 4);
      memcpy(&v16, f + 14, 2);
      bits = v16;
      have_fmt = true;
    } else if (!memcmp(ch, "data", 4)) {
      if (!have_fmt) RS_FAIL(path << ": data chunk before fmt chunk");
      if (fmt != 1 || bits != 16) RS_FAIL(path << ": only 16-bit PCM WAVE files are supported");
      if (channels != 1) RS_FAIL(path << ": only mono WAVE files are supported (got " << channels << " channels)");
      if ((float)rate != expect_rate)
        RS_FAIL(path << ": sampling frequency mismatch, expected " << expect_rate << ", got " << rate);
      long long n = std::min<long long>(sz, size - p);
      if (sz == 0xffffffffu || sz == 0) n = size - p;  // streamed headers
      if (n / 2 > 0x7fffffff) RS_FAIL(path << ": too long");
      w.data_off = p;
      w.n_samples = (int)(n / 2);
      return w;
    }
    p += (long long)sz + (sz & 1);
  }
  RS_FAIL(path << ": no data chunk");
}

}  // namespace rs

extern "C" {

int rs_decode_pcm(rs_decoder *d_, const int16_t *const *pcm, const int32_t *num_samples, int32_t n, rs_result **out,
                  char *err, size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  if (!d || !out) RS_FAIL("rs_decode_pcm: null argument");
  *out = DecodePcm(d, pcm, num_samples, n);
  return 0;
  API_GUARD_END(1)
}

int rs_decode_wavs(rs_decoder *d_, const char *const *paths, int32_t n, rs_result **out, char *err, size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  if (!d || !out) RS_FAIL("rs_decode_wavs: null argument");
  if (n < 0 || (n && !paths)) RS_FAIL("rs_decode_wavs: bad argument");
  // headers first (sizes fix the batch layout), then the staging threads read the samples into pinned memory
  std::vector<WavFile> files(n);
  std::vector<const int16_t *> ptrs(n, nullptr);
  std::vector<int32_t> ns(n);
  for (int i = 0; i < n; i++)
    if (!paths[i]) RS_FAIL("rs_decode_wavs: path " << i << " is NULL");
  const float rate = d->model->m.mfcc.samp_freq;
  // open + header parse is ~10 us of system calls per file and nothing else can start before the sizes are known:
  // a long list is opened by a few threads (the error of the lowest index is the one reported, as a loop would)
  const int n_open = n >= 64 ? (int)std::min(4u, std::max(1u, std::thread::hardware_concurrency())) : 1;
  std::vector<std::string> open_err(n_open);
  std::vector<int> open_err_at(n_open, n);
  auto open_range = [&](int w) {
    for (int i = w; i < n; i += n_open) {
      try {
        files[i] = WavOpen(paths[i], rate);
        ns[i] = files[i].n_samples;
      } catch (const std::exception &e) {
        open_err[w] = e.what();
        open_err_at[w] = i;
        return;
      }
    }
  };
  {
    std::vector<std::thread> th;
    for (int w = 1; w < n_open; w++) th.emplace_back(open_range, w);
    open_range(0);
    for (auto &t : th) t.join();
  }
  {
    int first = n, who = -1;
    for (int w = 0; w < n_open; w++)
      if (open_err_at[w] < first) first = open_err_at[w], who = w;
    if (who >= 0) RS_FAIL(open_err[who]);
  }
  const FillFn fill = [&](int u, int16_t *dst) { files[u].Read(paths[u], dst); };
  *out = DecodePcm(d, ptrs.data(), ns.data(), n, false, &fill);
  return 0;
  API_GUARD_END(1)
}

int rs_decode_loglikes(rs_decoder *d_, const float *const *loglikes, const int32_t *num_frames, int32_t n, rs_result **out,
                       char *err, size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  if (!d || !out) RS_FAIL("rs_decode_loglikes: null argument");
  *out = DecodeLoglikes(d, loglikes, num_frames, n);
  return 0;
  API_GUARD_END(1)
}

void rs_result_free(rs_result *r) {
  if (!r) return;
  delete[] r->n_hyp;
  delete[] r->word_offset;
  delete[] r->word_ids;
  delete[] r->graph_cost;
  delete[] r->acoustic_cost;
  delete[] r->num_frames;
  delete[] r->status;
  delete[] r->hyp_offset;
  delete[] r->hyp_word_offset;
  delete[] r->hyp_word_ids;
  delete[] r->hyp_graph_cost;
  delete[] r->hyp_acoustic_cost;
  delete r;
}

rs_stream *rs_stream_open(rs_decoder *d_, char *err, size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  if (!d) RS_FAIL("rs_stream_open: null decoder");
  StreamImpl *s = new StreamImpl();
  s->dec = d;
  return reinterpret_cast<rs_stream *>(s);
  API_GUARD_END(nullptr)
}

int rs_stream_accept(rs_stream *s_, const int16_t *pcm, int32_t num_samples, char *err, size_t errlen) {
  API_GUARD_BEGIN
  StreamImpl *s = reinterpret_cast<StreamImpl *>(s_);
  if (!s || num_samples < 0 || (num_samples && !pcm)) RS_FAIL("rs_stream_accept: bad argument");
  s->pcm.insert(s->pcm.end(), pcm, pcm + num_samples);
  return 0;
  API_GUARD_END(1)
}

int rs_streams_finish(rs_stream *const *streams, int32_t n, rs_result **out, char *err, size_t errlen) {
  API_GUARD_BEGIN
  if (n < 0 || !out || (n && !streams)) RS_FAIL("rs_streams_finish: bad argument");
  if (n == 0) {
    *out = NewResult(0);
    return 0;
  }
  for (int i = 0; i < n; i++)
    if (!streams[i]) RS_FAIL("rs_streams_finish: stream " << i << " is NULL (closed?)");
  DecoderImpl *d = reinterpret_cast<StreamImpl *>(streams[0])->dec;
  std::vector<const int16_t *> ptrs(n);
  std::vector<int32_t> ns(n);
  for (int i = 0; i < n; i++) {
    StreamImpl *s = reinterpret_cast<StreamImpl *>(streams[i]);
    if (s->dec != d) RS_FAIL("rs_streams_finish: streams belong to different decoders");
    ptrs[i] = s->pcm.data();
    ns[i] = (int32_t)s->pcm.size();
  }
  *out = DecodePcm(d, ptrs.data(), ns.data(), n, /*online=*/true);
  for (int i = 0; i < n; i++) reinterpret_cast<StreamImpl *>(streams[i])->pcm.clear();
  return 0;
  API_GUARD_END(1)
}

int rs_stream_finish(rs_stream *s, rs_result **out, char *err, size_t errlen) {
  rs_stream *arr[1] = {s};
  return rs_streams_finish(arr, 1, out, err, errlen);
}

void rs_stream_close(rs_stream *s) { delete reinterpret_cast<StreamImpl *>(s); }

int rs_model_check(const char *final_mdl, const char *online_conf, char *out, size_t outlen, char *err, size_t errlen) {
  API_GUARD_BEGIN
  Model m;
  LoadModel(final_mdl, online_conf, &m);
  std::ostringstream os;
  os << "pdfs " << m.trans.num_pdfs << " tids " << m.trans.tid2pdf.size() - 1 << " sf " << m.frame_subsampling_factor
     << " feat_dim " << m.mfcc.num_ceps << " ivector_dim " << (m.has_ivector ? m.ie.ivector_dim : 0) << " num_gauss "
     << (m.has_ivector ? m.ubm.num_gauss : 0) << " priors " << m.log_priors.size() << "\n"
     << DescribePlan(m.plan);
  SetErr(out, outlen, os.str());
  return 0;
  API_GUARD_END(1)
}

int rs_graph_check(const char *hclg_fst, const char *words_txt, int64_t *counts, char *err, size_t errlen) {
  API_GUARD_BEGIN
  Graph g;
  LoadGraph(hclg_fst, words_txt ? words_txt : "", &g);
  if (counts) {
    counts[0] = g.num_states;
    counts[1] = (int64_t)g.e_next.size();
    counts[2] = (int64_t)g.p_next.size();
    counts[3] = g.start;
    int64_t nf = 0;
    for (float f : g.final_cost) nf += std::isfinite(f);
    counts[4] = nf;
    counts[5] = (int64_t)g.words.size();
  }
  return 0;
  API_GUARD_END(1)
}

int rs_debug_gemm(int device, const float *src, int rows, int k, const int *offsets, int n_offsets, int stride,
                  const float *w, int n, const float *bias, int relu, int path, int iters, float *out, float *ms,
                  char *err, size_t errlen) {
  API_GUARD_BEGIN
  if (!src || !w || !out || rows < 1 || k < 1 || n < 1 || n_offsets < 1 || n_offsets > kTcMaxSlabs || stride < 1)
    RS_FAIL("rs_debug_gemm: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) RS_FAIL("no such CUDA device");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  std::vector<void *> owned;
  struct Free {
    std::vector<void *> *v;
    ~Free() {
      for (void *p : *v) cudaFree(p);
    }
  } free_all{&owned};
  const bool tc = path != 0, split_out = path == 2;
  const int ld = RoundUp(k, tc ? 8 : 4), m = std::max(rows / stride, 1), ktot = k * n_offsets;
  const int out_ld = RoundUp(n, split_out ? 8 : 4);
  // source: plain fp32 for the CUDA-core path, two fp16 planes for the tensor-core path
  const void *d_src = nullptr, *d_src_lo = nullptr;
  if (tc) {
    std::vector<__half> hi((size_t)rows * ld, __float2half_rn(0.f)), lo((size_t)rows * ld, __float2half_rn(0.f));
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < k; c++)
        if (!TcSplitHost(src[(size_t)r * k + c], &hi[(size_t)r * ld + c], &lo[(size_t)r * ld + c]))
          RS_FAIL("rs_debug_gemm: source value out of fp16 range");
    d_src = Upload(hi, &owned);
    d_src_lo = Upload(lo, &owned);
  } else {
    std::vector<float> plain((size_t)rows * ld, 0.f);
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < k; c++) plain[(size_t)r * ld + c] = src[(size_t)r * k + c];
    d_src = Upload(plain, &owned);
  }
  std::vector<float> zeros((size_t)m * out_ld, 0.f);
  float *d_out = Upload(zeros, &owned);  // fp32 matrix, or hi | lo fp16 planes (same bytes)
  __half *d_out_lo = split_out ? reinterpret_cast<__half *>(d_out) + (size_t)m * out_ld : nullptr;
  int *d_flag = Upload(std::vector<int>(4, 0), &owned);
  std::vector<float> bias_pad;
  const float *d_bias = nullptr;
  if (bias) {
    bias_pad.assign(bias, bias + n);
    bias_pad.resize(RoundUp(n, 8), 0.f);
    d_bias = Upload(bias_pad, &owned);
  }
  GemmParams g{};
  g.n_slabs = n_offsets;
  for (int s = 0; s < n_offsets; s++) {
    GemmSlab &gs = g.slabs[s];
    gs.src = d_src;
    gs.src_lo = d_src_lo;
    gs.ld = ld;
    gs.rows = rows;
    gs.k = k;
    gs.wcol = s * k;
    gs.num = stride;
    gs.den = 1;
    gs.shift = offsets[s];
  }
  g.out = d_out;
  g.out_lo = d_out_lo;
  g.range_flag = d_flag;
  g.out_ld = out_ld;
  g.m = m;
  g.n = n;
  g.out_step = stride;
  if (d_bias) {
    g.ops[g.n_ops].type = EpiOp::kBias;
    g.ops[g.n_ops].v0 = d_bias;
    g.n_ops++;
  }
  if (relu) g.ops[g.n_ops++].type = EpiOp::kRelu;
  TcParams t{};
  if (tc) {
    if (prop.major != 10) RS_FAIL("the tensor-core path needs sm_100a");
    std::vector<std::pair<int, int>> cols;
    for (int s = 0; s < n_offsets; s++) cols.push_back({s * k, k});
    std::vector<__half> whi, wlo;
    std::vector<int> k0;
    int kp = 0;
    TcPackWeights(w, n, ktot, cols, &whi, &wlo, &k0, &kp);
    const __half *d_whi = Upload(whi, &owned), *d_wlo = Upload(wlo, &owned);
    t.bn = TcTileN(n);
    TcEncodeMap(&t.w_hi, d_whi, n, kp, kp, t.bn);
    TcEncodeMap(&t.w_lo, d_wlo, n, kp, kp, t.bn);
    t.n_slabs = n_offsets;
    for (int s = 0; s < n_offsets; s++) {
      const int sh = offsets[s], rem = ((sh % stride) + stride) % stride;
      const long long view_rows = rows > rem ? (rows - 1 - rem) / stride + 1 : 0;
      TcEncodeMap(&t.a_hi[s], reinterpret_cast<const __half *>(d_src) + (size_t)rem * ld, view_rows, k, (long long)stride * ld, kTcBM);
      TcEncodeMap(&t.a_lo[s], reinterpret_cast<const __half *>(d_src_lo) + (size_t)rem * ld, view_rows, k, (long long)stride * ld, kTcBM);
      t.slabs[s].kblocks = (k + kTcBK - 1) / kTcBK;
      t.slabs[s].wk0 = k0[s];
      t.slabs[s].yshift = (sh - rem) / stride;
    }
    t.out_hi = d_out;
    t.out_lo = d_out_lo;
    t.out_ld = out_ld;
    t.m = m;
    t.n = n;
    t.n_ops = g.n_ops;
    for (int i = 0; i < g.n_ops; i++) t.ops[i] = g.ops[i];
    t.range_flag = d_flag;
    TcConfigure(&t);
  } else {
    std::vector<float> wv(w, w + (size_t)n * ktot);
    g.w = Upload(wv, &owned);
    g.ktot = ktot;
  }
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  auto launch = [&]() {
    if (tc) LaunchGemmTc(t, prop.multiProcessorCount, 0);
    else LaunchGemm(g, 0);
  };
  launch();  // warm-up (and the result)
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  if (iters > 0) {
    CUDA_OK(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; i++) launch();
    CUDA_OK(cudaEventRecord(e1, 0));
    CUDA_OK(cudaEventSynchronize(e1));
    float t_ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&t_ms, e0, e1));
    if (ms) *ms = t_ms / iters;
  } else if (ms) {
    *ms = 0.f;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  std::vector<float> h((size_t)m * out_ld);
  CUDA_OK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  if (split_out) {
    const __half *hh = reinterpret_cast<const __half *>(h.data()), *hl = hh + (size_t)m * out_ld;
    for (int r = 0; r < m; r++)
      for (int c = 0; c < n; c++)
        out[(size_t)r * n + c] = __half2float(hh[(size_t)r * out_ld + c]) + __half2float(hl[(size_t)r * out_ld + c]) / kSplitScale;
  } else {
    for (int r = 0; r < m; r++)
      for (int c = 0; c < n; c++) out[(size_t)r * n + c] = h[(size_t)r * out_ld + c];
  }
  return 0;
  API_GUARD_END(1)
}

int rs_debug_read_matrix(const char *path, float *dst, int32_t *rows, int32_t *cols, char *err, size_t errlen) {
  API_GUARD_BEGIN
  if (!path || !rows || !cols) RS_FAIL("rs_debug_read_matrix: bad argument");
  KaldiReader r(path);
  MatrixF m = r.ReadMatrixF();
  if (dst) {
    if (*rows != m.rows || *cols != m.cols) RS_FAIL("rs_debug_read_matrix: destination is " << *rows << " x " << *cols << ", matrix " << m.rows << " x " << m.cols);
    std::copy(m.d.begin(), m.d.end(), dst);
  }
  *rows = m.rows;
  *cols = m.cols;
  return 0;
  API_GUARD_END(1)
}

int rs_debug_strict_decode(const char *hclg_fst, const int32_t *tid2pdf, int32_t n_tids, const float *loglikes,
                           int32_t n_frames, int32_t num_pdfs, const rs_decoder_opts *opts, int32_t nbest, float acoustic_scale,
                           int32_t *word_offset, int32_t *word_ids, int32_t max_words, float *cost, int32_t *lattice_size,
                           char *err, size_t errlen) {
  API_GUARD_BEGIN
  if (!hclg_fst || !tid2pdf || !loglikes || !opts || !word_offset || !word_ids || !cost || nbest < 1 || n_frames < 0)
    RS_FAIL("rs_debug_strict_decode: bad argument");
  Graph g;
  LoadGraph(hclg_fst, "", &g);
  std::vector<int32_t> epdf(g.e_next.size());
  for (size_t i = 0; i < epdf.size(); i++) {
    const int il = g.e_ilabel[i];
    if (il <= 0 || il >= n_tids) RS_FAIL("HCLG has input label " << il << " outside the transition-id table");
    epdf[i] = tid2pdf[il];
    if (epdf[i] < 0 || epdf[i] >= num_pdfs) RS_FAIL("pdf id out of range");
  }
  StrictOptions so;
  so.beam = opts->beam;
  so.beam_delta = opts->beam_delta;
  so.lattice_beam = opts->lattice_beam;
  so.max_active = opts->max_active;
  so.min_active = opts->min_active;
  so.max_words = std::max(opts->max_words, 1);
  const bool lattice = nbest > 1 || acoustic_scale != 1.0f;
  StrictResult sr;
  StrictDecode(g, epdf.data(), loglikes, num_pdfs, n_frames, so, lattice, &sr);
  if (lattice_size) {
    lattice_size[0] = sr.n_nodes;
    lattice_size[1] = (int32_t)sr.lattice.size();
  }
  std::vector<NbestHyp> hyps;
  if (sr.decoded) {
    if (lattice) {
      LatticeNbest(sr.lattice.data(), (int)sr.lattice.size(), sr.n_nodes, nbest, acoustic_scale, &hyps);
    } else {
      NbestHyp h;
      h.words = sr.words;
      h.graph = sr.graph;
      h.acoustic = sr.acoustic;
      hyps.push_back(h);
    }
  }
  int w = 0;
  for (size_t h = 0; h < hyps.size(); h++) {
    word_offset[h] = w;
    if (w + (int)hyps[h].words.size() > max_words) RS_FAIL("rs_debug_strict_decode: word buffer too small");
    for (int id : hyps[h].words) word_ids[w++] = id;
    cost[2 * h] = hyps[h].graph;
    cost[2 * h + 1] = hyps[h].acoustic;
  }
  word_offset[hyps.size()] = w;
  return (int)hyps.size();
  API_GUARD_END(-1)
}

int rs_decoder_timings(const rs_decoder *d_, rs_timings *t) {
  const DecoderImpl *d = reinterpret_cast<const DecoderImpl *>(d_);
  if (!d || !t) return 1;
  *t = d->last;
  return 0;
}

int rs_debug_fetch(rs_decoder *d_, int32_t what, int32_t utt, float *dst, int32_t *rows, int32_t *cols, char *err,
                   size_t errlen) {
  API_GUARD_BEGIN
  DecoderImpl *d = reinterpret_cast<DecoderImpl *>(d_);
  if (!d) RS_FAIL("rs_debug_fetch: null decoder");
  const auto &B = d->batch;
  if (utt < 0 || utt >= B.n) RS_FAIL("rs_debug_fetch: utterance index out of range");
  const Model &m = d->model->m;
  CUDA_OK(cudaSetDevice(d->model->device));
  const float *src = nullptr;
  int r = 0, c = 0, ld = 0;
  if (what == 5) {
    // pruned state-level lattice of the last n-best call: rows of (src, dst, olabel, graph cost, acoustic cost)
    if ((int)d->lat_hdr.size() != B.n) RS_FAIL("rs_debug_fetch: the last call did not build lattices (rs_decoder_set_nbest)");
    auto sl = d->strict_lat.find(utt);
    if (sl != d->strict_lat.end()) {
      if (rows) *rows = (int)sl->second.size();
      if (cols) *cols = 5;
      for (size_t i = 0; dst && i < sl->second.size(); i++) {
        const LatticeArc &a = sl->second[i];
        float *o = dst + i * 5;
        o[0] = (float)a.src, o[1] = (float)a.dst, o[2] = (float)a.olabel, o[3] = a.graph, o[4] = a.acoustic;
      }
      return 0;
    }
    const LatticeHeader &h = d->lat_hdr[utt];
    if (rows) *rows = h.ok ? h.n_arcs : 0;
    if (cols) *cols = 5;
    if (dst && h.ok) {
      const LatticeArc *a = reinterpret_cast<const LatticeArc *>(d->h_lat.p) + h.arc_begin;
      for (int i = 0; i < h.n_arcs; i++) {
        float *o = dst + (size_t)i * 5;
        o[0] = (float)a[i].src, o[1] = (float)a[i].dst, o[2] = (float)a[i].olabel, o[3] = a[i].graph, o[4] = a[i].acoustic;
      }
    }
    return 0;
  }
  if (what == 6) {
    // a9 / a10: the UBM posteriors of the last call, rows of (gaussian, weight) x num_gselect, unused entries (-1, 0)
    if (B.from_loglikes || !m.has_ivector || !d->d_post_idx.p) RS_FAIL("rs_debug_fetch: the last call computed no posteriors");
    const int S = m.ivec.num_gselect, T = B.num_frames[utt];
    if (rows) *rows = T;
    if (cols) *cols = 2 * S;
    if (dst && T > 0) {
      std::vector<int> idx((size_t)T * S);
      std::vector<float> w((size_t)T * S);
      CUDA_OK(cudaMemcpy(idx.data(), d->d_post_idx.as<int>() + (size_t)B.frame_offset[utt] * S, idx.size() * sizeof(int), cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(w.data(), d->d_post_w.as<float>() + (size_t)B.frame_offset[utt] * S, w.size() * sizeof(float), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < idx.size(); i++) {
        dst[2 * i] = (float)idx[i];
        dst[2 * i + 1] = idx[i] >= 0 ? w[i] : 0.f;
      }
    }
    return 0;
  }
  if (what == 2) {
    r = B.n_out[utt];
    c = m.trans.num_pdfs;
    ld = B.ll_ld;
    src = B.loglikes + (size_t)B.ll_row0[utt] * ld;
  } else {
    if (B.from_loglikes) RS_FAIL("rs_debug_fetch: the last call did not run the feature stages");
    if (what == 0 || what == 3) {
      r = B.num_frames[utt];
      c = ld = m.mfcc.num_ceps;
      const DevBuf &b = what == 0 ? d->d_mfcc : d->d_mfcc_norm;
      if (!b.p) RS_FAIL("rs_debug_fetch: buffer not available");
      src = b.as<float>() + (size_t)B.frame_offset[utt] * ld;
    } else if (what == 1) {
      if (!m.has_ivector) RS_FAIL("rs_debug_fetch: model has no iVector input");
      r = B.v_begin[utt + 1] - B.v_begin[utt];  // one row per iVector solve (1 when decoding offline)
      c = m.ie.ivector_dim;
      ld = RoundUp(c, 4);
      src = d->slots[m.plan.buffers[m.plan.ivector_buffer].slot].as<float>() + (size_t)B.v_begin[utt] * ld;
    } else if (what == 4) {
      if (!m.has_ivector) RS_FAIL("rs_debug_fetch: model has no iVector input");
      r = B.num_frames[utt];
      c = ld = m.lda.rows;
      src = d->d_xnorm.as<float>() + (size_t)B.frame_offset[utt] * ld;
    } else {
      RS_FAIL("rs_debug_fetch: unknown item " << what);
    }
  }
  if (rows) *rows = r;
  if (cols) *cols = c;
  if (dst && r > 0)
    CUDA_OK(cudaMemcpy2D(dst, sizeof(float) * c, src, sizeof(float) * ld, sizeof(float) * c, r, cudaMemcpyDeviceToHost));
  return 0;
  API_GUARD_END(1)
}

}  // extern "C"
