// Host-side model artefacts of the hot path, parsed from the files rhasspy-speech trains
// (reference rhasspy_speech/transcribe_wav.py:43-57 for the layout):
//   online.conf -> mfcc.conf, ivector_extractor.conf -> splice.conf, online_cmvn.conf,
//   final.mat / final.dubm / final.ie / global_cmvn.stats ; final.mdl ; HCLG.fst ; words.txt
// Each parser cites the reference reader it is compatible with.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "kaldi_io.h"

namespace rs {

// --- option structs (defaults = the reference's) -------------------------------------------
// kaldi/src/feat/feature-window.h:53-105, mel-computations.h:56-74, feature-mfcc.h:51-81
struct MfccOptions {
  float samp_freq = 16000.f, frame_shift_ms = 10.f, frame_length_ms = 25.f;
  float dither = 1.0f, preemph_coeff = 0.97f, blackman_coeff = 0.42f;
  bool remove_dc_offset = true, round_to_power_of_two = true, snip_edges = true;
  std::string window_type = "povey";
  int num_bins = 23;
  float low_freq = 20.f, high_freq = 0.f;
  int num_ceps = 13;
  bool use_energy = true, raw_energy = true, htk_compat = false;
  float energy_floor = 0.f, cepstral_lifter = 22.f;
  int WindowShift() const { return (int)(samp_freq * 0.001f * frame_shift_ms); }
  int WindowSize() const { return (int)(samp_freq * 0.001f * frame_length_ms); }
  int PaddedWindowSize() const {
    int n = WindowSize();
    if (!round_to_power_of_two) return n;
    int p = 1;
    while (p < n) p <<= 1;
    return p;
  }
};

// kaldi/src/feat/online-feature.h:219-227
struct CmvnOptions {
  int cmn_window = 600, speaker_frames = 600, global_frames = 200, modulus = 20;
  bool normalize_mean = true, normalize_variance = false;
};

// kaldi/src/online2/online-ivector-feature.h:105-163
struct IvectorOptions {
  int splice_left = 4, splice_right = 4;
  CmvnOptions cmvn;
  bool online_cmvn_iextractor = false;
  int ivector_period = 10, num_gselect = 5, num_cg_iters = 15;
  float min_post = 0.025f, posterior_scale = 0.1f, max_count = 0.f;
  float max_remembered_frames = 1000.f;
};

// --- data ------------------------------------------------------------------------------------
struct DiagGmm {  // kaldi/src/gmm/diag-gmm.cc:729-755 (+ ComputeGconsts :94-124)
  int num_gauss = 0, dim = 0;
  std::vector<float> gconsts, weights;
  MatrixF means_invvars, inv_vars;  // [G x D]
};

struct IvectorExtractor {  // kaldi/src/ivector/ivector-extractor.cc:828-847, 182-218
  int num_gauss = 0, feat_dim = 0, ivector_dim = 0;
  double prior_offset = 0.0;
  std::vector<double> sigma_inv_m;  // [G][D][R]  Sigma_g^{-1} M_g
  std::vector<double> u;            // [G][R(R+1)/2]  packed lower triangle of M_g^T Sigma_g^{-1} M_g
};

struct TransitionModel {  // kaldi/src/hmm/transition-model.cc:394-420, 144-188
  std::vector<int32_t> tid2pdf;  // index 0 unused (transition-ids are 1-based)
  int num_pdfs = 0;
};

// --- nnet3 -----------------------------------------------------------------------------------
struct Component {
  std::string type, name;
  int in_dim = 0, out_dim = 0;
  // affine-like (FixedAffine / Affine / NaturalGradientAffine / Linear / Tdnn)
  MatrixF linear;                  // [out x (in * n_offsets)]
  std::vector<float> bias;         // may be empty
  std::vector<int32_t> time_offsets;  // Tdnn only; {0} otherwise
  // BatchNorm (test mode): y = x*scale + offset ; also FixedScale/FixedBias/PerElement*
  std::vector<float> scale, offset;
  int block_dim = 0;
  float dropout_scale = 1.f;       // identity in test mode
};

struct DescTerm {
  int node = -1;       // index into Nnet3::nodes
  float scale = 1.f;
  int t_offset = 0;
  bool const_time = false;  // ReplaceIndex(x, t, 0): one row per utterance
};
struct DescPart {  // one column block of an Append(...); the terms are summed
  std::vector<DescTerm> terms;
  int dim = 0;
};
typedef std::vector<DescPart> Descriptor;

struct Node {
  enum Kind { kInput, kComponent, kOutput, kDimRange } kind = kInput;
  std::string name;
  int dim = 0;
  int component = -1;
  Descriptor input;
  int dim_offset = 0;  // kDimRange
};

struct Nnet3 {
  std::vector<Node> nodes;
  std::vector<Component> components;
  int left_context = 0, right_context = 0;
  std::vector<float> priors;
  int FindNode(const std::string &name) const;
};

// --- the compiled forward plan --------------------------------------------------------------
// Every buffer lives on one global time axis: row i of a buffer with step s is time i*s.
// Utterance u owns times [origin_u, origin_u + T_u); between utterances there is a gap of at
// least (left+right) model context, so a layer is one dense operation over the whole axis.
struct EpiOp {
  enum Type { kBias, kRelu, kScaleOffset, kScale, kAddScaled, kUttBias } type;
  int vec0 = -1, vec1 = -1;  // indices into Plan::vectors
  float alpha = 1.f;
  int buffer = -1;           // kAddScaled: other buffer (same step); kUttBias: utt-bias matrix id
};
struct Slab {
  int src = -1;      // buffer
  int t_offset = 0;  // input time = output time + t_offset
  int k = 0;         // columns taken from the source (all of it)
  int wcol = 0;      // first column in the step's weight matrix
};
struct Step {
  enum Type { kGemm, kElementwise, kLogSoftmax, kUttGemm } type = kGemm;
  int out = -1;               // buffer written
  int n = 0;                  // output columns
  int weight = -1;            // index into Plan::matrices ([n x ktot] row-major)
  int ktot = 0;
  std::vector<Slab> slabs;    // kGemm: inputs ; kElementwise: summed terms (alpha in Slab::wcol unused)
  std::vector<float> term_scale;  // kElementwise: per-slab scale
  int col_offset = 0;         // kElementwise: first output column (Append of parts)
  std::vector<EpiOp> ops;
  std::string name;
};
struct PlanBuffer {
  int dim = 0, step = 1;
  bool per_utt = false;  // one row per utterance (ivector, utt-bias)
  bool is_input = false, is_ivector = false, is_output = false;
  int slot = -1;         // physical allocation after liveness analysis
  std::string name;
};
struct Plan {
  std::vector<PlanBuffer> buffers;
  std::vector<Step> steps;
  std::vector<MatrixF> matrices;
  std::vector<std::vector<float>> vectors;
  int input_buffer = -1, ivector_buffer = -1, output_buffer = -1;
  int left_context = 0, right_context = 0;  // in input frames
  int align = 1;                            // lcm of all buffer steps
  int frame_subsampling_factor = 1;
  int num_slots = 0;
  std::vector<int> slot_dim_step;           // unused placeholder for diagnostics
  double flops_per_axis_row(int step) const;
};

struct Model {
  MfccOptions mfcc;
  bool has_ivector = false;
  IvectorOptions ivec;
  MatrixF lda;  // final.mat [D_out x (D*(L+1+R)) (+1)]
  MatrixD global_cmvn;  // [2 x (D+1)]
  DiagGmm ubm;
  IvectorExtractor ie;
  bool nnet_cmvn = false;  // --cmvn-config given in online.conf: CMVN on the nnet input too
  CmvnOptions nnet_cmvn_opts;
  MatrixD nnet_global_cmvn;
  int frame_subsampling_factor = 1;
  int frames_per_chunk = 24;  // online decoding: nnet chunk (and iVector period) before rounding to sf
  float acoustic_scale = 1.0f;
  TransitionModel trans;
  Nnet3 nnet;
  Plan plan;
  std::vector<float> log_priors;
};

// kaldi/src/util/parse-options.cc ReadConfigFile: one --name=value per line, '#' comments
std::map<std::string, std::string> ReadConfigFile(const std::string &path);
void LoadModel(const std::string &final_mdl, const std::string &online_conf, Model *m);
void CompilePlan(const Nnet3 &nnet, int frame_subsampling_factor, Plan *plan);
std::string DescribePlan(const Plan &plan);

// --- graph -----------------------------------------------------------------------------------
struct Graph {  // ConstFst / VectorFst <StdArc>; kaldi/openfst/src/include/fst/const-fst.h:192-232
  int64_t start = -1;
  int32_t num_states = 0;
  // emitting arcs (ilabel != 0) and epsilon-input arcs, each CSR by source state, original order
  std::vector<uint32_t> e_begin, p_begin;  // [S+1]
  std::vector<int32_t> e_ilabel, e_olabel, e_next, e_src;
  std::vector<float> e_weight;
  std::vector<int32_t> p_olabel, p_next, p_src;
  std::vector<float> p_weight;
  std::vector<float> final_cost;  // +inf = not final
  std::vector<std::string> words;  // id -> symbol (words.txt)
};
void LoadGraph(const std::string &hclg_fst, const std::string &words_txt, Graph *g);

}  // namespace rs
