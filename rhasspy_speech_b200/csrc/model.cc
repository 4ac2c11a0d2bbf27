// Loaders for the artefacts on the hot path.  See model.h for the structures.
#include "model.h"

#include <algorithm>
#include <cmath>
#include <functional>
#include <numeric>
#include <set>

namespace rs {

// ---------------------------------------------------------------------------------------------
// conf files: kaldi/src/util/parse-options.cc (ReadConfigFile): "--name=value", '#' comments,
// names are normalised ('_' -> '-').

static std::string Trim(const std::string &s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

std::map<std::string, std::string> ReadConfigFile(const std::string &path) {
  std::ifstream f(path);
  if (!f) RS_FAIL("cannot open config file " << path);
  std::map<std::string, std::string> kv;
  std::string line;
  while (std::getline(f, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    line = Trim(line);
    if (line.empty()) continue;
    if (line.size() < 3 || line[0] != '-' || line[1] != '-')
      RS_FAIL(path << ": bad config line '" << line << "'");
    std::string key, val = "true";
    size_t eq = line.find('=');
    if (eq == std::string::npos) {
      key = line.substr(2);
    } else {
      key = line.substr(2, eq - 2);
      val = Trim(line.substr(eq + 1));
    }
    for (char &c : key)
      if (c == '_') c = '-';
    kv[key] = val;
  }
  return kv;
}

static bool ToBool(const std::string &v) {
  if (v == "true" || v == "t" || v == "T" || v == "1" || v == "") return true;
  if (v == "false" || v == "f" || v == "F" || v == "0") return false;
  RS_FAIL("bad boolean option value '" << v << "'");
}

template <typename T>
static void Take(std::map<std::string, std::string> &kv, const char *name, T *dst);
template <>
void Take<float>(std::map<std::string, std::string> &kv, const char *name, float *dst) {
  auto it = kv.find(name);
  if (it == kv.end()) return;
  *dst = (float)atof(it->second.c_str());
  kv.erase(it);
}
template <>
void Take<int>(std::map<std::string, std::string> &kv, const char *name, int *dst) {
  auto it = kv.find(name);
  if (it == kv.end()) return;
  *dst = atoi(it->second.c_str());
  kv.erase(it);
}
template <>
void Take<bool>(std::map<std::string, std::string> &kv, const char *name, bool *dst) {
  auto it = kv.find(name);
  if (it == kv.end()) return;
  *dst = ToBool(it->second);
  kv.erase(it);
}
template <>
void Take<std::string>(std::map<std::string, std::string> &kv, const char *name, std::string *dst) {
  auto it = kv.find(name);
  if (it == kv.end()) return;
  *dst = it->second;
  kv.erase(it);
}

static void ParseMfccConf(const std::string &path, MfccOptions *o) {
  auto kv = ReadConfigFile(path);
  Take(kv, "sample-frequency", &o->samp_freq);
  Take(kv, "frame-shift", &o->frame_shift_ms);
  Take(kv, "frame-length", &o->frame_length_ms);
  Take(kv, "dither", &o->dither);
  Take(kv, "preemphasis-coefficient", &o->preemph_coeff);
  Take(kv, "remove-dc-offset", &o->remove_dc_offset);
  Take(kv, "window-type", &o->window_type);
  Take(kv, "round-to-power-of-two", &o->round_to_power_of_two);
  Take(kv, "blackman-coeff", &o->blackman_coeff);
  Take(kv, "snip-edges", &o->snip_edges);
  Take(kv, "num-mel-bins", &o->num_bins);
  Take(kv, "low-freq", &o->low_freq);
  Take(kv, "high-freq", &o->high_freq);
  Take(kv, "num-ceps", &o->num_ceps);
  Take(kv, "use-energy", &o->use_energy);
  Take(kv, "energy-floor", &o->energy_floor);
  Take(kv, "raw-energy", &o->raw_energy);
  Take(kv, "cepstral-lifter", &o->cepstral_lifter);
  Take(kv, "htk-compat", &o->htk_compat);
  bool allow_downsample = false, allow_upsample = false;
  Take(kv, "allow-downsample", &allow_downsample);
  Take(kv, "allow-upsample", &allow_upsample);
  float vtln_low = 100, vtln_high = -500;
  Take(kv, "vtln-low", &vtln_low);
  Take(kv, "vtln-high", &vtln_high);
  bool debug_mel = false;
  Take(kv, "debug-mel", &debug_mel);
  if (!kv.empty()) RS_FAIL(path << ": unsupported MFCC option --" << kv.begin()->first);
  if (!o->snip_edges) RS_FAIL(path << ": --snip-edges=false is not supported");
  if (o->htk_compat) RS_FAIL(path << ": --htk-compat=true is not supported");
  if (o->num_ceps > o->num_bins) RS_FAIL(path << ": num-ceps cannot be larger than num-mel-bins");
  if (o->PaddedWindowSize() & (o->PaddedWindowSize() - 1))
    RS_FAIL(path << ": the window must be padded to a power of two (--round-to-power-of-two=true)");
}

static void ParseCmvnConf(const std::string &path, CmvnOptions *o) {
  auto kv = ReadConfigFile(path);
  Take(kv, "cmn-window", &o->cmn_window);
  Take(kv, "global-frames", &o->global_frames);
  Take(kv, "speaker-frames", &o->speaker_frames);
  Take(kv, "norm-vars", &o->normalize_variance);
  Take(kv, "norm-means", &o->normalize_mean);
  std::string skip;
  Take(kv, "skip-dims", &skip);
  if (!skip.empty()) RS_FAIL(path << ": --skip-dims is not supported");
  if (!kv.empty()) RS_FAIL(path << ": unsupported CMVN option --" << kv.begin()->first);
  if (!(o->speaker_frames <= o->cmn_window && o->global_frames <= o->speaker_frames))
    RS_FAIL(path << ": inconsistent CMVN frame counts");
}

// ---------------------------------------------------------------------------------------------

static void ReadDiagGmm(const std::string &path, DiagGmm *g) {
  KaldiReader r(path);
  std::string t = r.ReadToken();
  if (t != "<DiagGMMBegin>" && t != "<DiagGMM>") RS_FAIL(path << ": expected <DiagGMM>, got " << t);
  t = r.ReadToken();
  if (t == "<GCONSTS>") {
    r.ReadVectorF();
    r.ExpectToken("<WEIGHTS>");
  } else if (t != "<WEIGHTS>") {
    RS_FAIL(path << ": expected <WEIGHTS> or <GCONSTS>, got " << t);
  }
  g->weights = r.ReadVectorF();
  r.ExpectToken("<MEANS_INVVARS>");
  g->means_invvars = r.ReadMatrixF();
  r.ExpectToken("<INV_VARS>");
  g->inv_vars = r.ReadMatrixF();
  g->num_gauss = g->inv_vars.rows;
  g->dim = g->inv_vars.cols;
  if ((int)g->weights.size() != g->num_gauss || g->means_invvars.rows != g->num_gauss || g->means_invvars.cols != g->dim)
    RS_FAIL(path << ": inconsistent DiagGmm dimensions");
  // ComputeGconsts (diag-gmm.cc:94-124): float accumulator, double increments
  g->gconsts.resize(g->num_gauss);
  float offset = (float)(-0.5 * 1.8378770664093454835606594728112 * g->dim);
  for (int m = 0; m < g->num_gauss; m++) {
    float gc = logf(g->weights[m]) + offset;
    for (int d = 0; d < g->dim; d++) {
      float iv = g->inv_vars(m, d), mi = g->means_invvars(m, d);
      gc += 0.5 * logf(iv) - 0.5 * mi * mi / iv;
    }
    if (std::isnan(gc)) RS_FAIL(path << ": NaN gconst");
    if (std::isinf(gc)) gc = gc > 0 ? -gc : gc;
    g->gconsts[m] = gc;
  }
}

static void ReadIvectorExtractor(const std::string &path, IvectorExtractor *ie) {
  KaldiReader r(path);
  r.ExpectToken("<IvectorExtractor>");
  r.ExpectToken("<w>");
  MatrixD w = r.ReadMatrixD();
  if (w.rows != 0) RS_FAIL(path << ": iVector extractors with iVector-dependent weights are not supported");
  r.ExpectToken("<w_vec>");
  r.ReadVectorD();
  r.ExpectToken("<M>");
  int G = r.ReadInt32();
  if (G <= 0) RS_FAIL(path << ": bad Gaussian count");
  std::vector<MatrixD> M(G);
  for (int i = 0; i < G; i++) M[i] = r.ReadMatrixD();
  r.ExpectToken("<SigmaInv>");
  int D = M[0].rows, R = M[0].cols;
  ie->num_gauss = G;
  ie->feat_dim = D;
  ie->ivector_dim = R;
  int P = R * (R + 1) / 2;
  ie->sigma_inv_m.assign((size_t)G * D * R, 0.0);
  ie->u.assign((size_t)G * P, 0.0);
  std::vector<double> S((size_t)D * D);
  for (int g = 0; g < G; g++) {
    int dim = 0;
    std::vector<double> sp = r.ReadSpMatrixD(&dim);
    if (dim != D || M[g].rows != D || M[g].cols != R) RS_FAIL(path << ": inconsistent extractor dimensions");
    for (int i = 0, k = 0; i < D; i++)
      for (int j = 0; j <= i; j++, k++) S[(size_t)i * D + j] = S[(size_t)j * D + i] = sp[k];
    double *sim = &ie->sigma_inv_m[(size_t)g * D * R];
    for (int i = 0; i < D; i++)
      for (int j = 0; j < D; j++) {
        double s = S[(size_t)i * D + j];
        if (s == 0.0) continue;
        const double *mrow = &M[g].d[(size_t)j * R];
        double *o = sim + (size_t)i * R;
        for (int c = 0; c < R; c++) o[c] += s * mrow[c];
      }
    // U_g = M^T (Sigma^-1 M), packed lower triangle
    double *u = &ie->u[(size_t)g * P];
    for (int a = 0, k = 0; a < R; a++)
      for (int b = 0; b <= a; b++, k++) {
        double acc = 0.0;
        for (int i = 0; i < D; i++) acc += M[g].d[(size_t)i * R + a] * sim[(size_t)i * R + b];
        u[k] = acc;
      }
  }
  r.ExpectToken("<IvectorOffset>");
  ie->prior_offset = r.ReadDouble();
  r.ExpectToken("</IvectorExtractor>");
}

// ---------------------------------------------------------------------------------------------
// TransitionModel (hmm/transition-model.cc:394-420 Read, :144-188 ComputeDerived;
// hmm/hmm-topology.cc:39-160 Read)

struct TopoState {
  int fwd = -1, slf = -1;
  std::vector<std::pair<int, float>> trans;
};

static void ReadTransitionModel(KaldiReader &r, TransitionModel *tm) {
  r.ExpectToken("<TransitionModel>");
  r.ExpectToken("<Topology>");
  std::vector<std::vector<TopoState>> entries;
  std::vector<int32_t> phone2idx;
  if (!r.binary()) {
    while (true) {
      std::string t = r.ReadToken();
      if (t == "</Topology>") break;
      if (t != "<TopologyEntry>") RS_FAIL(r.path() << ": expected <TopologyEntry>, got " << t);
      r.ExpectToken("<ForPhones>");
      std::vector<int> phones;
      while (true) {
        std::string s = r.ReadToken();
        if (s == "</ForPhones>") break;
        phones.push_back(atoi(s.c_str()));
      }
      std::vector<TopoState> entry;
      t = r.ReadToken();
      while (t != "</TopologyEntry>") {
        if (t != "<State>") RS_FAIL(r.path() << ": expected <State>, got " << t);
        int idx = r.ReadInt32();
        if (idx != (int)entry.size()) RS_FAIL(r.path() << ": topology states out of order");
        TopoState st;
        t = r.ReadToken();
        if (t == "<PdfClass>") {
          st.fwd = st.slf = r.ReadInt32();
          t = r.ReadToken();
        } else if (t == "<ForwardPdfClass>") {
          st.fwd = r.ReadInt32();
          r.ExpectToken("<SelfLoopPdfClass>");
          st.slf = r.ReadInt32();
          t = r.ReadToken();
        }
        while (t == "<Transition>") {
          int dst = r.ReadInt32();
          float p = r.ReadFloat();
          st.trans.push_back({dst, p});
          t = r.ReadToken();
        }
        if (t != "</State>") RS_FAIL(r.path() << ": expected </State>, got " << t);
        entry.push_back(st);
        t = r.ReadToken();
      }
      int my = (int)entries.size();
      entries.push_back(entry);
      for (int p : phones) {
        if ((int)phone2idx.size() <= p) phone2idx.resize(p + 1, -1);
        phone2idx[p] = my;
      }
    }
  } else {
    r.ReadIntVector();  // phones
    phone2idx = r.ReadIntVector();
    int sz = r.ReadInt32();
    bool is_hmm = true;
    if (sz == -1) {
      is_hmm = false;
      sz = r.ReadInt32();
    }
    entries.resize(sz);
    for (int i = 0; i < sz; i++) {
      int n = r.ReadInt32();
      entries[i].resize(n);
      for (int j = 0; j < n; j++) {
        entries[i][j].fwd = r.ReadInt32();
        entries[i][j].slf = is_hmm ? entries[i][j].fwd : r.ReadInt32();
        int nt = r.ReadInt32();
        entries[i][j].trans.resize(nt);
        for (int k = 0; k < nt; k++) {
          entries[i][j].trans[k].first = r.ReadInt32();
          entries[i][j].trans[k].second = r.ReadFloat();
        }
      }
    }
    r.ExpectToken("</Topology>");
  }
  std::string tok = r.ReadToken();
  if (tok != "<Tuples>" && tok != "<Triples>") RS_FAIL(r.path() << ": expected <Tuples>/<Triples>, got " << tok);
  int n = r.ReadInt32();
  tm->tid2pdf.assign(1, 0);
  tm->num_pdfs = 0;
  for (int i = 0; i < n; i++) {
    int phone = r.ReadInt32(), hs = r.ReadInt32(), fwd = r.ReadInt32();
    int slf = tok == "<Tuples>" ? r.ReadInt32() : fwd;
    if (phone < 0 || phone >= (int)phone2idx.size() || phone2idx[phone] < 0)
      RS_FAIL(r.path() << ": phone " << phone << " has no topology");
    const auto &entry = entries[phone2idx[phone]];
    if (hs < 0 || hs >= (int)entry.size()) RS_FAIL(r.path() << ": bad hmm-state in transition model");
    for (const auto &tr : entry[hs].trans) tm->tid2pdf.push_back(tr.first == hs ? slf : fwd);
    tm->num_pdfs = std::max(tm->num_pdfs, std::max(fwd, slf) + 1);
  }
  tok = r.ReadToken();
  if (tok != "</Tuples>" && tok != "</Triples>") RS_FAIL(r.path() << ": expected </Tuples>, got " << tok);
  r.ExpectToken("<LogProbs>");
  r.ReadVectorF();
  r.ExpectToken("</LogProbs>");
  r.ExpectToken("</TransitionModel>");
}

// ---------------------------------------------------------------------------------------------
// nnet3 components.  Field types per component follow the reference's Read() functions
// (nnet-simple-component.cc, nnet-tdnn-component.cc:410-455, nnet-normalize-component.cc:591-614,
// nnet-general-component.cc:1636-1671, nnet-component-itf.cc:263-304,481-540).

enum FieldType { kF, kFF, kI, kII, kB, kD, kVec, kMat, kIVec, kFlag };

static void SkipField(KaldiReader &r, FieldType t) {
  switch (t) {
    case kF: r.ReadFloat(); break;
    case kFF: r.ReadFloat(); r.ReadFloat(); break;
    case kI: r.ReadInt32(); break;
    case kII: r.ReadInt32(); r.ReadInt32(); break;
    case kB: r.ReadBool(); break;
    case kD: r.ReadDouble(); break;
    case kVec: r.ReadVectorD(); break;
    case kMat: r.ReadMatrixD(); break;
    case kIVec: r.ReadIntVector(); break;
    case kFlag: break;
  }
}

static void ReadComponent(KaldiReader &r, Component *c) {
  std::string open = r.ReadToken();
  if (open.size() < 3 || open[0] != '<' || open.back() != '>') RS_FAIL(r.path() << ": bad component tag " << open);
  c->type = open.substr(1, open.size() - 2);
  const std::string close = "</" + c->type + ">";
  std::map<std::string, FieldType> f = {
      {"<LearningRateFactor>", kF}, {"<IsGradient>", kB}, {"<MaxChange>", kF}, {"<L2Regularize>", kF},
      {"<LearningRate>", kF}, {"<OrthonormalConstraint>", kF}, {"<UseNaturalGradient>", kB},
      {"<NumSamplesHistory>", kF}, {"<AlphaInOut>", kFF}, {"<Alpha>", kF}, {"<RankInOut>", kII},
      {"<RankIn>", kI}, {"<RankOut>", kI}, {"<UpdatePeriod>", kI}, {"<MaxChangePerSample>", kF},
      {"<UpdateCount>", kD}, {"<ActiveScalingCount>", kD}, {"<MaxChangeScaleStats>", kD}, {"<Rank>", kI}};
  bool nonlinear = false;
  const std::string &T = c->type;
  if (T == "RectifiedLinearComponent" || T == "LogSoftmaxComponent" || T == "SigmoidComponent" ||
      T == "TanhComponent" || T == "SoftmaxComponent") {
    nonlinear = true;
    f = {{"<Dim>", kI}, {"<BlockDim>", kI}, {"<ValueAvg>", kVec}, {"<DerivAvg>", kVec}, {"<Count>", kD},
         {"<OderivRms>", kVec}, {"<OderivCount>", kD}, {"<NumDimsSelfRepaired>", kD}, {"<NumDimsProcessed>", kD},
         {"<SelfRepairLowerThreshold>", kF}, {"<SelfRepairUpperThreshold>", kF}, {"<SelfRepairScale>", kF}};
  } else if (T == "BatchNormComponent") {
    f = {{"<Dim>", kI}, {"<BlockDim>", kI}, {"<Epsilon>", kF}, {"<TargetRms>", kF}, {"<TestMode>", kB},
         {"<Count>", kD}, {"<StatsMean>", kVec}, {"<StatsVar>", kVec}};
  } else if (T == "GeneralDropoutComponent") {
    f = {{"<Dim>", kI}, {"<BlockDim>", kI}, {"<TimePeriod>", kI}, {"<DropoutProportion>", kF},
         {"<SpecAugmentMaxProportion>", kF}, {"<SpecAugmentMaxRegions>", kI}, {"<TestMode>", kFlag}, {"<Continuous>", kFlag}};
  } else if (T == "DropoutComponent") {
    f = {{"<Dim>", kI}, {"<DropoutProportion>", kF}, {"<DropoutPerFrame>", kB}, {"<TestMode>", kB}};
  } else if (T == "NoOpComponent") {
    // old format stored NonlinearComponent stats with float counts (nnet-simple-component.cc:493-527)
    f = {{"<Dim>", kI}, {"<BackpropScale>", kF}, {"<ValueAvg>", kVec}, {"<DerivAvg>", kVec}, {"<Count>", kF},
         {"<OderivRms>", kVec}, {"<OderivCount>", kF}, {"<NumDimsSelfRepaired>", kF}, {"<NumDimsProcessed>", kF}};
  } else if (T == "FixedAffineComponent" || T == "AffineComponent" || T == "NaturalGradientAffineComponent" ||
             T == "LinearComponent" || T == "TdnnComponent") {
    f["<LinearParams>"] = kMat;
    f["<Params>"] = kMat;
    f["<BiasParams>"] = kVec;
    f["<TimeOffsets>"] = kIVec;
  } else if (T == "FixedScaleComponent") {
    f = {{"<Scales>", kVec}};
  } else if (T == "FixedBiasComponent") {
    f = {{"<Bias>", kVec}};
  } else if (T == "PerElementScaleComponent" || T == "NaturalGradientPerElementScaleComponent") {
    f["<Params>"] = kVec;
  } else if (T == "PerElementOffsetComponent") {
    f["<Offsets>"] = kVec;
    f["<Dim>"] = kI;
  } else if (T == "ScaleAndOffsetComponent") {
    f["<Dim>"] = kI;
    f["<Scales>"] = kVec;
    f["<Offsets>"] = kVec;
  } else {
    RS_FAIL(r.path() << ": nnet3 component type " << T << " is not supported by this decoder");
  }
  float epsilon = 1e-3f, target_rms = 1.f;
  double count = 0.0;
  std::vector<float> stats_mean, stats_var;
  int dim = -1;
  c->time_offsets = {0};
  while (true) {
    std::string t = r.ReadToken();
    if (t[0] != '<') t = "<" + t;
    if (t == close) break;
    if (t == open) continue;
    // NaturalGradientAffineComponent tolerates either tag at the end (nnet-simple-component.cc:2844-2849)
    auto it = f.find(t);
    if (it == f.end()) RS_FAIL(r.path() << ": unexpected field " << t << " in " << T);
    if (t == "<LinearParams>" || (t == "<Params>" && it->second == kMat)) {
      c->linear = r.ReadMatrixF();
    } else if (t == "<BiasParams>") {
      c->bias = r.ReadVectorF();
    } else if (t == "<TimeOffsets>") {
      c->time_offsets = r.ReadIntVector();
    } else if (t == "<Dim>") {
      dim = r.ReadInt32();
    } else if (t == "<BlockDim>") {
      c->block_dim = r.ReadInt32();
    } else if (t == "<Epsilon>") {
      epsilon = r.ReadFloat();
    } else if (t == "<TargetRms>") {
      target_rms = r.ReadFloat();
    } else if (t == "<Count>" && T == "BatchNormComponent") {
      count = r.ReadDouble();
    } else if (t == "<StatsMean>") {
      stats_mean = r.ReadVectorF();
    } else if (t == "<StatsVar>") {
      stats_var = r.ReadVectorF();
    } else if (t == "<Scales>" || (t == "<Params>" && it->second == kVec)) {
      c->scale = r.ReadVectorF();
    } else if (t == "<Offsets>" || t == "<Bias>") {
      c->offset = r.ReadVectorF();
    } else {
      SkipField(r, it->second);
    }
  }
  (void)nonlinear;
  if (!c->linear.d.empty() || T == "LinearComponent" || T == "TdnnComponent" || T.find("Affine") != std::string::npos) {
    if (c->linear.rows <= 0) RS_FAIL(r.path() << ": component of type " << T << " has no parameters");
    c->out_dim = c->linear.rows;
    int noff = (int)c->time_offsets.size();
    if (noff < 1 || c->linear.cols % noff) RS_FAIL(r.path() << ": TdnnComponent parameter/offset mismatch");
    c->in_dim = c->linear.cols / noff;
    if (!c->bias.empty() && (int)c->bias.size() != c->out_dim) RS_FAIL(r.path() << ": bias dimension mismatch");
  } else if (T == "BatchNormComponent") {
    // ComputeDerived (nnet-normalize-component.cc:209-246), test mode forced by the reference binaries
    // (online2-wav-nnet3-latgen-faster.cc:169); float arithmetic as in CuVector.
    if (c->block_dim <= 0) c->block_dim = dim;
    if (count == 0.0) RS_FAIL(r.path() << ": BatchNormComponent without statistics");
    int bd = c->block_dim;
    if ((int)stats_mean.size() != bd || (int)stats_var.size() != bd || dim % bd) RS_FAIL(r.path() << ": bad BatchNorm dims");
    c->in_dim = c->out_dim = dim;
    c->scale.resize(dim);
    c->offset.resize(dim);
    for (int i = 0; i < bd; i++) {
      // Read(): sumsq = (var + mean*mean) * count ; sum = mean * count
      float sum = stats_mean[i], sumsq = stats_var[i] + stats_mean[i] * stats_mean[i];
      sum = sum * (float)count;
      sumsq = sumsq * (float)count;
      float off = sum * (float)(-1.0 / count);
      float sc = sumsq * (float)(1.0 / count);
      sc = sc + (-1.0f) * off * off;
      if (sc < 0.f) sc = 0.f;
      sc += epsilon;
      sc = powf(sc, -0.5f);
      sc *= target_rms;
      off *= sc;
      for (int b = 0; b < dim / bd; b++) {
        c->scale[b * bd + i] = sc;
        c->offset[b * bd + i] = off;
      }
    }
  } else if (T == "FixedScaleComponent" || T == "PerElementScaleComponent" || T == "NaturalGradientPerElementScaleComponent") {
    c->in_dim = c->out_dim = (int)c->scale.size();
  } else if (T == "FixedBiasComponent" || T == "PerElementOffsetComponent") {
    c->in_dim = c->out_dim = (int)c->offset.size();
  } else if (T == "ScaleAndOffsetComponent") {
    if (dim <= 0 || c->scale.empty() || dim % (int)c->scale.size()) RS_FAIL(r.path() << ": bad ScaleAndOffsetComponent");
    int bd = (int)c->scale.size();
    std::vector<float> s(dim), o(dim);
    for (int i = 0; i < dim; i++) {
      s[i] = c->scale[i % bd];
      o[i] = c->offset[i % bd];
    }
    c->scale = s;
    c->offset = o;
    c->in_dim = c->out_dim = dim;
  } else {
    if (dim <= 0) RS_FAIL(r.path() << ": component " << T << " without <Dim>");
    c->in_dim = c->out_dim = dim;
  }
}

// ---------------------------------------------------------------------------------------------
// config lines + descriptors (nnet3/nnet-nnet.cc ReadConfig, nnet-descriptor.cc)

static std::vector<std::string> TokenizeDescriptor(const std::string &s) {
  std::vector<std::string> out;
  std::string cur;
  for (char ch : s) {
    if (ch == '(' || ch == ')' || ch == ',' || isspace((unsigned char)ch)) {
      if (!cur.empty()) out.push_back(cur), cur.clear();
      if (ch == '(' || ch == ')' || ch == ',') out.push_back(std::string(1, ch));
    } else {
      cur.push_back(ch);
    }
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

struct DescParser {
  const Nnet3 &net;
  std::vector<std::string> tok;
  size_t p = 0;
  std::string line;
  const std::string &Next() {
    if (p >= tok.size()) RS_FAIL("descriptor ends early: " << line);
    return tok[p++];
  }
  void Expect(const char *s) {
    if (Next() != s) RS_FAIL("expected '" << s << "' in descriptor: " << line);
  }
  Descriptor Parse() {
    std::string t = Next();
    if (p < tok.size() && tok[p] == "(") {
      p++;
      Descriptor d;
      if (t == "Append") {
        while (true) {
          Descriptor sub = Parse();
          d.insert(d.end(), sub.begin(), sub.end());
          std::string n = Next();
          if (n == ")") break;
          if (n != ",") RS_FAIL("bad Append in descriptor: " << line);
        }
      } else if (t == "Sum") {
        d = Parse();
        while (true) {
          std::string n = Next();
          if (n == ")") break;
          if (n != ",") RS_FAIL("bad Sum in descriptor: " << line);
          Descriptor b = Parse();
          if (d.size() != 1 || b.size() != 1 || d[0].dim != b[0].dim)
            RS_FAIL("Sum() over Append() blocks is not supported: " << line);
          d[0].terms.insert(d[0].terms.end(), b[0].terms.begin(), b[0].terms.end());
        }
      } else if (t == "Offset") {
        d = Parse();
        Expect(",");
        int off = atoi(Next().c_str());
        std::string n = Next();
        if (n == ",") {
          if (atoi(Next().c_str()) != 0) RS_FAIL("Offset() with an x offset is not supported: " << line);
          n = Next();
        }
        if (n != ")") RS_FAIL("bad Offset in descriptor: " << line);
        for (auto &part : d)
          for (auto &term : part.terms)
            if (!term.const_time) term.t_offset += off;
      } else if (t == "Scale") {
        float a = (float)atof(Next().c_str());
        Expect(",");
        d = Parse();
        Expect(")");
        for (auto &part : d)
          for (auto &term : part.terms) term.scale *= a;
      } else if (t == "ReplaceIndex") {
        d = Parse();
        Expect(",");
        std::string var = Next();
        Expect(",");
        int val = atoi(Next().c_str());
        Expect(")");
        if (var != "t" || val != 0) RS_FAIL("only ReplaceIndex(x, t, 0) is supported: " << line);
        for (auto &part : d)
          for (auto &term : part.terms) {
            term.const_time = true;
            term.t_offset = 0;
          }
      } else if (t == "IfDefined" || t == "Failover") {
        RS_FAIL("descriptor function " << t << " is not supported: " << line);
      } else {
        RS_FAIL("unknown descriptor function " << t << ": " << line);
      }
      return d;
    }
    int n = net.FindNode(t);
    if (n < 0) RS_FAIL("descriptor refers to unknown node '" << t << "': " << line);
    DescPart part;
    DescTerm term;
    term.node = n;
    part.terms.push_back(term);
    part.dim = net.nodes[n].dim;
    return Descriptor{part};
  }
};

int Nnet3::FindNode(const std::string &name) const {
  for (size_t i = 0; i < nodes.size(); i++)
    if (nodes[i].name == name && nodes[i].kind != Node::kOutput) return (int)i;
  return -1;
}

static std::map<std::string, std::string> ParseConfigLine(const std::string &line, std::string *first) {
  // "component-node name=x component=y input=Append(a, b)": values may contain spaces
  std::map<std::string, std::string> kv;
  std::istringstream is(line);
  is >> *first;
  std::string rest;
  std::getline(is, rest);
  std::vector<std::pair<size_t, size_t>> keys;  // (start of key, position of '=')
  for (size_t i = 0; i < rest.size(); i++) {
    if (rest[i] == '=') {
      size_t s = i;
      while (s > 0 && (isalnum((unsigned char)rest[s - 1]) || rest[s - 1] == '-' || rest[s - 1] == '_')) s--;
      if (s < i && (s == 0 || isspace((unsigned char)rest[s - 1]))) keys.push_back({s, i});
    }
  }
  for (size_t k = 0; k < keys.size(); k++) {
    size_t vend = k + 1 < keys.size() ? keys[k + 1].first : rest.size();
    kv[rest.substr(keys[k].first, keys[k].second - keys[k].first)] =
        Trim(rest.substr(keys[k].second + 1, vend - keys[k].second - 1));
  }
  return kv;
}

static void ReadNnet3(KaldiReader &r, Nnet3 *net) {
  r.ExpectToken("<Nnet3>");
  std::string l = r.ReadLine();
  if (!Trim(l).empty()) RS_FAIL(r.path() << ": expected newline after <Nnet3>");
  std::vector<std::string> lines;
  while (true) {
    if (r.eof()) RS_FAIL(r.path() << ": unterminated nnet3 config section");
    l = r.ReadLine();
    if (Trim(l).empty()) break;
    lines.push_back(l);
  }
  r.ExpectToken("<NumComponents>");
  int nc = r.ReadInt32();
  net->components.resize(nc);
  std::map<std::string, int> comp_index;
  for (int i = 0; i < nc; i++) {
    r.ExpectToken("<ComponentName>");
    std::string name = r.ReadToken();
    ReadComponent(r, &net->components[i]);
    net->components[i].name = name;
    comp_index[name] = i;
  }
  r.ExpectToken("</Nnet3>");
  // pass 1: create nodes with dims; pass 2: parse descriptors (they may refer forward)
  struct Pending {
    int node;
    std::string desc;
  };
  std::vector<Pending> pending;
  for (const std::string &line : lines) {
    std::string first;
    auto kv = ParseConfigLine(line, &first);
    Node n;
    n.name = kv["name"];
    if (first == "input-node") {
      n.kind = Node::kInput;
      n.dim = atoi(kv["dim"].c_str());
    } else if (first == "component-node") {
      n.kind = Node::kComponent;
      auto it = comp_index.find(kv["component"]);
      if (it == comp_index.end()) RS_FAIL(r.path() << ": unknown component in: " << line);
      n.component = it->second;
      n.dim = net->components[n.component].out_dim;
      pending.push_back({(int)net->nodes.size(), kv["input"]});
    } else if (first == "output-node") {
      n.kind = Node::kOutput;
      pending.push_back({(int)net->nodes.size(), kv["input"]});
    } else if (first == "dim-range-node") {
      n.kind = Node::kDimRange;
      n.dim = atoi(kv["dim"].c_str());
      n.dim_offset = atoi(kv["dim-offset"].c_str());
      pending.push_back({(int)net->nodes.size(), kv["input-node"]});
    } else if (first == "component") {
      continue;
    } else {
      RS_FAIL(r.path() << ": unsupported nnet3 config line: " << line);
    }
    net->nodes.push_back(n);
  }
  for (const Pending &pd : pending) {
    DescParser dp{*net, TokenizeDescriptor(pd.desc), 0, pd.desc};
    net->nodes[pd.node].input = dp.Parse();
    if (dp.p != dp.tok.size()) RS_FAIL(r.path() << ": trailing tokens in descriptor: " << pd.desc);
    // dims of parts that refer to later-defined nodes were unknown during parsing: refresh
    for (auto &part : net->nodes[pd.node].input) part.dim = net->nodes[part.terms[0].node].dim;
    if (net->nodes[pd.node].kind == Node::kOutput) {
      int d = 0;
      for (auto &part : net->nodes[pd.node].input) d += part.dim;
      net->nodes[pd.node].dim = d;
    }
  }
  for (auto &n : net->nodes)
    if (n.kind == Node::kDimRange) RS_FAIL(r.path() << ": dim-range-node is not supported (recurrent model?)");
}

// ---------------------------------------------------------------------------------------------

void LoadModel(const std::string &final_mdl, const std::string &online_conf, Model *m) {
  auto kv = ReadConfigFile(online_conf);
  std::string feature_type = "mfcc", mfcc_config, ivector_config, cmvn_config, global_cmvn_stats;
  Take(kv, "feature-type", &feature_type);
  Take(kv, "mfcc-config", &mfcc_config);
  Take(kv, "ivector-extraction-config", &ivector_config);
  Take(kv, "cmvn-config", &cmvn_config);
  Take(kv, "global-cmvn-stats", &global_cmvn_stats);
  Take(kv, "frame-subsampling-factor", &m->frame_subsampling_factor);
  bool add_pitch = false;
  Take(kv, "add-pitch", &add_pitch);
  int extra_left_initial = 0, frames_per_chunk = 24;  // NnetSimpleLoopedComputationOptions (decodable-simple-looped.h:54-58)
  Take(kv, "extra-left-context-initial", &extra_left_initial);
  Take(kv, "frames-per-chunk", &frames_per_chunk);
  // endpointing and silence weighting are inactive as the reference invokes the decoders
  // (transcribe_wav.py:49 --do-endpointing=false; online-ivector-feature.h:431-436)
  for (auto it = kv.begin(); it != kv.end();) {
    if (it->first.compare(0, 9, "endpoint.") == 0 || it->first == "silence-weight" || it->first == "silence-phones" ||
        it->first == "max-state-duration" || it->first == "online-pitch-config" || it->first == "fbank-config" ||
        it->first == "plp-config" || it->first == "acoustic-scale" || it->first == "debug-computation")
      it = kv.erase(it);
    else
      ++it;
  }
  if (!kv.empty()) RS_FAIL(online_conf << ": unsupported option --" << kv.begin()->first);
  if (feature_type != "mfcc") RS_FAIL(online_conf << ": only --feature-type=mfcc is supported, got " << feature_type);
  if (add_pitch) RS_FAIL(online_conf << ": --add-pitch=true is not supported");
  if (extra_left_initial != 0) RS_FAIL(online_conf << ": --extra-left-context-initial != 0 is not supported");
  if (frames_per_chunk < 1) RS_FAIL(online_conf << ": bad --frames-per-chunk");
  m->frames_per_chunk = frames_per_chunk;
  if (m->frame_subsampling_factor < 1) RS_FAIL(online_conf << ": bad --frame-subsampling-factor");
  if (!mfcc_config.empty()) ParseMfccConf(mfcc_config, &m->mfcc);
  if (!cmvn_config.empty()) {
    m->nnet_cmvn = true;
    ParseCmvnConf(cmvn_config, &m->nnet_cmvn_opts);
    if (global_cmvn_stats.empty()) RS_FAIL(online_conf << ": --cmvn-config needs --global-cmvn-stats");
    KaldiReader r(global_cmvn_stats);
    m->nnet_global_cmvn = r.ReadMatrixD();
  }
  if (!ivector_config.empty()) {
    m->has_ivector = true;
    auto iv = ReadConfigFile(ivector_config);
    std::string splice_config, cmvn_conf, lda, gstats, ubm, ie;
    Take(iv, "splice-config", &splice_config);
    Take(iv, "cmvn-config", &cmvn_conf);
    Take(iv, "lda-matrix", &lda);
    Take(iv, "global-cmvn-stats", &gstats);
    Take(iv, "diag-ubm", &ubm);
    Take(iv, "ivector-extractor", &ie);
    Take(iv, "online-cmvn-iextractor", &m->ivec.online_cmvn_iextractor);
    Take(iv, "ivector-period", &m->ivec.ivector_period);
    Take(iv, "num-gselect", &m->ivec.num_gselect);
    Take(iv, "min-post", &m->ivec.min_post);
    Take(iv, "posterior-scale", &m->ivec.posterior_scale);
    Take(iv, "max-count", &m->ivec.max_count);
    Take(iv, "num-cg-iters", &m->ivec.num_cg_iters);
    Take(iv, "max-remembered-frames", &m->ivec.max_remembered_frames);
    bool umr = true, greedy = false;
    Take(iv, "use-most-recent-ivector", &umr);
    Take(iv, "greedy-ivector-extractor", &greedy);
    // the two iVector schedules built here are the defaults of online-ivector-feature.h:93-112: the most recent estimate
    // for every frame, no greedy history (online-ivector-feature.cc:348-353); anything else must not be answered silently
    if (!umr) RS_FAIL(ivector_config << ": --use-most-recent-ivector=false (periodic iVector history) is not supported");
    if (greedy) RS_FAIL(ivector_config << ": --greedy-ivector-extractor=true is not supported");
    if (!iv.empty()) RS_FAIL(ivector_config << ": unsupported option --" << iv.begin()->first);
    if (lda.empty() || gstats.empty() || ubm.empty() || ie.empty())
      RS_FAIL(ivector_config << ": --lda-matrix, --global-cmvn-stats, --diag-ubm and --ivector-extractor are required");
    if (!splice_config.empty()) {
      auto sp = ReadConfigFile(splice_config);
      Take(sp, "left-context", &m->ivec.splice_left);
      Take(sp, "right-context", &m->ivec.splice_right);
      if (!sp.empty()) RS_FAIL(splice_config << ": unsupported option --" << sp.begin()->first);
    }
    if (!cmvn_conf.empty()) ParseCmvnConf(cmvn_conf, &m->ivec.cmvn);
    {
      KaldiReader r(lda);
      m->lda = r.ReadMatrixF();
    }
    {
      KaldiReader r(gstats);
      m->global_cmvn = r.ReadMatrixD();
    }
    ReadDiagGmm(ubm, &m->ubm);
    ReadIvectorExtractor(ie, &m->ie);
    int D = m->mfcc.num_ceps, ns = m->ivec.splice_left + 1 + m->ivec.splice_right;
    if (m->lda.cols != D * ns && m->lda.cols != D * ns + 1) RS_FAIL(lda << ": LDA matrix does not match spliced feature dim");
    if (m->lda.rows != m->ubm.dim || m->ubm.dim != m->ie.feat_dim || m->ubm.num_gauss != m->ie.num_gauss)
      RS_FAIL(ivector_config << ": LDA / UBM / extractor dimensions disagree");
    if (m->global_cmvn.rows != 2 || m->global_cmvn.cols != D + 1) RS_FAIL(gstats << ": bad global CMVN stats shape");
    if (m->ivec.num_gselect < 1) RS_FAIL(ivector_config << ": --num-gselect must be >= 1");
  }
  KaldiReader r(final_mdl);
  ReadTransitionModel(r, &m->trans);
  ReadNnet3(r, &m->nnet);
  r.ExpectToken("<LeftContext>");
  m->nnet.left_context = r.ReadInt32();
  r.ExpectToken("<RightContext>");
  m->nnet.right_context = r.ReadInt32();
  r.ExpectToken("<Priors>");
  m->nnet.priors = r.ReadVectorF();
  m->log_priors.clear();
  for (float p : m->nnet.priors) m->log_priors.push_back(logf(p));
  CompilePlan(m->nnet, m->frame_subsampling_factor, &m->plan);
  const Plan &pl = m->plan;
  if (pl.buffers[pl.output_buffer].dim != m->trans.num_pdfs)
    RS_FAIL(final_mdl << ": nnet output dim " << pl.buffers[pl.output_buffer].dim << " != number of pdfs " << m->trans.num_pdfs);
  if (!m->log_priors.empty() && (int)m->log_priors.size() != m->trans.num_pdfs) RS_FAIL(final_mdl << ": priors dim mismatch");
  if (pl.buffers[pl.input_buffer].dim != m->mfcc.num_ceps) RS_FAIL(final_mdl << ": nnet input dim != MFCC dim");
  if ((pl.ivector_buffer >= 0) != m->has_ivector) RS_FAIL(final_mdl << ": model and online.conf disagree about iVectors");
  if (m->has_ivector && pl.buffers[pl.ivector_buffer].dim != m->ie.ivector_dim) RS_FAIL(final_mdl << ": iVector dim mismatch");
}

// ---------------------------------------------------------------------------------------------
// plan compiler

static int Gcd(int a, int b) {
  a = std::abs(a);
  b = std::abs(b);
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

static bool IsAffine(const Component &c) {
  return c.type == "FixedAffineComponent" || c.type == "AffineComponent" || c.type == "NaturalGradientAffineComponent" ||
         c.type == "LinearComponent" || c.type == "TdnnComponent";
}

void CompilePlan(const Nnet3 &net, int sf, Plan *plan) {
  *plan = Plan();
  plan->frame_subsampling_factor = sf;
  int out_node = -1;
  for (size_t i = 0; i < net.nodes.size(); i++)
    if (net.nodes[i].kind == Node::kOutput && net.nodes[i].name == "output") out_node = (int)i;
  if (out_node < 0) RS_FAIL("nnet3 model has no output-node named 'output'");
  const int N = (int)net.nodes.size();
  // reachable set + topological order
  std::vector<int> order, state(N, 0);
  std::function<void(int)> visit = [&](int n) {
    if (state[n] == 2) return;
    if (state[n] == 1) RS_FAIL("recurrent nnet3 models are not supported (cycle at node " << net.nodes[n].name << ")");
    state[n] = 1;
    for (const auto &part : net.nodes[n].input)
      for (const auto &term : part.terms) visit(term.node);
    state[n] = 2;
    order.push_back(n);
  };
  visit(out_node);
  // consumers
  std::vector<std::vector<int>> consumers(N);
  for (int n : order)
    for (const auto &part : net.nodes[n].input)
      for (const auto &term : part.terms) consumers[term.node].push_back(n);
  // time steps per node: gcd over consumers' steps and all offsets used to reach this node
  std::vector<int> step(N, 0), lo(N, 0), hi(N, 0);  // lo/hi: needed context relative to the output range
  step[out_node] = sf;
  for (auto it = order.rbegin(); it != order.rend(); ++it) {
    int n = *it;
    const Node &nd = net.nodes[n];
    std::vector<int> offs = {0};
    if (nd.kind == Node::kComponent && net.components[nd.component].type == "TdnnComponent")
      offs = net.components[nd.component].time_offsets;
    for (const auto &part : nd.input)
      for (const auto &term : part.terms) {
        if (term.const_time) continue;
        for (int o : offs) {
          int tot = o + term.t_offset;
          step[term.node] = Gcd(Gcd(step[term.node], step[n]), tot);
          lo[term.node] = std::min(lo[term.node], lo[n] + tot);
          hi[term.node] = std::max(hi[term.node], hi[n] + tot);
        }
      }
  }
  // buffers: one per input / component node (fused nodes are redirected below)
  std::vector<int> node_buffer(N, -1);
  auto new_buffer = [&](const std::string &name, int dim, int st) {
    PlanBuffer b;
    b.name = name;
    b.dim = dim;
    b.step = std::max(st, 1);
    plan->buffers.push_back(b);
    return (int)plan->buffers.size() - 1;
  };
  auto add_vec = [&](const std::vector<float> &v) {
    plan->vectors.push_back(v);
    return (int)plan->vectors.size() - 1;
  };
  for (int n : order) {
    const Node &nd = net.nodes[n];
    if (nd.kind != Node::kInput) continue;
    int b = new_buffer(nd.name, nd.dim, step[n]);
    node_buffer[n] = b;
    if (nd.name == "input") {
      plan->input_buffer = b;
      plan->buffers[b].is_input = true;
      plan->left_context = -lo[n];
      plan->right_context = hi[n];
    } else if (nd.name == "ivector") {
      plan->ivector_buffer = b;
      plan->buffers[b].is_ivector = true;
      plan->buffers[b].per_utt = true;
    } else {
      RS_FAIL("unsupported nnet3 input node '" << nd.name << "'");
    }
  }
  if (plan->input_buffer < 0) RS_FAIL("nnet3 model has no input-node named 'input'");

  // an elementwise node can be fused into its producer's step when it consumes exactly that node
  auto elementwise_ops = [&](const Component &c, std::vector<EpiOp> *ops) -> bool {
    EpiOp op;
    if (c.type == "RectifiedLinearComponent") {
      op.type = EpiOp::kRelu;
      ops->push_back(op);
    } else if (c.type == "BatchNormComponent" || c.type == "ScaleAndOffsetComponent") {
      op.type = EpiOp::kScaleOffset;
      op.vec0 = add_vec(c.scale);
      op.vec1 = add_vec(c.offset);
      ops->push_back(op);
    } else if (c.type == "FixedScaleComponent" || c.type == "PerElementScaleComponent" ||
               c.type == "NaturalGradientPerElementScaleComponent") {
      op.type = EpiOp::kScaleOffset;
      op.vec0 = add_vec(c.scale);
      op.vec1 = add_vec(std::vector<float>(c.scale.size(), 0.f));
      ops->push_back(op);
    } else if (c.type == "FixedBiasComponent" || c.type == "PerElementOffsetComponent") {
      op.type = EpiOp::kBias;
      op.vec0 = add_vec(c.offset);
      ops->push_back(op);
    } else if (c.type == "GeneralDropoutComponent" || c.type == "DropoutComponent" || c.type == "NoOpComponent") {
      // identity in test mode (online2-wav-nnet3-latgen-faster.cc:169-170 forces test mode)
    } else {
      return false;
    }
    return true;
  };

  std::vector<int> node_step_index(N, -1);  // plan step that produces the node's buffer
  for (int n : order) {
    const Node &nd = net.nodes[n];
    if (nd.kind == Node::kInput) continue;
    if (nd.kind == Node::kOutput) {
      // the output node is a plain alias of its (single) input
      if (nd.input.size() != 1 || nd.input[0].terms.size() != 1 || nd.input[0].terms[0].t_offset != 0 ||
          nd.input[0].terms[0].scale != 1.f || nd.input[0].terms[0].const_time)
        RS_FAIL("output-node with a non-trivial descriptor is not supported");
      node_buffer[n] = node_buffer[nd.input[0].terms[0].node];
      continue;
    }
    const Component &c = net.components[nd.component];
    int in_dim = 0;
    for (const auto &part : nd.input) in_dim += part.dim;
    if (in_dim != c.in_dim) RS_FAIL("node " << nd.name << ": input dim " << in_dim << " != component input dim " << c.in_dim);
    if (IsAffine(c)) {
      Step st;
      st.type = Step::kGemm;
      st.name = nd.name;
      st.n = c.out_dim;
      st.ktot = c.linear.cols;
      plan->matrices.push_back(c.linear);
      st.weight = (int)plan->matrices.size() - 1;
      std::vector<int> offs = c.time_offsets;
      // per-utterance (const_time) parts become a separate tiny GEMM whose result is a per-utt bias
      Step utt;
      utt.type = Step::kUttGemm;
      utt.name = nd.name + ".utt";
      utt.n = c.out_dim;
      utt.weight = st.weight;
      utt.ktot = st.ktot;
      for (size_t oi = 0; oi < offs.size(); oi++) {
        int col = (int)oi * c.in_dim;
        for (const auto &part : nd.input) {
          if (part.terms.size() != 1 || part.terms[0].scale != 1.f) {
            RS_FAIL("node " << nd.name << ": Sum()/Scale() inside the input of an affine component is not supported; "
                    "only as the input of an elementwise node");
          }
          const DescTerm &t = part.terms[0];
          Slab s;
          s.src = node_buffer[t.node];
          s.k = part.dim;
          s.wcol = col;
          if (t.const_time) {
            s.t_offset = 0;
            utt.slabs.push_back(s);
          } else {
            s.t_offset = offs[oi] + t.t_offset;
            st.slabs.push_back(s);
          }
          col += part.dim;
        }
      }
      int b = new_buffer(nd.name, c.out_dim, step[n]);
      node_buffer[n] = b;
      st.out = b;
      if (!utt.slabs.empty()) {
        int ub = new_buffer(nd.name + ".uttbias", c.out_dim, 1);
        plan->buffers[ub].per_utt = true;
        utt.out = ub;
        if (!c.bias.empty()) {
          EpiOp op;
          op.type = EpiOp::kBias;
          op.vec0 = add_vec(c.bias);
          utt.ops.push_back(op);
        }
        plan->steps.push_back(utt);
        EpiOp op;
        op.type = EpiOp::kUttBias;
        op.buffer = ub;
        st.ops.push_back(op);
      } else if (!c.bias.empty()) {
        EpiOp op;
        op.type = EpiOp::kBias;
        op.vec0 = add_vec(c.bias);
        st.ops.push_back(op);
      }
      if (st.slabs.empty()) RS_FAIL("node " << nd.name << ": affine component with only per-utterance inputs");
      plan->steps.push_back(st);
      node_step_index[n] = (int)plan->steps.size() - 1;
      continue;
    }
    if (c.type == "LogSoftmaxComponent") {
      if (nd.input.size() != 1 || nd.input[0].terms.size() != 1 || nd.input[0].terms[0].t_offset != 0 ||
          nd.input[0].terms[0].scale != 1.f)
        RS_FAIL("node " << nd.name << ": LogSoftmax with a non-trivial input is not supported");
      Step st;
      st.type = Step::kLogSoftmax;
      st.name = nd.name;
      st.n = c.out_dim;
      Slab s;
      s.src = node_buffer[nd.input[0].terms[0].node];
      s.k = c.out_dim;
      st.slabs.push_back(s);
      st.out = new_buffer(nd.name, c.out_dim, step[n]);
      node_buffer[n] = st.out;
      plan->steps.push_back(st);
      node_step_index[n] = (int)plan->steps.size() - 1;
      continue;
    }
    std::vector<EpiOp> ops;
    if (!elementwise_ops(c, &ops)) RS_FAIL("node " << nd.name << ": component type " << c.type << " is not supported");
    // Try to fuse into the producing step.  Pattern A: input is exactly one node, unscaled, no
    // offset, produced by a step, and this node is its only consumer.  Pattern B (TDNN-F bypass):
    // Sum(Scale(a, other), x) where x matches pattern A and 'other' lives on the same time grid.
    int fuse_src = -1, fuse_term = -1;
    if (nd.input.size() == 1) {
      const auto &terms = nd.input[0].terms;
      for (size_t ti = 0; ti < terms.size(); ti++) {
        const DescTerm &t = terms[ti];
        if (t.const_time || t.t_offset != 0 || t.scale != 1.f) continue;
        if (node_step_index[t.node] < 0 || consumers[t.node].size() != 1) continue;
        if (plan->steps[node_step_index[t.node]].type == Step::kLogSoftmax) continue;
        if (step[t.node] != step[n]) continue;
        // the candidate must be the most recently produced of the terms (others already exist)
        fuse_src = t.node;
        fuse_term = (int)ti;
      }
      if (fuse_src >= 0) {
        for (size_t ti = 0; ti < terms.size(); ti++) {
          if ((int)ti == fuse_term) continue;
          const DescTerm &t = terms[ti];
          // the other summand may live on a finer time grid (its step divides ours)
          if (t.const_time || t.t_offset != 0 || node_buffer[t.node] < 0 ||
              std::max(step[n], 1) % plan->buffers[node_buffer[t.node]].step != 0)
            fuse_src = -1;
        }
      }
    }
    if (fuse_src >= 0) {
      Step &st = plan->steps[node_step_index[fuse_src]];
      const auto &terms = nd.input[0].terms;
      for (size_t ti = 0; ti < terms.size(); ti++) {
        if ((int)ti == fuse_term) continue;
        EpiOp op;
        op.type = EpiOp::kAddScaled;
        op.alpha = terms[ti].scale;
        op.buffer = node_buffer[terms[ti].node];
        st.ops.push_back(op);
      }
      st.ops.insert(st.ops.end(), ops.begin(), ops.end());
      st.name += "+" + nd.name;
      node_buffer[n] = st.out;
      plan->buffers[st.out].name = nd.name;
      node_step_index[n] = node_step_index[fuse_src];
      continue;
    }
    // generic elementwise step(s): one per Append part
    int b = new_buffer(nd.name, c.out_dim, step[n]);
    node_buffer[n] = b;
    int col = 0;
    for (size_t pi = 0; pi < nd.input.size(); pi++) {
      const auto &part = nd.input[pi];
      Step st;
      st.type = Step::kElementwise;
      st.name = nd.name;
      st.out = b;
      st.n = part.dim;
      st.col_offset = col;
      for (const auto &t : part.terms) {
        if (t.const_time) RS_FAIL("node " << nd.name << ": per-utterance input to an elementwise node is not supported");
        Slab s;
        s.src = node_buffer[t.node];
        s.t_offset = t.t_offset;
        s.k = part.dim;
        st.slabs.push_back(s);
        st.term_scale.push_back(t.scale);
      }
      // ops carry full-width vectors; slice them for this part
      for (EpiOp op : ops) {
        if (op.vec0 >= 0 && nd.input.size() > 1) {
          std::vector<float> v(plan->vectors[op.vec0].begin() + col, plan->vectors[op.vec0].begin() + col + part.dim);
          op.vec0 = add_vec(v);
        }
        if (op.vec1 >= 0 && nd.input.size() > 1) {
          std::vector<float> v(plan->vectors[op.vec1].begin() + col, plan->vectors[op.vec1].begin() + col + part.dim);
          op.vec1 = add_vec(v);
        }
        st.ops.push_back(op);
      }
      plan->steps.push_back(st);
      col += part.dim;
    }
    node_step_index[n] = (int)plan->steps.size() - 1;
  }
  plan->output_buffer = node_buffer[out_node];
  plan->buffers[plan->output_buffer].is_output = true;
  if (plan->buffers[plan->output_buffer].step != sf)
    RS_FAIL("internal: output step " << plan->buffers[plan->output_buffer].step << " != frame-subsampling-factor " << sf);
  int align = 1;
  for (const auto &b : plan->buffers)
    if (!b.per_utt) align = align / Gcd(align, b.step) * b.step;
  plan->align = align;

  // liveness -> physical slots (buffers of equal (dim, step) share storage once dead)
  const int NB = (int)plan->buffers.size();
  std::vector<int> last_use(NB, -1);
  for (size_t si = 0; si < plan->steps.size(); si++) {
    const Step &st = plan->steps[si];
    for (const auto &s : st.slabs) last_use[s.src] = (int)si;
    for (const auto &op : st.ops)
      if (op.buffer >= 0) last_use[op.buffer] = (int)si;
    last_use[st.out] = std::max(last_use[st.out], (int)si);
  }
  last_use[plan->output_buffer] = (int)plan->steps.size() + 1;
  last_use[plan->input_buffer] = (int)plan->steps.size() + 1;
  if (plan->ivector_buffer >= 0) last_use[plan->ivector_buffer] = (int)plan->steps.size() + 1;
  struct Slot {
    int dim, step;
    bool per_utt;
    int free_at;
  };
  std::vector<Slot> slots;
  auto assign = [&](int b, int at) {
    PlanBuffer &pb = plan->buffers[b];
    if (pb.slot >= 0) return;
    for (size_t s = 0; s < slots.size(); s++)
      if (slots[s].free_at < at && slots[s].dim == pb.dim && slots[s].step == pb.step && slots[s].per_utt == pb.per_utt &&
          !pb.is_input && !pb.is_ivector && !pb.is_output) {
        pb.slot = (int)s;
        slots[s].free_at = last_use[b];
        return;
      }
    slots.push_back({pb.dim, pb.step, pb.per_utt, (pb.is_input || pb.is_ivector || pb.is_output) ? (1 << 30) : last_use[b]});
    pb.slot = (int)slots.size() - 1;
  };
  assign(plan->input_buffer, -1);
  if (plan->ivector_buffer >= 0) assign(plan->ivector_buffer, -1);
  for (size_t si = 0; si < plan->steps.size(); si++) assign(plan->steps[si].out, (int)si);
  plan->num_slots = (int)slots.size();
}

double Plan::flops_per_axis_row(int /*step*/) const { return 0.0; }

std::string DescribePlan(const Plan &plan) {
  std::ostringstream os;
  os << "plan: " << plan.steps.size() << " steps, " << plan.buffers.size() << " buffers in " << plan.num_slots
     << " slots, context -" << plan.left_context << "/+" << plan.right_context << ", align " << plan.align << "\n";
  for (const Step &st : plan.steps) {
    static const char *names[] = {"gemm", "elementwise", "logsoftmax", "uttgemm"};
    os << "  " << names[st.type] << " " << st.name << " -> buf" << st.out << "(slot " << plan.buffers[st.out].slot << ", step "
       << plan.buffers[st.out].step << ") n=" << st.n << " k=" << st.ktot << " slabs[";
    for (const auto &s : st.slabs) os << " buf" << s.src << "@" << s.t_offset << ":k" << s.k << "/w" << s.wcol;
    os << " ] ops[";
    static const char *on[] = {"bias", "relu", "scaleoffset", "scale", "addscaled", "uttbias"};
    for (const auto &op : st.ops) {
      os << " " << on[op.type];
      if (op.type == EpiOp::kAddScaled) os << "(" << op.alpha << "*buf" << op.buffer << ")";
    }
    os << " ]\n";
  }
  return os.str();
}

}  // namespace rs
