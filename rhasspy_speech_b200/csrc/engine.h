// Device-side parameter blocks and launchers shared by the stage kernels and the host engine.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace rs {

// ---------------------------------------------------------------------------- stage (i) MFCC
struct FeatParams {
  // batch
  const int16_t *pcm;        // all utterances back to back
  const int64_t *pcm_offset; // [n_utts] first sample of each utterance
  const int *num_frames;     // [n_utts]
  const int *frame_offset;   // [n_utts] first row of each utterance in `mfcc`
  float *mfcc;               // [total_frames, num_ceps]
  // options
  int shift, length, padded, logn;  // logn = log2(padded / 2)
  float preemph, dither, energy_floor;
  int remove_dc, use_energy, raw_energy;
  uint32_t seed;
  int num_bins, num_ceps;
  // tables (device)
  const float *window;          // [length]
  const uint16_t *level_offsets;  // split-radix block offsets, grouped by log2(block size)
  int level_start[12], level_count[12];
  const float *twiddle;         // per level: cn, spcn, smcn, c3n, spc3n, smc3n (each m/4-2 long)
  int twiddle_start[12];
  const uint16_t *perm;         // bit-reversal permutation of the complex FFT output [padded/2]
  const float *kn;              // real-FFT twiddles (re, im) for k = 1 .. padded/4
  const int *mel_offset, *mel_len, *mel_start;  // [num_bins]
  const float *mel_weights;
  const float *dct;             // [num_ceps, num_bins]
  const float *mel_weights_t;   // [mel_max_len][num_bins] tap-major copy of mel_weights, zero beyond a bin's length
  int mel_max_len;
  const float *dct_t;           // [num_bins][num_ceps]
  const float *lifter;          // [num_ceps] or null
};
void LaunchMfcc(const FeatParams &p, int n_utts, int max_frames, cudaStream_t stream);

// --------------------------------------------------------------------- stage (i) CMVN + iVector
struct CmvnParams {
  const float *in;   // [total_frames, dim]
  float *out;        // [total_frames, dim]
  const int *num_frames, *frame_offset;
  const double *global_stats;  // [2, dim+1]
  int dim, cmn_window, global_frames;
  int normalize_mean, normalize_variance;
};
void LaunchCmvn(const CmvnParams &p, int n_utts, cudaStream_t stream);

struct IvecParams {
  const float *mfcc, *mfcc_norm;      // [total_frames, dim]
  const int *num_frames, *frame_offset;
  int n_utts, total_frames, max_frames;
  // iVector solves: solve j accumulates the statistics of the first v_num_frames[j] frames of the
  // utterance whose features start at v_frame_offset[j]; utterance u owns solves [v_begin[u], v_begin[u+1])
  // and runs them in order, each CG warm-started from the previous one.  Offline decoding has one solve
  // per utterance over all its frames; online decoding one per nnet chunk (online-ivector-feature.cc:248-279).
  int v_n;
  const int *v_num_frames, *v_frame_offset, *v_begin;
  int dim, left, right;               // splice
  const float *lda_t;                 // [K = dim*(left+1+right)][ldim]  (transposed), bias may be null
  const float *lda_bias;
  int ldim;                           // LDA output dim == UBM dim
  float *x_raw, *x_norm;              // [total_frames, ldim]
  // UBM
  int num_gauss;
  const float *gconsts;               // [G]
  const float *means_invvars_t;       // [ldim][G]
  const float *inv_vars_t;            // [ldim][G]
  int num_gselect;
  float min_post, posterior_scale;
  int *post_idx;                      // [total_frames, num_gselect]  (-1 = unused)
  float *post_w;                      // [total_frames, num_gselect]
  // extractor
  int ivector_dim;
  const double *sigma_inv_m;          // [G][ldim][R]
  const double *u;                    // [G][R(R+1)/2]
  double prior_offset;
  float max_count;
  int num_cg_iters;
  int online_cmvn_iextractor;
  double *wf;                         // [v_n][G][ldim] weighted feature sums
  float *gw;                          // [v_n][G] per-Gaussian total weights (float, as the reference)
  double *linear_part;                // [linear_chunks][v_n][R] split-K partial sums of the linear term
  int linear_chunks;                  // ceil(G * ldim / 512)
  double *quad;                       // [v_n][R(R+1)/2] packed lower triangle
  float *ivector;                     // [v_n][ivector_ld] nnet input (prior offset removed), one row per solve
  int ivector_ld;
};
void LaunchIvector(const IvecParams &p, cudaStream_t stream);

// ------------------------------------------------------------------------ stage (ii) nnet
constexpr int kMaxSlabs = 8;
constexpr int kMaxOps = 8;
struct GemmSlab {
  const void *src;      // fp32 matrix, or the hi plane of a split buffer (split.cuh)
  const void *src_lo;   // lo plane of a split buffer, or null
  int ld;        // row stride of src in elements
  int rows;      // valid rows of src (indices are clamped into [0, rows))
  int k;         // columns used
  int wcol;      // first weight column
  int num, den;  // src_row = (out_row * num + shift) / den   (den divides exactly on valid rows)
  int shift;
};
struct DevOp {
  int type;            // EpiOp::Type
  const float *v0, *v1;
  float alpha;
  const void *buf;     // kAddScaled: other activation buffer (fp32 or hi plane) ; kUttBias: fp32 [n_utts, ld]
  const void *buf_lo;  // kAddScaled: lo plane of a split buffer, or null
  int buf_ld, buf_rows;
  int num, den;        // kAddScaled: other_row = out_row * num / den ; kUttBias: axis time = out_row * num
};
struct GemmParams {
  GemmSlab slabs[kMaxSlabs];
  int n_slabs;
  const float *w;  // [n, ktot] row-major
  int ktot;
  void *out;
  void *out_lo;    // not null: store the result split into two fp16 planes (split.cuh)
  int *range_flag; // set if a split store saturated
  int out_ld;
  int m, n;        // output rows / columns
  DevOp ops[kMaxOps];
  int n_ops;
  const int *row_utt;  // [axis_len] utterance owning each time step (for kUttBias)
  int out_step;
};
void LaunchGemm(const GemmParams &p, cudaStream_t stream);
void LaunchElementwise(const GemmParams &p, const float *term_scale_host, int col_offset, cudaStream_t stream);
// log-softmax over the rows of `in` into p.out (plain fp32), then p.ops
void LaunchLogSoftmax(const void *in, const void *in_lo, int in_ld, const GemmParams &p, cudaStream_t stream);

struct AssembleParams {
  const float *feats;  // [total_frames, dim]
  const int *num_frames, *frame_offset, *origin;  // per utt
  void *dst;           // nnet input buffer on the global axis [axis_len, ld]
  void *dst_lo;        // not null: split store
  int *range_flag;
  int dim, ld, left, right, axis_len;
};
void LaunchAssembleInput(const AssembleParams &p, int n_utts, int max_rows, cudaStream_t stream);

// ------------------------------------------------------------------------ stage (iii) decoder
struct DevGraph {
  int num_states;
  int start;
  unsigned num_earcs, num_parcs;
  const unsigned *e_begin, *p_begin;  // [S+1]
  const int4 *earc;                   // {next, pdf, weight bits, olabel}
  const int *e_src;
  const int4 *parc;                   // {next, 0, weight bits, olabel}
  const int *p_src;
  const float *final_cost;
  int eps_flat;  // no epsilon arc leads to a state that has epsilon arcs itself (decode_small.cu)
};

struct DecodeConfig {
  float beam, beam_delta;
  int max_active, min_active;
  int tok_cap;      // max tokens inserted per frame per lane
  int hash_size;    // power of two >= 2 * tok_cap ; identity addressing if num_states <= hash_size
  int arena_cap;    // max tokens per utterance (traceback records)
  int max_words;
  int profile;      // debug: block 0 prints its per-phase clock counts (RS_B200_DECODE_PROFILE=1)
  int smem_slots;   // > 0: state tables in shared memory, addressed by state id (power of two >= num_states)
  float lattice_beam;  // lattice mode: links worse than their destination by more than this are not recorded
  int small_cache_arcs, small_ll_stage;  // decode_small.cu: arcs / log-likelihood rows staged in shared memory
};

struct LaneWorkspace {  // one per resident CTA; all pointers are device memory
  int *hkey[2];
  unsigned long long *hval[2];
  int *hidx[2];
  int *inq;          // frontier membership flag per slot (shared by both tables, always cleared)
  int *ins_list[2];  // slots inserted into table k, in insertion order
  int *tok_state[2];
  float *tok_cost[2];
  int *tok_slot;     // scratch: slot of each alive token of the frame being finalised
  unsigned *pfx;     // [tok_cap + 1] exclusive prefix of emitting out-degrees
  int *frontier[2];
  int2 *arena;       // {prev token gid, arc id}
};

// State-level lattice of one batch (GetRawLattice, lattice-faster-decoder.cc:106-189): written by
// decode_kernel<true>, pruned by lattice_prune_kernel, compacted by lattice_emit_kernel.
// Token time t = 0 .. n_frames (t = 0: closure of the start state); utterance u owns the slices
// tok[u * tok_cap ..], link[u * link_cap ..], tok_base[u * (max_t + 2) ..], link_pos[u * (2 * max_t + 4) ..].
struct LatticeBuf {
  int2 *tok;           // {state, forward cost bits} per token, in arena order (time-major)
  float *extra;        // extra_cost per token (>= 0, +inf = pruned)
  int *newid;          // compact node id of the surviving tokens
  int4 *link;          // {source token, destination token, arc id, slack bits}
  int4 *surv;          // links inside the lattice beam: {source token, destination token, arc id, time}
  int *tok_base;       // [max_t + 2] first token of each time; [n_frames + 1] = token count
  int *link_pos;       // [2 * max_t + 4]: [2t] start of the emitting links (t-1 -> t), [2t+1] start of the epsilon links at t
  float *cost_offset;  // [max_t + 1] offset added to the acoustic costs of the frame leaving time t (:733)
  int tok_cap, link_cap, surv_cap, max_t;
};

// compact lattice as the host receives it: one header per utterance, then its arcs in `arcs`
struct LatticeHeader {
  int arc_begin, n_arcs, n_nodes, ok, n_links;
  int pad[3];  // pad[0]: a link the reference holds only under some visiting orders would have survived the pruning
};
struct LatticeArc {  // dst == -1: final weight of src (graph = final cost, acoustic = 0)
  int src, dst, olabel;
  float graph, acoustic;
};

struct DecodeParams {
  DevGraph g;
  DecodeConfig cfg;
  LatticeBuf lat;         // only read by the lattice instantiation
  const float *loglikes;  // [rows, ld]
  int ld;
  const int *ll_row0;     // [n_utts] first row of each utterance
  const int *n_frames;    // [n_utts]
  int n_utts;
  LaneWorkspace *lanes;   // [gridDim.x]
  int *next_utt;          // work counter
  // decode_small.cu: one traceback arena for the batch, utterance u owns records [small_arena_off[u], small_arena_off[u + 1])
  int2 *small_arena;
  const long long *small_arena_off;
  // outputs
  int *words;             // [n_utts, max_words]
  int *n_words;           // [n_utts]  (-1 = nothing decoded)
  float *cost;            // [n_utts, 2] graph, acoustic
  int *status;            // [n_utts] 0 ok, bit0 token overflow, bit1 arena overflow, bit2 no tokens, bit3 word overflow
  unsigned long long *counters;  // [n_utts, 4] tokens expanded, arcs visited, tokens created, records written
};
void LaunchDecode(const DecodeParams &p, int n_lanes, cudaStream_t stream, bool lattice = false);
// a20: PruneForwardLinksFinal / PruneForwardLinks (lattice-faster-decoder.cc:299-458) over the recorded lattice,
// then renumbering + compaction of the surviving arcs into `arcs` (global cursor `cursor`), one CTA per utterance
void LaunchLatticePrune(const DecodeParams &p, float lattice_beam, LatticeHeader *headers, LatticeArc *arcs,
                        int arcs_cap, int *cursor, cudaStream_t stream);
// Small graphs (decode_small.cu): the reference's token order reproduced on the device -- exact unconditionally.
// The reference's token hash never has fewer than 1000 buckets, so up to 1000 states map one state to a bucket.
constexpr int kSmallMaxStates = 1000, kSmallMaxEarcs = 16384, kSmallMaxParcs = 4096;
bool DecodeSmallSupports(const DevGraph &g);
void LaunchDecodeSmall(const DecodeParams &p, cudaStream_t stream, bool lattice = false);
int DecodeCtaThreads();
size_t DecodeSmemBytes(int slots);

}  // namespace rs
