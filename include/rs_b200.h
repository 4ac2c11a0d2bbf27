/* rs_b200.h -- C ABI of the B200 batched utterance decoder (librs_b200.so).
 *
 * The reference has no FFI on this path: rhasspy-speech reaches the Kaldi decoder through argv and
 * stdin/stdout of child processes.  Each entry point below names the process-level interface it
 * replaces (paths relative to the reference checkout); INTEGRATION.md shows the ctypes binding that
 * rhasspy_speech/transcribe_wav.py and transcribe_stream.py would use instead of tools.async_run*.
 *
 * Ownership: the caller owns every input buffer for the duration of the call; the library owns
 * models, graphs, decoders and device memory until the matching *_free; results are allocated by
 * the library and released with rs_result_free.  rs_model / rs_graph are immutable and may be
 * shared by decoders; an rs_decoder (and its streams) must be used from one thread at a time.
 *
 * Errors: functions returning a pointer return NULL, functions returning int return non-zero, and
 * write a NUL-terminated message into `err` (may be NULL).  The Python layer turns that message into
 * the RuntimeError the reference raises when a Kaldi binary exits non-zero (rhasspy_speech/tools.py:81-88).
 */
#ifndef RS_B200_H_
#define RS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rs_model rs_model;
typedef struct rs_graph rs_graph;
typedef struct rs_decoder rs_decoder;
typedef struct rs_stream rs_stream;

/* Decoder options = the command-line flags rhasspy passes to online2-wav-nnet3-latgen-faster /
 * online2-cli-nnet3-decode-faster (rhasspy_speech/transcribe_wav.py:47-60, transcribe_stream.py:56-66)
 * plus LatticeFasterDecoderConfig defaults (kaldi/src/decoder/lattice-faster-decoder.h:38-92). */
typedef struct rs_decoder_opts {
  float beam;           /* --beam            (24.0) */
  int32_t max_active;   /* --max-active      (7000) */
  int32_t min_active;   /* --min-active      (200)  */
  float lattice_beam;   /* --lattice-beam    (8.0); prunes the lattice the n-best tail reads (rs_decoder_set_nbest) */
  float acoustic_scale; /* --acoustic-scale  (1.0; rhasspy always passes 1.0 to the decoder) */
  float beam_delta;     /* --beam-delta      (0.5)  */
  int32_t max_tokens_per_frame; /* device capacity per lane and frame (65536) */
  int32_t max_tokens_per_utt;   /* traceback arena per lane (4194304) */
  int32_t max_words;            /* word ids returned per hypothesis (256) */
  int32_t num_lanes;            /* resident decoder CTAs; 0 = 2 per SM */
  uint32_t dither_seed;         /* only used when mfcc.conf asks for dither */
  int32_t strict_fallback;      /* utterances whose device search was order-sensitive (status bit 4) are decoded again by
                                 * the strict-order host decoder (csrc/strict_decode.cc) from the log-likelihoods still
                                 * resident on the device: 1 (default) = those utterances, 0 = never (flag only),
                                 * 2 = every utterance (test hook) */
} rs_decoder_opts;

typedef struct rs_result {
  int32_t n_utts;
  int32_t *n_hyp;        /* [n_utts] hypotheses per utterance: 0 (nothing decoded) .. nbest */
  int32_t *word_offset;  /* [n_utts + 1] into word_ids */
  int32_t *word_ids;     /* olabels of the best path, in order (words.txt ids) */
  float *graph_cost;     /* [n_utts] of the best path */
  float *acoustic_cost;  /* [n_utts] of the best path */
  int32_t *num_frames;   /* [n_utts] decoded (subsampled) frames */
  int32_t *status;       /* [n_utts] 0 ok; errors: bit0 token capacity, bit1 arena capacity, bit2 no surviving tokens,
                          * bit3 word capacity; information: bit4 (16) the device search met a frame on which the
                          * reference's order-dependent pruning (the tokens its transient next_cutoff admits,
                          * lattice-faster-decoder.cc:780-787) could have changed a cutoff or the set of expanded tokens
                          * (decode.cu, "safe frame" rules).  With strict_fallback != 0 (default) such an utterance has been
                          * decoded again by the strict-order host decoder and the result returned is the reference's;
                          * with strict_fallback == 0 it is the device result, not guaranteed word-identical;
                          * bit5 (32) n-best requested but the lattice did not fit its device buffers: only the best path
                          * is returned; bit6 (64) the result comes from the strict-order host decoder; bits 8-11: which safe-frame
                          * rule raised bit 4 (min-active count, max-active count, extra inside the cutoff, last frame) and
                          * bit 12 (4096): a forward link only some visiting orders create would have survived the lattice
                          * pruning -- diagnostics */
  /* every hypothesis, best first (what lattice-to-nbest | nbest-to-linear print as utt-1 .. utt-n and the two
   * cost archives of nbest-to-linear): hypothesis h of utterance u is entry hyp_offset[u] + h */
  int32_t *hyp_offset;       /* [n_utts + 1] */
  int32_t *hyp_word_offset;  /* [hyp_offset[n_utts] + 1] into hyp_word_ids */
  int32_t *hyp_word_ids;
  float *hyp_graph_cost;     /* [hyp_offset[n_utts]] */
  float *hyp_acoustic_cost;  /* [hyp_offset[n_utts]] unscaled */
} rs_result;

typedef struct rs_timings {
  float h2d_ms, feature_ms, nnet_ms, decode_ms, d2h_ms, total_ms; /* CUDA-event times of the last call */
  double audio_seconds;
  uint64_t frames_decoded, tokens_expanded, arcs_visited, tokens_created, records_written;
  uint64_t nnet_flops;    /* algorithmic FLOPs of the acoustic model for the last batch */
  uint64_t h2d_bytes, d2h_bytes;
  int32_t kernel_launches;
  uint64_t nnet_bytes;    /* algorithmic HBM bytes of the acoustic model: every layer input once, bypass input, output */
  uint64_t lattice_states, lattice_arcs; /* n-best calls: states and arcs (final weights included) of the pruned
                                          * state-level lattices of the batch -- what GetRawLattice would return */
  uint64_t lattice_links_recorded;       /* forward links recorded before pruning */
  int32_t strict_utts;                   /* utterances of the last call decoded again by the strict-order host decoder */
  float strict_ms;                       /* wall time of that (log-likelihood D2H + host search, all threads) */
} rs_timings;

void rs_decoder_opts_default(rs_decoder_opts *opts);
/* CUDA devices visible to the process (0 without a driver): the multi-device pool of the Python layer deals a request
 * list over them; every device holds its own rs_model / rs_graph / rs_decoder replica. */
int rs_device_count(void);

/* Replaces the per-process model load of online2-wav-nnet3-latgen-faster.cc:150-181 (feature
 * pipeline info from --config=online.conf, TransitionModel + AmNnetSimple from final.mdl). */
rs_model *rs_model_load(const char *final_mdl, const char *online_conf, int device, char *err, size_t errlen);
void rs_model_free(rs_model *m);
int rs_model_info(const rs_model *m, int32_t *num_pdfs, int32_t *frame_subsampling_factor, int32_t *ivector_dim,
                  int32_t *feat_dim, int32_t *left_context, int32_t *right_context);

/* Replaces ReadFstKaldiGeneric(HCLG.fst) (online2-wav-nnet3-latgen-faster.cc:181) and the
 * --word-symbol-table=words.txt argument. */
rs_graph *rs_graph_load(const char *hclg_fst, const char *words_txt, int device, char *err, size_t errlen);
void rs_graph_free(rs_graph *g);
int rs_graph_info(const rs_graph *g, int32_t *num_states, int64_t *num_arcs, int32_t *num_words);
/* words.txt lookup (utils/int2sym.pl -f 2- words.txt, transcribe_wav.py:77-85); NULL if unknown */
const char *rs_graph_word(const rs_graph *g, int32_t id);

rs_decoder *rs_decoder_create(rs_model *m, rs_graph *g, const rs_decoder_opts *opts, char *err, size_t errlen);
void rs_decoder_free(rs_decoder *d);
/* Page-locked host memory for audio.  Utterances that lie back to back, in call order, inside one rs_host_alloc block are
 * copied to the device straight from that block by rs_decode_pcm; any other buffer is first packed into the decoder's
 * own pinned staging area (a host memcpy of the whole batch).  The reference has no counterpart: its audio travels
 * through a pipe into the child process (transcribe_stream.py:68-82). */
void *rs_host_alloc(size_t bytes, char *err, size_t errlen);
void rs_host_free(void *p);
/* Hot graph swap: binds the decoder to another HCLG without rebuilding its device workspace.  The
 * reference re-reads HCLG.fst in every call (ReadFstKaldiGeneric, online2-wav-nnet3-latgen-faster.cc:181), so a graph
 * retrained by KaldiTrainer._mkgraph (rhasspy_speech/kaldi.py:409-425) is picked up by the next transcription; a resident
 * decoder gets the same behaviour from rs_graph_load(new files) + this call.  The previous rs_graph stays valid and
 * is freed by its owner.  On error the decoder keeps its previous graph. */
int rs_decoder_set_graph(rs_decoder *d, rs_graph *g, char *err, size_t errlen);
/* Replaces the `lattice-to-nbest --n=<nbest> --acoustic-scale=<acoustic_scale>` stage of the pipeline
 * (transcribe_wav.py:62-75, kaldi/src/latbin/lattice-to-nbest.cc:84-113) for every later decode call on this
 * decoder.  nbest == 1 with scale 1.0 (the default) returns the device back-trace of the best path; anything
 * else records the state-level lattice on the device (GetRawLattice, lattice-faster-decoder.cc:106-189), prunes
 * it with --lattice-beam (:299-458) and returns the n cheapest distinct word sequences under
 * graph + acoustic_scale * acoustic, each with the costs of its best path.
 * The lattices of a batch share a device budget (environment RS_B200_LATTICE_MB, default 8192, split evenly over the
 * utterances of the call: 76 bytes per token).  When a lattice does not fit, the decode stage is run again with four
 * times the budget, up to RS_B200_LATTICE_MAX_MB (default 65536); the grown budget is kept for later calls (an
 * ARPA-shaped HCLG at 6 k tokens per frame needs ~60 MB per 4 s utterance).  An utterance that still does not fit returns
 * its best path and carries status bit 5. */
int rs_decoder_set_nbest(rs_decoder *d, int32_t nbest, float acoustic_scale, char *err, size_t errlen);

/* Host staging of a call (no reference counterpart: the reference reads one WAV per process, feat/wave-reader.cc).
 * on != 0 (default): the audio goes to the device in a few items on a copy stream and the MFCC kernel of an item runs
 * under the copy of the next; on == 0: one stream, copies first, then every kernel -- the stage times of rs_timings
 * are then those of a batch already resident in device memory (what bench.py reports as `value`). */
int rs_decoder_set_staging_overlap(rs_decoder *d, int32_t on, char *err, size_t errlen);

/* Replaces one run of `online2-wav-nnet3-latgen-faster --online=false ... | lattice-to-nbest --n=1 |
 * nbest-to-linear` per utterance (transcribe_wav.py:45-75), for n utterances at once.
 * pcm[i] = 16 kHz mono s16 samples, unscaled as WaveData reads them (feat/wave-reader.cc:153-320). */
int rs_decode_pcm(rs_decoder *d, const int16_t *const *pcm, const int32_t *num_samples, int32_t n, rs_result **out,
                  char *err, size_t errlen);
/* Same, reading RIFF/WAVE PCM16 files ('scp:echo utt WAV|', transcribe_wav.py:59). A sampling-rate
 * mismatch is an error as in the reference (feat/online-feature.cc:98-103). */
int rs_decode_wavs(rs_decoder *d, const char *const *paths, int32_t n, rs_result **out, char *err, size_t errlen);
/* Stage (iii) only, over caller-provided log-likelihood matrices [num_frames[i] x num_pdfs]: what
 * latgen-faster-mapped does (kaldi/src/bin/latgen-faster-mapped.cc); used by the parity tests. */
int rs_decode_loglikes(rs_decoder *d, const float *const *loglikes, const int32_t *num_frames, int32_t n,
                       rs_result **out, char *err, size_t errlen);
void rs_result_free(rs_result *r);

/* Streaming surface (online2-cli-nnet3-decode-faster fed raw s16le on stdin,
 * transcribe_stream.py:51-82).  Audio is buffered and the utterance is decoded at finish WITH THE ONLINE
 * SEMANTICS of that binary: it re-chunks its input into reads of 1024 samples, so its result is a function
 * of the audio only -- each nnet chunk uses the iVector estimated (warm-started CG) from the frames that had
 * arrived when the chunk became ready (decodable-online-looped.cc:56-84, 186-194).  rs_decode_pcm / _wavs
 * use the offline schedule of online2-wav-nnet3-latgen-faster --online=false (one iVector per utterance). */
rs_stream *rs_stream_open(rs_decoder *d, char *err, size_t errlen);
int rs_stream_accept(rs_stream *s, const int16_t *pcm, int32_t num_samples, char *err, size_t errlen);
int rs_stream_finish(rs_stream *s, rs_result **out, char *err, size_t errlen);
void rs_stream_close(rs_stream *s);
/* Finish many streams of one decoder in a single batch. */
int rs_streams_finish(rs_stream *const *streams, int32_t n, rs_result **out, char *err, size_t errlen);

/* Instrumentation of the last decode call on this decoder. */
int rs_decoder_timings(const rs_decoder *d, rs_timings *t);
/* Intermediate results of the last rs_decode_pcm / rs_decode_wavs call, for the parity tests:
 * what = 0 MFCC [T x dim], 1 iVector [solves x dim] (1 row offline, one per CG solve online), 2 log-likelihoods [T/sf x num_pdfs],
 *        3 CMVN-normalised MFCC, 4 LDA features (normalised stream),
 *        5 (after an n-best call) the pruned state-level lattice: rows of (src, dst, olabel, graph cost, acoustic cost),
 *          dst = -1 marks a final weight, state 0 is the start;
 *        6 UBM posteriors (rows a9 / a10): per frame num_gselect pairs (gaussian index, weight), unused pairs (-1, 0).
 * Call with dst == NULL to query rows/cols. */
int rs_debug_fetch(rs_decoder *d, int32_t what, int32_t utt, float *dst, int32_t *rows, int32_t *cols, char *err,
                   size_t errlen);
/* Test hook for the affine-layer kernels: one TDNN-style layer (TdnnComponent::Propagate,
 * kaldi/src/nnet3/nnet-tdnn-component.cc:181-211) on a caller-provided activation matrix.
 *   out[r, :] = sum_i src[r * stride + offsets[i], :] * w[:, i*k : (i+1)*k]^T (+ bias) (ReLU)
 * src [rows x k], w [n x (k * n_offsets)], out [max(rows / stride, 1) x n], all row-major fp32.
 * path 0 = fp32 CUDA-core kernel, 1 = tcgen05 3xTF32 kernel, 2 = same with the split (two-plane) store.
 * Rows whose source row falls outside [0, rows) are unspecified (clamped by path 0, zero by 1/2).
 * iters > 0 additionally times that many launches (CUDA events) and returns the mean in *ms. */
int rs_debug_gemm(int device, const float *src, int rows, int k, const int *offsets, int n_offsets, int stride,
                  const float *w, int n, const float *bias, int relu, int path, int iters, float *out, float *ms,
                  char *err, size_t errlen);
/* Test hook for the host half of the n-best tail: n-best of a caller-provided state-level lattice (arcs
 * src/dst/olabel/graph/acoustic; dst == -1 marks a final weight; node 0 is the start; node ids ascend with time).
 * Fills up to n hypotheses: words into word_ids (capacity max_words in total) delimited by word_offset[n + 1],
 * costs into cost[2 * n] (graph, acoustic).  Returns the number of hypotheses, < 0 on error. */
int rs_debug_lattice_nbest(const int32_t *src, const int32_t *dst, const int32_t *olabel, const float *graph,
                           const float *acoustic, int32_t n_arcs, int32_t n_nodes, int32_t n, float acoustic_scale,
                           int32_t *word_offset, int32_t *word_ids, int32_t max_words, float *cost);
/* Test hook for the Kaldi object reader (host only): a Matrix<float> file as Matrix::Read accepts it -- binary FM / DM,
 * CompressedMatrix CM / CM2 / CM3 (kaldi/src/matrix/kaldi-matrix.cc:1475-1513, compressed-matrix.cc:565-660), or text.
 * Call with dst == NULL to query the shape, then with *rows / *cols set to it. */
int rs_debug_read_matrix(const char *path, float *dst, int32_t *rows, int32_t *cols, char *err, size_t errlen);
/* Test hook for the strict-order host decoder (csrc/strict_decode.cc; host only, no GPU): what latgen-faster-mapped
 * (kaldi/src/bin/latgen-faster-mapped.cc) | lattice-to-nbest --n=nbest --acoustic-scale | nbest-to-linear print for
 * one log-likelihood matrix [n_frames x num_pdfs], with the reference's order-dependent pruning reproduced
 * (lattice-faster-decoder.cc:780-787, util/hash-list-inl.h:156-194).  tid2pdf[0 .. n_tids) maps HCLG input labels to
 * pdfs.  Outputs as rs_debug_lattice_nbest; lattice_size (may be NULL) = {states, arcs incl. final weights} of the
 * pruned state-level lattice when nbest > 1 or acoustic_scale != 1.  Returns the number of hypotheses, < 0 on error. */
int rs_debug_strict_decode(const char *hclg_fst, const int32_t *tid2pdf, int32_t n_tids, const float *loglikes,
                           int32_t n_frames, int32_t num_pdfs, const rs_decoder_opts *opts, int32_t nbest, float acoustic_scale,
                           int32_t *word_offset, int32_t *word_ids, int32_t max_words, float *cost, int32_t *lattice_size,
                           char *err, size_t errlen);
/* Fuzzy matcher in process (host only, no GPU): replaces the seven-process OpenFst pipeline of
 * rhasspy_speech/transcribe_util.py:46-60 (fstcompile | fstcompose - G.fuzzy.fst | fstshortestpath | fstrmepsilon |
 * fsttopsort | fstproject --project_type=output | fstprint).  rs_fuzzy_load reads lang_dir/G.fuzzy.fst (OpenFst vector
 * or const FST over StdArc, embedded symbol tables skipped) and lang_dir/words.txt.  rs_fuzzy_match takes the n-best
 * hypotheses (word ids of hypothesis k = word_ids[hyp_offset[k] .. hyp_offset[k+1]), rank k penalised 0.1 * k per
 * word, transcribe_util.py:28-40) and returns the output word ids of the cheapest match and the cost the reference
 * computes from fstprint's arc lines.  Returns 0 = match, 1 = no path (the reference's `None`), < 0 error. */
typedef struct rs_fuzzy rs_fuzzy;
rs_fuzzy *rs_fuzzy_load(const char *g_fuzzy_fst, const char *words_txt, char *err, size_t errlen);
void rs_fuzzy_free(rs_fuzzy *f);
int rs_fuzzy_match(const rs_fuzzy *f, const int32_t *word_ids, const int32_t *hyp_offset, int32_t n_hyp, int32_t *out_ids,
                   int32_t max_out, int32_t *n_out, float *cost, char *err, size_t errlen);
const char *rs_fuzzy_word(const rs_fuzzy *f, int32_t id);
/* Text description of the compiled acoustic-model plan (one line per launch). */
const char *rs_model_plan(const rs_model *m);

/* Host-only validation of the artefacts (no GPU needed): parse final.mdl + online.conf and every
 * file they name / HCLG.fst + words.txt exactly as the loaders above do, and report what was found.
 * rs_model_check writes a summary + the compiled plan into `out`; rs_graph_check fills
 * counts[6] = {states, emitting arcs, epsilon-input arcs, start state, final states, word symbols}. */
int rs_model_check(const char *final_mdl, const char *online_conf, char *out, size_t outlen, char *err, size_t errlen);
int rs_graph_check(const char *hclg_fst, const char *words_txt, int64_t *counts, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* RS_B200_H_ */
