// Strict-order host decoder: the search of LatticeFasterDecoder with the reference's token ORDER reproduced, for
// the utterances whose device decode was order-sensitive (rs_result.status bit 4, see decode.cu "safe frame" rules).
#pragma once
#include <cstdint>
#include <vector>

#include "engine.h"
#include "model.h"

namespace rs {

struct StrictOptions {
  float beam = 24.f, beam_delta = 0.5f, lattice_beam = 8.f, hash_ratio = 2.f;
  int max_active = 7000, min_active = 200;
  int max_words = 256;
};

struct StrictResult {
  bool decoded = false;       // false: no surviving token (the reference prints nothing for the utterance)
  bool word_overflow = false;
  std::vector<int> words;     // olabels of the best path
  float graph = 0.f, acoustic = 0.f;
  // want_lattice: the pruned state-level lattice in the layout lattice_prune_kernel produces (node 0 = start,
  // node ids ascend with time, dst == -1 = final weight)
  std::vector<LatticeArc> lattice;
  int n_nodes = 0;
  uint64_t tokens_expanded = 0, arcs_visited = 0, tokens_created = 0;
};

// The arcs of a graph as 12-byte records in CSR order (Graph keeps one array per field; the search touches next state,
// pdf and weight of every arc it visits and nothing else): built once per (graph, model) pair.
struct StrictArcs {
  struct Arc {
    int32_t next, pdf;  // pdf = -1 on epsilon arcs
    float weight;
  };
  std::vector<Arc> emitting, epsilon;
  std::vector<uint64_t> has_epsilon;  // one bit per state: it has epsilon-input arcs (16 KB for 127 k states)
};
void BuildStrictArcs(const Graph &g, const int32_t *e_pdf, StrictArcs *out);

// e_pdf[a] = pdf of emitting arc a (transition-id -> pdf applied); loglikes [n_frames x ld]
void StrictDecode(const Graph &g, const int32_t *e_pdf, const StrictArcs &arcs, const float *loglikes, int ld, int n_frames,
                  const StrictOptions &opt, bool want_lattice, StrictResult *out);
// the same, building the arc records for this one call
void StrictDecode(const Graph &g, const int32_t *e_pdf, const float *loglikes, int ld, int n_frames,
                  const StrictOptions &opt, bool want_lattice, StrictResult *out);

}  // namespace rs
