// Host side of the n-best tail (a22): n cheapest distinct word sequences of a pruned state-level lattice.
#pragma once
#include <vector>

#include "engine.h"

namespace rs {

struct NbestHyp {
  std::vector<int> words;
  float graph = 0.f, acoustic = 0.f;  // weight of the word sequence's best path
};

// `arcs` as lattice_prune_kernel writes them: node 0 is the start, node ids ascend with time, dst == -1 marks
// a final weight.  acoustic_scale is lattice-to-nbest's --acoustic-scale: it ranks, the costs stay unscaled.
void LatticeNbest(const LatticeArc *arcs, int n_arcs, int n_nodes, int n, float acoustic_scale,
                  std::vector<NbestHyp> *out, int max_expansions = 100000);

}  // namespace rs
