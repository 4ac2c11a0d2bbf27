// a22: n-best word sequences of a pruned state-level lattice (host side of the n-best tail).
//
// The reference turns the decoder's raw lattice into n hypotheses in three steps
//   DeterminizeLatticePhonePrunedWrapper (kaldi/src/lat/determinize-lattice-pruned.cc:1488, called from
//     online-nnet3-decoding.cc:66-79): one path per distinct word sequence, carrying the (graph, acoustic)
//     weight of that sequence's best alignment under LatticeWeight's order (sum, then graph cost);
//   lattice-to-nbest (kaldi/src/latbin/lattice-to-nbest.cc:84-113): acoustic costs times --acoustic-scale,
//     fst::ShortestPath(n) over the determinised lattice, keys utt-1 .. utt-n;
//   nbest-to-linear (nbest-to-linear.cc:70-92): the word ids and the two costs of every path.
// The product of those steps is the n cheapest DISTINCT word sequences, each with the weight of its best path.
// Determinising the whole lattice to read off n <= a handful of paths is the expensive way round; here the
// word-prefix tree is explored best-first instead.  A tree node is the set of lattice states reachable from
// the start with exactly that word prefix (the subset the determiniser would build for it), its priority the
// cheapest completion of any of its states, known exactly from one backward pass -- so nodes leave the queue in
// the order of their best word sequence and only the subsets along the n winners are ever built.
#include "nbest.h"

#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <map>
#include <queue>

namespace rs {

namespace {

struct W {  // LatticeWeight: compared on the sum, ties on the graph cost (lattice-weight.h Compare)
  double g, a;
  double sum() const { return g + a; }
};
inline bool Better(const W &x, const W &y) {
  const double sx = x.sum(), sy = y.sum();
  if (sx < sy) return true;
  if (sx > sy) return false;
  return x.g < y.g;
}

struct Elem {
  int node;
  W w;
};

struct TreeNode {
  int parent = -1;
  int word = 0;
  bool complete = false;  // a finished hypothesis (final weight included)
  W w{0, 0};              // complete: weight of the hypothesis
  std::vector<Elem> seeds;
};

}  // namespace

void LatticeNbest(const LatticeArc *arcs, int n_arcs, int n_nodes, int n, float acoustic_scale,
                  std::vector<NbestHyp> *out, int max_expansions) {
  out->clear();
  if (n_nodes <= 0 || n_arcs <= 0 || n <= 0) return;
  const double kInf = std::numeric_limits<double>::infinity();
  // CSR by source node; final weights apart
  std::vector<int> begin(n_nodes + 1, 0);
  std::vector<double> final_g(n_nodes, kInf);
  for (int i = 0; i < n_arcs; i++) {
    const LatticeArc &a = arcs[i];
    if (a.src < 0 || a.src >= n_nodes) continue;
    if (a.dst < 0)
      final_g[a.src] = std::min(final_g[a.src], (double)a.graph);
    else
      begin[a.src + 1]++;
  }
  for (int i = 0; i < n_nodes; i++) begin[i + 1] += begin[i];
  std::vector<int> order(begin[n_nodes]), fill(begin.begin(), begin.end() - 1);
  for (int i = 0; i < n_arcs; i++) {
    const LatticeArc &a = arcs[i];
    if (a.src < 0 || a.src >= n_nodes || a.dst < 0 || a.dst >= n_nodes) continue;
    order[fill[a.src]++] = i;
  }
  // backward pass: cheapest completion (sum of both costs) of every node.  Node ids are time-major, so one
  // descending sweep settles everything but the epsilon links inside one time; repeat until stable.
  std::vector<double> beta(final_g);
  for (int it = 0; it <= n_nodes; it++) {
    bool changed = false;
    for (int s = n_nodes - 1; s >= 0; s--) {
      double b = beta[s];
      for (int k = begin[s]; k < begin[s + 1]; k++) {
        const LatticeArc &a = arcs[order[k]];
        const double c = (double)a.graph + (double)a.acoustic + beta[a.dst];
        if (c < b) b = c;
      }
      if (b < beta[s]) {
        beta[s] = b;
        changed = true;
      }
    }
    if (!changed) break;
  }
  if (beta[0] == kInf) return;

  std::vector<TreeNode> tree;
  typedef std::pair<double, int> QE;  // (priority, tree node); equal priorities leave in creation order
  std::priority_queue<QE, std::vector<QE>, std::greater<QE>> queue;
  tree.emplace_back();
  tree[0].seeds.push_back(Elem{0, W{0, 0}});
  queue.push(QE(beta[0], 0));

  // with a re-ranking scale the order of the hypotheses can change, so enumerate every distinct word
  // sequence of the (beam-pruned) lattice up to the expansion limit and rank afterwards
  const bool rerank = acoustic_scale != 1.0f;
  const size_t want = rerank ? (size_t)std::max(n * 64, 512) : (size_t)n;

  std::vector<int> slot(n_nodes, -1);  // closure: index into `closure` of each touched node
  std::vector<Elem> closure;
  std::vector<NbestHyp> found;
  int expansions = 0;
  while (!queue.empty() && found.size() < want) {
    const int id = queue.top().second;
    queue.pop();
    if (tree[id].complete) {
      NbestHyp h;
      h.graph = (float)tree[id].w.g;
      h.acoustic = (float)tree[id].w.a;
      for (int p = tree[id].parent; p > 0; p = tree[p].parent) h.words.push_back(tree[p].word);
      std::reverse(h.words.begin(), h.words.end());
      found.push_back(std::move(h));
      continue;
    }
    if (++expansions > max_expansions) break;
    // ---- closure of the seeds over links without a word label, keeping the best weight per lattice state
    closure.clear();
    std::priority_queue<int, std::vector<int>, std::greater<int>> work;  // lattice states, earliest first
    auto relax = [&](int node, const W &w) {
      int &sl = slot[node];
      if (sl < 0) {
        sl = (int)closure.size();
        closure.push_back(Elem{node, w});
        work.push(node);
      } else if (Better(w, closure[sl].w)) {
        closure[sl].w = w;
        work.push(node);
      }
    };
    for (const Elem &e : tree[id].seeds) relax(e.node, e.w);
    std::vector<Elem>().swap(tree[id].seeds);
    while (!work.empty()) {
      const int s = work.top();
      work.pop();
      while (!work.empty() && work.top() == s) work.pop();
      const W w = closure[slot[s]].w;
      for (int k = begin[s]; k < begin[s + 1]; k++) {
        const LatticeArc &a = arcs[order[k]];
        if (a.olabel != 0) continue;
        relax(a.dst, W{w.g + a.graph, w.a + a.acoustic});
      }
    }
    // ---- this prefix as a whole hypothesis, and its extensions by one word
    bool has_final = false;
    W best_final{kInf, 0};
    std::map<int, int> child;  // word -> tree node
    for (const Elem &e : closure) {
      if (final_g[e.node] != kInf) {
        const W f{e.w.g + final_g[e.node], e.w.a};
        if (!has_final || Better(f, best_final)) best_final = f;
        has_final = true;
      }
      for (int k = begin[e.node]; k < begin[e.node + 1]; k++) {
        const LatticeArc &a = arcs[order[k]];
        if (a.olabel == 0 || beta[a.dst] == kInf) continue;
        auto it = child.find(a.olabel);
        if (it == child.end()) {
          it = child.emplace(a.olabel, (int)tree.size()).first;
          tree.emplace_back();
          tree.back().parent = id;
          tree.back().word = a.olabel;
        }
        tree[it->second].seeds.push_back(Elem{a.dst, W{e.w.g + a.graph, e.w.a + a.acoustic}});
      }
    }
    for (const Elem &e : closure) slot[e.node] = -1;
    if (has_final) {
      tree.emplace_back();
      tree.back().parent = id;  // the word chain of a hypothesis is read from the prefix node upwards
      tree.back().complete = true;
      tree.back().w = best_final;
      queue.push(QE(best_final.sum(), (int)tree.size() - 1));
    }
    for (const auto &kv : child) {
      double prio = kInf;
      for (const Elem &e : tree[kv.second].seeds) prio = std::min(prio, e.w.sum() + beta[e.node]);
      queue.push(QE(prio, kv.second));
    }
  }
  if (rerank) {
    const double s = acoustic_scale;
    std::stable_sort(found.begin(), found.end(), [s](const NbestHyp &x, const NbestHyp &y) {
      return (double)x.graph + s * x.acoustic < (double)y.graph + s * y.acoustic;
    });
  }
  if ((int)found.size() > n) found.resize(n);
  *out = std::move(found);
}

}  // namespace rs
