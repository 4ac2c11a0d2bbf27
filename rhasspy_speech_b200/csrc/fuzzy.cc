// Row f2: rhasspy-speech's fuzzy matcher ("Handling Out of Vocabulary", reference README.md:42-46) in process.
//
// The reference (rhasspy_speech/transcribe_util.py:11-88) writes the n-best hypotheses as a text FST -- one linear
// chain of word ids per hypothesis from state 0, every arc of hypothesis k weighted 0.1 * k (:28-40) -- and pipes it
// through  fstcompile | fstcompose - G.fuzzy.fst | fstshortestpath | fstrmepsilon | fsttopsort |
// fstproject --project_type=output | fstprint --osymbols=words.txt  (:46-60), then reads the words and sums the
// printed arc weights (:62-83).  G.fuzzy.fst is the sentence grammar with, on every state, an <eps>:<eps>/0 loop and a
// word:<eps>/1.0 loop per vocabulary word (rhasspy_speech/kaldi.py:360-389): input words can be skipped at cost 1.
//
// Here: one label-correcting shortest-path search over the product (hypothesis position, grammar state), built on
// the fly -- the composition is never materialised.  What the seven processes print is reproduced from the best path:
//   * fstrmepsilon removes the arcs whose input AND output are <eps> and adds their weight to the next remaining
//     arc -- or to the final weight when none follows; fstprint writes the final weight on a line the reference
//     does not read (:70-71), so the reported cost is the sum over the path up to its last arc that is not
//     <eps>:<eps>, final weights excluded;
//   * after fstproject the words are the output labels; the reference drops <eps> (:80-81).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rs_b200.h"
#include "model.h"

namespace rs {

struct FuzzyImpl {
  Graph g;  // G.fuzzy.fst: emitting arcs = arcs with an input word, "epsilon" arcs = <eps> input
  // every state carries one word:<eps> loop per vocabulary word, so a state's arcs are looked up by input label:
  // by_label[e_begin[s] .. e_begin[s+1]) = that state's emitting arcs sorted by (ilabel, original position)
  std::vector<uint32_t> by_label;
  bool non_negative = true;  // all arc and final weights >= 0: Dijkstra order with early termination is valid
  void Index() {
    by_label.resize(g.e_ilabel.size());
    for (size_t i = 0; i < by_label.size(); i++) by_label[i] = (uint32_t)i;
    for (int s = 0; s < g.num_states; s++)
      std::stable_sort(by_label.begin() + g.e_begin[s], by_label.begin() + g.e_begin[s + 1],
                       [&](uint32_t a, uint32_t b) { return g.e_ilabel[a] < g.e_ilabel[b]; });
    for (float w : g.e_weight) non_negative &= !(w < 0.f);
    for (float w : g.p_weight) non_negative &= !(w < 0.f);
    for (float w : g.final_cost) non_negative &= !(w < 0.f);
  }
};

namespace {

struct FzNode {
  int p = 0, s = 0;   // hypothesis position, grammar state
  float cost = std::numeric_limits<float>::infinity();
  int prev = -1;      // predecessor node
  int olabel = 0;
  float weight = 0.f;
  bool both_eps = false;  // the arc into this node was <eps>:<eps>
  bool queued = false;
};

}  // namespace

// hyps: word ids of hypothesis k are ids[offset[k] .. offset[k+1]).  Returns false when no path exists.
static bool FuzzyMatch(const FuzzyImpl &f, const int32_t *ids, const int32_t *offset, int n_hyp, std::vector<int> *words,
                       float *cost) {
  const Graph &g = f.g;
  const int S = g.num_states;
  if (S <= 0 || g.start < 0 || n_hyp <= 0) return false;
  // input positions: position 0 is the shared start state, hypothesis k owns positions base[k]+1 .. base[k]+len
  // (state numbering of hassil_fst.Fst.next_edge; only the structure matters)
  std::vector<char> pos_final;
  // position p > 0 of a chain has exactly one outgoing arc (or none at the chain's end); the start has one per chain
  struct Out { int word; float w; int to; };
  std::vector<std::vector<Out>> out(1);
  pos_final.assign(1, 0);
  double penalty = 0.0;  // Python float, += 0.1 per hypothesis (:39-40), converted to float by fstcompile
  for (int k = 0; k < n_hyp; k++) {
    int cur = 0;
    for (int i = offset[k]; i < offset[k + 1]; i++) {
      out.emplace_back();
      pos_final.push_back(0);
      const int to = (int)out.size() - 1;
      out[cur].push_back(Out{ids[i], (float)penalty, to});
      cur = to;
    }
    pos_final[cur] = 1;
    penalty += 0.1;
  }
  const int P = (int)out.size();
  // product states are created when first reached (the full product P x S would be huge for a large grammar)
  std::vector<FzNode> node;
  std::unordered_map<long long, int> index;
  auto id = [&](int p, int s) {
    const long long key = (long long)p * S + s;
    auto it = index.find(key);
    if (it != index.end()) return it->second;
    node.emplace_back();
    node.back().p = p;
    node.back().s = s;
    index.emplace(key, (int)node.size() - 1);
    return (int)node.size() - 1;
  };
  // Non-negative weights (the normal case: -log probabilities, penalties): Dijkstra order, and the search stops as
  // soon as no open node can beat the best complete match.  Otherwise: label-correcting with a FIFO queue.
  const bool dijkstra = f.non_negative;
  std::deque<int> queue;
  typedef std::pair<float, std::pair<long, int>> HeapItem;  // (cost, (sequence number, node)): ties in creation order
  std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>> heap;
  long seq = 0;
  auto relax = [&](int from, int to, float w, int olabel, bool both_eps) {
    const float c = node[from].cost + w;
    if (c < node[to].cost) {
      node[to].cost = c;
      node[to].prev = from;
      node[to].olabel = olabel;
      node[to].weight = w;
      node[to].both_eps = both_eps;
      if (dijkstra) {
        heap.push(HeapItem(c, std::make_pair(seq++, to)));
      } else if (!node[to].queued) {
        node[to].queued = true;
        queue.push_back(to);
      }
    }
  };
  const int start = id(0, (int)g.start);
  node[start].cost = 0.f;
  node[start].queued = true;
  if (dijkstra) heap.push(HeapItem(0.f, std::make_pair(seq++, start))); else queue.push_back(start);
  long relaxations = 0;
  const long limit = 64L * 1000 * 1000;
  float best_complete = std::numeric_limits<float>::infinity();
  while (dijkstra ? !heap.empty() : !queue.empty()) {
    int n;
    if (dijkstra) {
      const HeapItem top = heap.top();
      heap.pop();
      n = top.second.second;
      if (top.first > node[n].cost) continue;      // a stale entry
      if (top.first >= best_complete) break;       // nothing left that could be cheaper
      if (pos_final[node[n].p] && !std::isinf(g.final_cost[node[n].s]))
        best_complete = std::min(best_complete, node[n].cost + g.final_cost[node[n].s]);
    } else {
      n = queue.front();
      queue.pop_front();
      node[n].queued = false;
    }
    const int p = node[n].p, s = node[n].s;
    // grammar arcs with <eps> input: the hypothesis does not advance
    for (uint32_t a = g.p_begin[s]; a < g.p_begin[s + 1]; a++) {
      if (g.p_next[a] == s && g.p_olabel[a] == 0 && !(g.p_weight[a] < 0.f)) continue;  // the <eps>:<eps> self loop
      relax(n, id(p, g.p_next[a]), g.p_weight[a], g.p_olabel[a], g.p_olabel[a] == 0);
    }
    // a hypothesis word matched with a grammar arc of the same input label
    for (const Out &o : out[p]) {
      const uint32_t *lo = f.by_label.data() + g.e_begin[s], *hi = f.by_label.data() + g.e_begin[s + 1];
      const uint32_t *it = std::lower_bound(lo, hi, o.word, [&](uint32_t a, int w) { return g.e_ilabel[a] < w; });
      for (; it < hi && g.e_ilabel[*it] == o.word; ++it)
        relax(n, id(o.to, g.e_next[*it]), o.w + g.e_weight[*it], g.e_olabel[*it], false);
    }
    if (++relaxations > limit) return false;
  }
  // best final node: hypothesis at its end, grammar state final
  int best = -1;
  float best_cost = std::numeric_limits<float>::infinity();
  for (int n = 0; n < (int)node.size(); n++) {
    const float fc = g.final_cost[node[n].s];
    if (!pos_final[node[n].p] || std::isinf(fc) || std::isinf(node[n].cost)) continue;
    const float c = node[n].cost + fc;
    if (c < best_cost) {
      best_cost = c;
      best = n;
    }
  }
  (void)P;
  if (best < 0) return false;
  std::vector<int> path;
  for (int n = best; n != start && n >= 0; n = node[n].prev) path.push_back(n);
  words->clear();
  // path is in reverse order: skip the trailing <eps>:<eps> arcs, then sum every weight before them
  size_t first_counted = 0;
  while (first_counted < path.size() && node[path[first_counted]].both_eps) first_counted++;
  double sum = 0.0;
  for (size_t i = path.size(); i-- > first_counted;) {
    sum += node[path[i]].weight;
    if (node[path[i]].olabel != 0) words->push_back(node[path[i]].olabel);
  }
  *cost = (float)sum;
  return true;
}

}  // namespace rs

using namespace rs;

extern "C" {

static void FzErr(char *err, size_t errlen, const std::string &m) {
  if (err && errlen) {
    size_t n = std::min(errlen - 1, m.size());
    memcpy(err, m.data(), n);
    err[n] = 0;
  }
}

rs_fuzzy *rs_fuzzy_load(const char *g_fuzzy_fst, const char *words_txt, char *err, size_t errlen) {
  try {
    if (!g_fuzzy_fst) RS_FAIL("rs_fuzzy_load: path required");
    std::unique_ptr<FuzzyImpl> f(new FuzzyImpl());
    LoadGraph(g_fuzzy_fst, words_txt ? words_txt : "", &f->g);
    f->Index();
    return reinterpret_cast<rs_fuzzy *>(f.release());
  } catch (const std::exception &e) {
    FzErr(err, errlen, e.what());
    return nullptr;
  }
}

void rs_fuzzy_free(rs_fuzzy *f) { delete reinterpret_cast<FuzzyImpl *>(f); }

int rs_fuzzy_match(const rs_fuzzy *f_, const int32_t *word_ids, const int32_t *hyp_offset, int32_t n_hyp, int32_t *out_ids,
                   int32_t max_out, int32_t *n_out, float *cost, char *err, size_t errlen) {
  try {
    const FuzzyImpl *f = reinterpret_cast<const FuzzyImpl *>(f_);
    if (!f || !hyp_offset || !n_out || !cost || (n_hyp > 0 && hyp_offset[n_hyp] > 0 && !word_ids))
      RS_FAIL("rs_fuzzy_match: bad argument");
    std::vector<int> words;
    float c = 0.f;
    *n_out = 0;
    *cost = 0.f;
    if (!FuzzyMatch(*f, word_ids, hyp_offset, n_hyp, &words, &c)) return 1;  // no path: the pipeline prints nothing
    if ((int)words.size() > max_out) RS_FAIL("rs_fuzzy_match: output capacity too small");
    for (size_t i = 0; i < words.size(); i++) out_ids[i] = words[i];
    *n_out = (int32_t)words.size();
    *cost = c;
    return 0;
  } catch (const std::exception &e) {
    FzErr(err, errlen, e.what());
    return -1;
  }
}

const char *rs_fuzzy_word(const rs_fuzzy *f_, int32_t id) {
  const FuzzyImpl *f = reinterpret_cast<const FuzzyImpl *>(f_);
  if (!f || id < 0 || id >= (int)f->g.words.size() || f->g.words[id].empty()) return nullptr;
  return f->g.words[id].c_str();
}

}  // extern "C"
