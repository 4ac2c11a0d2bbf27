// Strict-order host decoder ("hard parts"): LatticeFasterDecoder's search with the ORDER in which the
// reference visits tokens reproduced, so that the order-dependent parts of its pruning come out identically:
//   * the transient next_cutoff of ProcessEmitting (kaldi/src/decoder/lattice-faster-decoder.cc:780-787), which admits
//     tokens beyond the frame's final cutoff depending on how early they are reached ("extras");
//   * first-come-wins on equal costs in FindOrAddToken (:252-293);
//   * the LIFO queue of ProcessNonemitting (:846-884);
//   * the iteration order of the token hash (kaldi/src/util/hash-list-inl.h:156-194): buckets in the order they were
//     first occupied, elements of one bucket in insertion order, bucket = state % hash_size, hash_size grown by
//     PossiblyResizeHash (:219-225).
// The device decoder (decode.cu) keeps exactly the tokens inside each frame's final cutoff; it proves per frame that
// the extras cannot matter ("safe frame" rules there) and flags the utterance otherwise.  Flagged utterances are
// decoded again here from the log-likelihoods that are still resident on the device -- a second, exact opinion on a
// rare case, not a fallback for the search as a whole: an unflagged utterance never reaches this file.
// Float expressions follow the reference operand by operand (this file is compiled with -ffp-contract=off).
#include "strict_decode.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>

namespace rs {

namespace {

constexpr float kInf = std::numeric_limits<float>::infinity();

struct Tok {
  float tot, extra;
  int state;
  int back;      // token that set the current cost (-1: start token)
  int back_arc;  // arc it came through: emitting arc id, or num_earcs + epsilon arc id; -1: none
  int links;     // head of the forward-link list (lattice mode)
};
struct Link {
  int next_tok, arc;
  float graph, acoustic;
  int next;
};

// state -> token map whose iteration order is the reference HashList's
class OrderedStateMap {
 public:
  struct Elem {
    int state, tok, tail;
    float tot;  // the token's current cost, kept next to the key: a hit on an existing token does not touch the token array
  };
  struct Item {  // one entry of a released list: (state, token, cost)
    int first, second;
    float tot;
  };
  void SetSize(size_t n) {
    hash_size_ = n;
    // state % hash_size without a division per look-up (Lemire's fastmod: exact for 32-bit operands)
    mod_m_ = n ? ~uint64_t(0) / n + 1 : 0;
    if (n > buckets_.size()) buckets_.resize(n);
  }
  size_t Size() const { return hash_size_; }
  int Head() const { return head_; }
  size_t Index(int state) const {  // state % hash_size
    return hash_size_ <= 0xffffffffu ? (size_t)(((unsigned __int128)(mod_m_ * (uint32_t)state) * hash_size_) >> 64)
                                     : static_cast<size_t>(state) % hash_size_;
  }
  const Elem &At(int e) const { return pool_[e]; }
  void SetTok(int e, int tok, float tot) {
    pool_[e].tok = tok;
    pool_[e].tot = tot;
  }
  void SetTot(int e, float tot) { pool_[e].tot = tot; }
  // element of `state`, created (tok = -1) behind the last element of its bucket when absent
  int Insert(int state) {
    const size_t idx = Index(state);
    Bucket &b = buckets_[idx];
    const bool occupied = b.epoch == epoch_;
    if (occupied) {
      // the bucket's elements are a run of the list: from its first element (fixed when the bucket was opened: later
      // elements of the previous bucket are spliced in before it) to the one behind its last
      const int stop = pool_[b.last].tail;
      for (int e = b.first; e != stop; e = pool_[e].tail)
        if (pool_[e].state == state) return e;
    }
    const int e = (int)pool_.size();
    pool_.push_back(Elem{state, -1, -1, kInf});
    if (!occupied) {  // the bucket joins the end of the bucket chain, its element the end of the list
      if (tail_bucket_ < 0) head_ = e; else pool_[buckets_[tail_bucket_].last].tail = e;
      b.last = e;
      b.first = e;
      b.epoch = epoch_;
      tail_bucket_ = (int)idx;
    } else {
      pool_[e].tail = pool_[b.last].tail;
      pool_[b.last].tail = e;
      b.last = e;
    }
    return e;
  }
  // the (state, token) pairs in list order; the map is left empty
  void Release(std::vector<Item> *out) {
    out->clear();
    for (int e = head_; e >= 0; e = pool_[e].tail) out->push_back(Item{pool_[e].state, pool_[e].tok, pool_[e].tot});
    pool_.clear();
    head_ = -1;
    tail_bucket_ = -1;
    if (++epoch_ == 0) {  // 2^32 releases: start over with clean buckets
      for (auto &b : buckets_) b.epoch = 0;
      epoch_ = 1;
    }
  }

 private:
  struct Bucket {  // 12 bytes: the table of a 7000-token frame (14 k buckets) stays in the L2 of a core
    int first = -1, last = -1;
    uint32_t epoch = 0;
  };
  std::vector<Elem> pool_;
  std::vector<Bucket> buckets_;
  size_t hash_size_ = 0;
  uint64_t mod_m_ = 0;
  int head_ = -1;
  int tail_bucket_ = -1;
  uint32_t epoch_ = 1;
};

inline bool ApproxEq(float a, float b, float tol) {  // kaldi/src/base/kaldi-math.h:265-273
  if (a == b) return true;
  const float diff = std::fabs(a - b);
  if (diff == kInf || diff != diff) return false;
  return diff <= tol * (std::fabs(a) + std::fabs(b));
}

class Search {
 public:
  Search(const Graph &g, const int32_t *e_pdf, const StrictArcs &arcs, const float *ll, int ld, const StrictOptions &o, bool lattice)
      : g_(g), e_pdf_(e_pdf), ea_(arcs.emitting.data()), pa_(arcs.epsilon.data()), eps_bits_(arcs.has_epsilon.data()), ll_(ll),
        ld_(ld), o_(o), lattice_(lattice),
        NE_((int)g.e_next.size()) {}

  void Run(int n_frames, StrictResult *out) {
    map_.SetSize(1000);  // the constructor's toks_.SetSize(1000)
    // InitDecoding :56-73
    // (room for a max-active frontier on every frame: the token array of a 4 s utterance on an LM-sized graph reaches
    //  ~600 k entries and was copied five times over while it grew)
    toks_.reserve((size_t)(n_frames + 1) * (size_t)std::min(std::max(o_.max_active, 256), 6000));
    if (lattice_) links_.reserve(toks_.capacity() * 2);  // one link per admitted arc, ~2.5 per token on an LM-sized graph
    frame_begin_.push_back(0);
    {
      const int e = map_.Insert((int)g_.start);
      toks_.push_back(Tok{0.f, 0.f, (int)g_.start, -1, -1, -1});
      map_.SetTok(e, 0, 0.f);
    }
    ProcessNonemitting(o_.beam);
    for (int f = 0; f < n_frames; f++) {
      frame_begin_.push_back((int)toks_.size());
      const float cutoff = ProcessEmitting(f);
      ProcessNonemitting(cutoff);
    }
    frame_begin_.push_back((int)toks_.size());  // frame_begin_[t] .. frame_begin_[t + 1]: tokens of time t (0 .. n_frames)
    out->tokens_expanded = expanded_;
    out->arcs_visited = arcs_;
    out->tokens_created = toks_.size();
    BestPath(n_frames, out);
    if (debug_) fprintf(stderr, "strict: %d frames, %d unsafe\n", n_frames, unsafe_frames_);
    if (lattice_ && out->decoded) Lattice(n_frames, out);
  }

 private:
  bool HasEps(int s) const { return (eps_bits_[(size_t)s >> 6] >> (s & 63)) & 1u; }

  // FindOrAddToken :252-293
  int FindOrAdd(int state, float tot, int back, int back_arc, bool *changed) {
    const int e = map_.Insert(state);
    const int t = map_.At(e).tok;
    if (t < 0) {
      toks_.push_back(Tok{tot, 0.f, state, back, back_arc, -1});
      map_.SetTok(e, (int)toks_.size() - 1, tot);
      if (changed) *changed = true;
    } else if (map_.At(e).tot > tot) {
      map_.SetTot(e, tot);
      toks_[t].tot = tot;
      toks_[t].back = back;
      toks_[t].back_arc = back_arc;
      if (changed) *changed = true;
    } else if (changed) {
      *changed = false;
    }
    return e;
  }
  void AddLink(int tok, int next_tok, int arc, float graph, float acoustic) {
    links_.push_back(Link{next_tok, arc, graph, acoustic, toks_[tok].links});
    toks_[tok].links = (int)links_.size() - 1;
  }

  // GetCutoff :644-711
  float GetCutoff(const std::vector<OrderedStateMap::Item> &list, float *adaptive_beam, int *best) {
    float best_w = kInf;
    *best = -1;
    tmp_.clear();
    for (size_t i = 0; i < list.size(); i++) {
      const float w = list[i].tot;
      tmp_.push_back(w);
      if (w < best_w) {
        best_w = w;
        *best = (int)i;
      }
    }
    if (o_.max_active == std::numeric_limits<int>::max() && o_.min_active == 0) {
      *adaptive_beam = o_.beam;
      return best_w + o_.beam;
    }
    const float beam_cutoff = best_w + o_.beam;
    float min_active_cutoff = kInf, max_active_cutoff = kInf;
    // The reference runs nth_element for both limits on every frame and then only COMPARES the two order statistics
    // with the beam cutoff.  With s = the sorted costs: s[max_active] < beam_cutoff  <=>  more than max_active costs lie
    // below the beam cutoff, and s[min_active] > beam_cutoff  <=>  at most min_active costs lie at or below it.  Two
    // counts (one branch-free pass) decide both; a selection only runs when its limit binds -- same values, and the
    // selection was an eighth of the search time on an LM-sized graph.
    size_t below = 0, at_or_below = 0;
    for (const float w : tmp_) {
      below += w < beam_cutoff;
      at_or_below += w <= beam_cutoff;
    }
    const bool over_max = tmp_.size() > (size_t)o_.max_active;
    bool selected_max = false;
    if (over_max && below > (size_t)o_.max_active) {
      std::nth_element(tmp_.begin(), tmp_.begin() + o_.max_active, tmp_.end());
      max_active_cutoff = tmp_[o_.max_active];
      selected_max = true;
    }
    if (max_active_cutoff < beam_cutoff) {
      *adaptive_beam = max_active_cutoff - best_w + o_.beam_delta;
      return max_active_cutoff;
    }
    if (tmp_.size() > (size_t)o_.min_active) {
      if (o_.min_active == 0) {
        min_active_cutoff = best_w;
      } else if (at_or_below <= (size_t)o_.min_active) {
        // (after a max-active selection the min_active smallest costs sit in front of position max_active)
        std::nth_element(tmp_.begin(), tmp_.begin() + o_.min_active, selected_max ? tmp_.begin() + o_.max_active : tmp_.end());
        min_active_cutoff = tmp_[o_.min_active];
      } else {
        min_active_cutoff = -kInf;  // s[min_active] <= beam_cutoff: the beam decides below
      }
    }
    if (min_active_cutoff > beam_cutoff) {
      *adaptive_beam = min_active_cutoff - best_w + o_.beam_delta;
      return min_active_cutoff;
    }
    *adaptive_beam = o_.beam;
    return beam_cutoff;
  }

  // ProcessEmitting :714-804
  float ProcessEmitting(int frame) {
    map_.Release(&list_);
    float adaptive_beam;
    int best;
    if (debug_ && frame > 0) SafeFrameProbe(frame);
    const float cur_cutoff = GetCutoff(list_, &adaptive_beam, &best);
    if (debug_) fprintf(stderr, "strict frame %d tokens %zu adaptive_beam %g\n", frame, list_.size(), adaptive_beam);
    {  // PossiblyResizeHash :219-225
      const size_t new_sz = static_cast<size_t>(static_cast<float>(list_.size()) * o_.hash_ratio);
      if (new_sz > map_.Size()) map_.SetSize(new_sz);
    }
    const float *ll = ll_ + (size_t)frame * ld_;
    float next_cutoff = kInf, cost_offset = 0.f;
    if (best >= 0) {
      const int state = list_[best].first;
      const float tot = list_[best].tot;
      cost_offset = -tot;
      for (uint32_t a = g_.e_begin[state]; a < g_.e_begin[state + 1]; a++) {
        const float new_weight = ea_[a].weight + cost_offset - ll[ea_[a].pdf] + tot;
        if (new_weight + adaptive_beam < next_cutoff) next_cutoff = new_weight + adaptive_beam;
      }
    }
    cost_offsets_.push_back(cost_offset);
    if (debug_) {  // every state an arc inside the SEED cutoff reaches: the superset of what any visiting order admits
      sup_.clear();
      for (const auto &st : list_)
        if (toks_[st.second].tot <= cur_cutoff)
          for (uint32_t a = g_.e_begin[st.first]; a < g_.e_begin[st.first + 1]; a++) {
            const float t = toks_[st.second].tot + (cost_offset - ll[e_pdf_[a]]) + g_.e_weight[a];
            if (t < next_cutoff) {
              auto it = sup_.find(g_.e_next[a]);
              if (it == sup_.end() || t < it->second) sup_[g_.e_next[a]] = t;
            }
          }
    }
    const size_t n_list = list_.size();
    for (size_t li = 0; li < n_list; li++) {
      const auto &st = list_[li];
      // the list is known in advance and the graph is far larger than a core's caches: the CSR offsets of the token
      // eight places ahead and the arc records of the token four places ahead are requested now
      if (li + 8 < n_list) __builtin_prefetch(&g_.e_begin[list_[li + 8].first]);
      if (li + 4 < n_list) __builtin_prefetch(&ea_[g_.e_begin[list_[li + 4].first]]);
      const int state = st.first, tok = st.second;
      const float cur_cost = st.tot;  // (a token of this frame is never the target of an arc of this frame)
      if (cur_cost <= cur_cutoff) {
        expanded_++;
        const uint32_t a_end = g_.e_begin[state + 1];
        arcs_ += a_end - g_.e_begin[state];
        for (uint32_t a = g_.e_begin[state]; a < a_end; a++) {
          const StrictArcs::Arc arc = ea_[a];  // {next, pdf, weight} side by side: one cache line per state's arcs
          const float ac_cost = cost_offset - ll[arc.pdf], graph_cost = arc.weight;
          const float tot_cost = cur_cost + ac_cost + graph_cost;
          if (tot_cost >= next_cutoff) continue;
          else if (tot_cost + adaptive_beam < next_cutoff) next_cutoff = tot_cost + adaptive_beam;
          const int e = FindOrAdd(arc.next, tot_cost, tok, (int)a, nullptr);
          if (lattice_) AddLink(tok, map_.At(e).tok, (int)a, graph_cost, ac_cost);
        }
      }
    }
    prev_cutoff_ = next_cutoff;
    return next_cutoff;
  }

  // debug: would the device decoder's safe-frame rules (decode.cu) flag this frame?  A = tokens inside the previous
  // frame's final cutoff (what the device keeps), E = the other states arcs inside the seed cutoff reached
  void SafeFrameProbe(int frame) {
    std::vector<float> a_cost;
    std::map<int, float> in_list;
    for (const auto &st : list_) in_list[st.first] = toks_[st.second].tot;
    for (const auto &kv : in_list)
      if (kv.second < prev_cutoff_) a_cost.push_back(kv.second);
    int n_e = 0;
    float min_e = kInf;
    for (const auto &kv : sup_) {
      auto it = in_list.find(kv.first);
      const float c = it != in_list.end() ? std::min(it->second, kv.second) : kv.second;
      if (c >= prev_cutoff_) {
        n_e++;
        min_e = std::min(min_e, c);
      }
    }
    const int n = (int)a_cost.size();
    if (n == 0 || n_e == 0) return;
    const float best = *std::min_element(a_cost.begin(), a_cost.end()), beam_cutoff = best + o_.beam;
    float cur_cutoff = beam_cutoff;
    const char *why = nullptr;
    bool binding = false;
    if (n > o_.max_active) {
      std::nth_element(a_cost.begin(), a_cost.begin() + o_.max_active, a_cost.end());
      if (a_cost[o_.max_active] < beam_cutoff) {
        binding = true;
        cur_cutoff = a_cost[o_.max_active];
      }
    }
    if (!binding) {
      if (n <= o_.max_active && n + n_e > o_.max_active && min_e < beam_cutoff) why = "max-active count";
      if (n <= o_.min_active) {
        why = "min-active count";
        cur_cutoff = kInf;
      } else {
        int inside = 0;
        for (float c : a_cost) inside += c <= beam_cutoff;
        if (inside <= o_.min_active) {
          std::nth_element(a_cost.begin(), a_cost.begin() + o_.min_active, a_cost.end());
          cur_cutoff = std::max(beam_cutoff, a_cost[o_.min_active]);
        }
      }
    }
    if (!why && min_e <= cur_cutoff) why = "extra inside the cutoff";
    if (why) {
      unsafe_frames_++;
      fprintf(stderr, "UNSAFE frame %d: %s (A %d, E %d, min E - best %.3f, cutoff - best %.3f)\n", frame, why, n, n_e, min_e - best,
              cur_cutoff - best);
    }
  }

  // ProcessNonemitting :820-887
  void ProcessNonemitting(float cutoff) {
    queue_.clear();
    for (int e = map_.Head(); e >= 0; e = map_.At(e).tail)
      if (HasEps(map_.At(e).state)) queue_.push_back(e);
    while (!queue_.empty()) {
      const int e = queue_.back();
      queue_.pop_back();
      const int state = map_.At(e).state, tok = map_.At(e).tok;
      const float cur_cost = map_.At(e).tot;
      if (cur_cost >= cutoff) continue;
      toks_[tok].links = -1;  // DeleteForwardLinks: they are regenerated below
      for (uint32_t a = g_.p_begin[state]; a < g_.p_begin[state + 1]; a++) {
        arcs_++;
        const StrictArcs::Arc arc = pa_[a];
        const float graph_cost = arc.weight, tot_cost = cur_cost + graph_cost;
        if (tot_cost < cutoff) {
          bool changed;
          const int e_new = FindOrAdd(arc.next, tot_cost, tok, NE_ + (int)a, &changed);
          if (lattice_) AddLink(tok, map_.At(e_new).tok, NE_ + (int)a, graph_cost, 0.f);
          if (changed && HasEps(arc.next)) queue_.push_back(e_new);
        }
      }
    }
  }

  // best path with final probabilities (lattice-faster-online-decoder.cc:78-173); the cost sums run from the end of
  // the path to its start, as the device back-trace (decode.cu) forms them
  void BestPath(int n_frames, StrictResult *out) {
    const int b = frame_begin_[n_frames], e = frame_begin_[n_frames + 1];
    bool any_final = false;
    for (int t = b; t < e; t++) any_final |= g_.final_cost[toks_[t].state] != kInf;
    float best_cost = kInf;
    int best_tok = -1;
    for (int t = e - 1; t >= b; t--) {  // the frame's token list: newest first
      float cost = toks_[t].tot;
      if (any_final) {
        const float fc = g_.final_cost[toks_[t].state];
        cost = fc != kInf ? cost + fc : kInf;
      }
      if (cost < best_cost) {
        best_cost = cost;
        best_tok = t;
      }
    }
    any_final_ = any_final;
    if (best_tok < 0) return;
    out->decoded = true;
    float graph = any_final ? g_.final_cost[toks_[best_tok].state] : 0.f, acoustic = 0.f;
    int f = n_frames - 1;
    std::vector<int> rev;
    for (int t = best_tok; t >= 0; t = toks_[t].back) {
      const int arc = toks_[t].back_arc;
      if (arc < 0) continue;
      if (arc < NE_) {
        graph += g_.e_weight[arc];
        acoustic -= ll_[(size_t)f * ld_ + e_pdf_[arc]];
        f--;
        if (g_.e_olabel[arc]) rev.push_back(g_.e_olabel[arc]);
      } else {
        graph += g_.p_weight[arc - NE_];
        if (g_.p_olabel[arc - NE_]) rev.push_back(g_.p_olabel[arc - NE_]);
      }
    }
    out->graph = graph;
    out->acoustic = acoustic;
    if ((int)rev.size() > o_.max_words) {  // the device path keeps the last max_words words
      out->word_overflow = true;
      rev.resize(o_.max_words);
    }
    out->words.assign(rev.rbegin(), rev.rend());
  }

  // FinalizeDecoding :625-640 (PruneForwardLinksFinal :376-458, PruneForwardLinks :299-370, PruneTokensForFrame
  // :479-498) followed by GetRawLattice :106-189.  The periodic PruneActiveTokens (:506-533) is not replayed: it only
  // excises links on lower bounds of the extra costs computed here, i.e. links this pass excises as well.
  void Lattice(int n_frames, StrictResult *out) {
    const float lb = o_.lattice_beam;
    auto link_extra = [&](const Tok &tok, const Link &l) {
      const Tok &nt = toks_[l.next_tok];
      return nt.extra + ((tok.tot + l.acoustic + l.graph) - nt.tot);
    };
    // sweeps the links of `tok`: excises those beyond the lattice beam, returns min(tok_extra, cheapest kept link)
    auto sweep = [&](int t, float tok_extra) {
      Tok &tok = toks_[t];
      int *slot = &tok.links;
      while (*slot >= 0) {
        Link &l = links_[*slot];
        float x = link_extra(tok, l);
        if (x > lb) {
          *slot = l.next;
        } else {
          if (x < 0.f) x = 0.f;
          if (x < tok_extra) tok_extra = x;
          slot = &l.next;
        }
      }
      return tok_extra;
    };
    {  // last time, with the final costs
      const int b = frame_begin_[n_frames], e = frame_begin_[n_frames + 1];
      float best = kInf, best_final = kInf;
      for (int t = b; t < e; t++) {
        best = std::min(toks_[t].tot, best);
        best_final = std::min(toks_[t].tot + g_.final_cost[toks_[t].state], best_final);
      }
      const float final_best_cost = best_final != kInf ? best_final : best;
      bool changed = true;
      while (changed) {
        changed = false;
        for (int t = e - 1; t >= b; t--) {
          const float fc = !any_final_ ? 0.f : g_.final_cost[toks_[t].state];
          float x = sweep(t, toks_[t].tot + fc - final_best_cost);
          if (x > lb) x = kInf;
          if (!ApproxEq(toks_[t].extra, x, 1.0e-05f)) changed = true;
          toks_[t].extra = x;
        }
      }
    }
    for (int f = n_frames - 1; f >= 0; f--) {
      const int b = frame_begin_[f], e = frame_begin_[f + 1];
      bool changed = true;
      while (changed) {
        changed = false;
        for (int t = e - 1; t >= b; t--) {
          const float x = sweep(t, kInf);
          if (std::fabs(x - toks_[t].extra) > 0.f) changed = true;
          toks_[t].extra = x;
        }
      }
    }
    // surviving tokens, numbered time-major in creation order; their surviving links
    std::vector<int> id(toks_.size(), -1);
    int n = 0;
    for (size_t t = 0; t < toks_.size(); t++)
      if (toks_[t].extra != kInf) id[t] = n++;
    out->n_nodes = n;
    if (debug_)
      for (int f = 0; f <= n_frames; f++) {
        int alive = 0;
        for (int t = frame_begin_[f]; t < frame_begin_[f + 1]; t++) alive += id[t] >= 0;
        fprintf(stderr, "strict lattice time %d nodes %d of %d\n", f, alive, frame_begin_[f + 1] - frame_begin_[f]);
      }
    for (int f = 0; f <= n_frames; f++) {
      const float off = f < n_frames ? cost_offsets_[f] : 0.f;
      for (int t = frame_begin_[f]; t < frame_begin_[f + 1]; t++) {
        if (id[t] < 0) continue;
        for (int l = toks_[t].links; l >= 0; l = links_[l].next) {
          const Link &k = links_[l];
          if (id[k.next_tok] < 0) continue;
          LatticeArc a;
          a.src = id[t];
          a.dst = id[k.next_tok];
          if (k.arc < NE_) {
            a.olabel = g_.e_olabel[k.arc];
            a.acoustic = k.acoustic - off;
          } else {
            a.olabel = g_.p_olabel[k.arc - NE_];
            a.acoustic = k.acoustic;
          }
          a.graph = k.graph;
          out->lattice.push_back(a);
        }
        if (f == n_frames) {
          const float fc = g_.final_cost[toks_[t].state];
          if (!any_final_ || fc != kInf) out->lattice.push_back(LatticeArc{id[t], -1, 0, any_final_ ? fc : 0.f, 0.f});
        }
      }
    }
  }

  const Graph &g_;
  const int32_t *e_pdf_;
  const StrictArcs::Arc *ea_, *pa_;  // emitting / epsilon arcs as records (the search loops read nothing else of an arc)
  const uint64_t *eps_bits_;         // "state has epsilon arcs", asked for every token of every frame
  const float *ll_;
  const int ld_;
  const StrictOptions o_;
  const bool lattice_;
  const int NE_;
  OrderedStateMap map_;
  std::vector<Tok> toks_;
  std::vector<Link> links_;
  std::vector<int> frame_begin_;
  std::vector<float> cost_offsets_, tmp_;
  std::vector<OrderedStateMap::Item> list_;
  std::vector<int> queue_;
  bool any_final_ = false;
  std::map<int, float> sup_;
  float prev_cutoff_ = kInf;
  int unsafe_frames_ = 0;
  const bool debug_ = getenv("RS_B200_STRICT_DEBUG") != nullptr;
  uint64_t expanded_ = 0, arcs_ = 0;
};

}  // namespace

void BuildStrictArcs(const Graph &g, const int32_t *e_pdf, StrictArcs *out) {
  out->emitting.resize(g.e_next.size());
  for (size_t a = 0; a < g.e_next.size(); a++) out->emitting[a] = StrictArcs::Arc{g.e_next[a], e_pdf[a], g.e_weight[a]};
  out->epsilon.resize(g.p_next.size());
  for (size_t a = 0; a < g.p_next.size(); a++) out->epsilon[a] = StrictArcs::Arc{g.p_next[a], -1, g.p_weight[a]};
  out->has_epsilon.assign(((size_t)g.num_states + 63) / 64 + 1, 0);
  for (int s = 0; s < g.num_states; s++)
    if (g.p_begin[s + 1] > g.p_begin[s]) out->has_epsilon[(size_t)s >> 6] |= uint64_t(1) << (s & 63);
}

void StrictDecode(const Graph &g, const int32_t *e_pdf, const StrictArcs &arcs, const float *loglikes, int ld, int n_frames,
                  const StrictOptions &opt, bool want_lattice, StrictResult *out) {
  *out = StrictResult();
  if (n_frames <= 0 || g.num_states <= 0) return;
  if (arcs.emitting.size() != g.e_next.size() || arcs.epsilon.size() != g.p_next.size() ||
      arcs.has_epsilon.size() != ((size_t)g.num_states + 63) / 64 + 1)
    throw Error("StrictDecode: the arc records do not belong to this graph");
  Search s(g, e_pdf, arcs, loglikes, ld, opt, want_lattice);
  s.Run(n_frames, out);
}

void StrictDecode(const Graph &g, const int32_t *e_pdf, const float *loglikes, int ld, int n_frames,
                  const StrictOptions &opt, bool want_lattice, StrictResult *out) {
  StrictArcs arcs;
  BuildStrictArcs(g, e_pdf, &arcs);
  StrictDecode(g, e_pdf, arcs, loglikes, ld, n_frames, opt, want_lattice, out);
}

}  // namespace rs
