// HCLG.fst reader: OpenFst "const" and "vector" FSTs over StdArc (tropical, int32 labels).
// Formats: reference kaldi/openfst/src/lib/fst.cc:58-82 (header),
// include/fst/const-fst.h:192-232 (ConstFst body), include/fst/vector-fst.h:445-484 (VectorFst
// body), lib/symbol-table.cc (embedded symbol tables, skipped).  Which types the reference
// accepts: kaldi/src/fstext/kaldi-fst-io.cc:51-91 (ReadFstKaldiGeneric).
#include <cmath>
#include <cstring>
#include <limits>

#include "model.h"

namespace rs {

namespace {
struct Cursor {
  const std::string &b;
  size_t p = 0;
  const std::string &path;
  template <typename T>
  T Get() {
    if (p + sizeof(T) > b.size()) RS_FAIL(path << ": truncated FST file");
    T v;
    memcpy(&v, b.data() + p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  std::string Str() {
    int32_t n = Get<int32_t>();
    if (n < 0 || p + (size_t)n > b.size()) RS_FAIL(path << ": bad string in FST header");
    std::string s = b.substr(p, n);
    p += n;
    return s;
  }
  void Align16() { p = (p + 15) & ~(size_t)15; }
  void SkipSymbolTable() {
    int32_t magic = Get<int32_t>();
    if (magic != 2125658996) RS_FAIL(path << ": bad symbol table magic");
    Str();
    Get<int64_t>();
    int64_t n = Get<int64_t>();
    for (int64_t i = 0; i < n; i++) {
      Str();
      Get<int64_t>();
    }
  }
};
}  // namespace

void LoadGraph(const std::string &hclg_fst, const std::string &words_txt, Graph *g) {
  std::string buf;
  {
    std::ifstream f(hclg_fst, std::ios::binary);
    if (!f) RS_FAIL("cannot open " << hclg_fst);
    std::stringstream ss;
    ss << f.rdbuf();
    buf = ss.str();
  }
  Cursor c{buf, 0, hclg_fst};
  if (c.Get<int32_t>() != 2125659606) RS_FAIL(hclg_fst << ": not an OpenFst binary file (bad magic number)");
  std::string fsttype = c.Str(), arctype = c.Str();
  int32_t version = c.Get<int32_t>();
  int32_t flags = c.Get<int32_t>();
  c.Get<uint64_t>();  // properties
  int64_t start = c.Get<int64_t>(), ns = c.Get<int64_t>(), na = c.Get<int64_t>();
  if (arctype != "standard") RS_FAIL(hclg_fst << ": FST arc type '" << arctype << "' is not supported (need 'standard')");
  if (flags & 1) c.SkipSymbolTable();
  if (flags & 2) c.SkipSymbolTable();

  struct Arc {
    int32_t ilabel, olabel;
    float weight;
    int32_t next;
  };
  std::vector<std::vector<Arc>> per_state;  // only for vector FSTs
  std::vector<float> final_w;
  std::vector<uint32_t> pos, narcs;
  const Arc *arcs = nullptr;
  std::vector<Arc> arc_store;
  if (fsttype == "const") {
    bool aligned = (flags & 4) || version == 1;
    if (ns < 0 || na < 0) RS_FAIL(hclg_fst << ": bad ConstFst header");
    if (aligned) c.Align16();
    if (c.p + (size_t)ns * 20 > buf.size()) RS_FAIL(hclg_fst << ": truncated ConstFst state table");
    final_w.resize(ns);
    pos.resize(ns);
    narcs.resize(ns);
    for (int64_t s = 0; s < ns; s++) {
      final_w[s] = c.Get<float>();
      pos[s] = c.Get<uint32_t>();
      narcs[s] = c.Get<uint32_t>();
      c.Get<uint32_t>();
      c.Get<uint32_t>();
    }
    if (aligned) c.Align16();
    if (c.p + (size_t)na * 16 > buf.size()) RS_FAIL(hclg_fst << ": truncated ConstFst arc table");
    arc_store.resize(na);
    if (na) memcpy(arc_store.data(), buf.data() + c.p, (size_t)na * 16);
    arcs = arc_store.data();
    for (int64_t s = 0; s < ns; s++)
      if ((uint64_t)pos[s] + narcs[s] > (uint64_t)na) RS_FAIL(hclg_fst << ": ConstFst arc range out of bounds");
  } else if (fsttype == "vector") {
    int64_t s = 0;
    for (; ns == -1 || s < ns; s++) {
      if (c.p + 4 > buf.size()) break;
      final_w.push_back(c.Get<float>());
      int64_t n = c.Get<int64_t>();
      pos.push_back((uint32_t)arc_store.size());
      narcs.push_back((uint32_t)n);
      for (int64_t i = 0; i < n; i++) {
        Arc a;
        a.ilabel = c.Get<int32_t>();
        a.olabel = c.Get<int32_t>();
        a.weight = c.Get<float>();
        a.next = c.Get<int32_t>();
        arc_store.push_back(a);
      }
    }
    if (ns != -1 && s != ns) RS_FAIL(hclg_fst << ": unexpected end of VectorFst");
    ns = s;
    na = (int64_t)arc_store.size();
    arcs = arc_store.data();
  } else {
    RS_FAIL(hclg_fst << ": FST type '" << fsttype << "' is not supported (need 'const' or 'vector')");
  }
  if (ns >= (int64_t)1 << 31 || na >= (int64_t)1 << 31) RS_FAIL(hclg_fst << ": graph too large");
  if (start < 0 || start >= ns) RS_FAIL(hclg_fst << ": FST has no start state");
  g->start = start;
  g->num_states = (int32_t)ns;
  g->final_cost = final_w;
  g->e_begin.assign(ns + 1, 0);
  g->p_begin.assign(ns + 1, 0);
  for (int64_t s = 0; s < ns; s++) {
    uint32_t ne = 0, np = 0;
    for (uint32_t i = 0; i < narcs[s]; i++) {
      const Arc &a = arcs[pos[s] + i];
      if (a.next < 0 || a.next >= ns) RS_FAIL(hclg_fst << ": arc to a non-existent state");
      if (a.ilabel < 0) RS_FAIL(hclg_fst << ": negative input label");
      (a.ilabel != 0 ? ne : np)++;
    }
    g->e_begin[s + 1] = g->e_begin[s] + ne;
    g->p_begin[s + 1] = g->p_begin[s] + np;
  }
  size_t NE = g->e_begin[ns], NP = g->p_begin[ns];
  g->e_ilabel.resize(NE);
  g->e_olabel.resize(NE);
  g->e_next.resize(NE);
  g->e_src.resize(NE);
  g->e_weight.resize(NE);
  g->p_olabel.resize(NP);
  g->p_next.resize(NP);
  g->p_src.resize(NP);
  g->p_weight.resize(NP);
  for (int64_t s = 0; s < ns; s++) {
    uint32_t e = g->e_begin[s], p = g->p_begin[s];
    for (uint32_t i = 0; i < narcs[s]; i++) {
      const Arc &a = arcs[pos[s] + i];
      if (a.ilabel != 0) {
        g->e_ilabel[e] = a.ilabel;
        g->e_olabel[e] = a.olabel;
        g->e_next[e] = a.next;
        g->e_src[e] = (int32_t)s;
        g->e_weight[e] = a.weight;
        e++;
      } else {
        g->p_olabel[p] = a.olabel;
        g->p_next[p] = a.next;
        g->p_src[p] = (int32_t)s;
        g->p_weight[p] = a.weight;
        p++;
      }
    }
  }
  g->words.clear();
  if (!words_txt.empty()) {
    std::ifstream f(words_txt);
    if (!f) RS_FAIL("cannot open " << words_txt);
    std::string sym;
    long id;
    while (f >> sym >> id) {
      if (id < 0) continue;
      if ((size_t)id >= g->words.size()) g->words.resize(id + 1);
      g->words[id] = sym;
    }
  }
}

}  // namespace rs
