// Stage (iii): token-passing beam search over HCLG, one resident CTA per utterance lane.
//
// Replaces LatticeFasterOnlineDecoder as the reference drives it
// (kaldi/src/decoder/lattice-faster-decoder.cc):
//   InitDecoding :56-73, GetCutoff :644-711, ProcessEmitting :714-804, ProcessNonemitting :820-887,
//   FindOrAddToken :252-293; best path: lattice-faster-online-decoder.cc:56-173 (BestPathEnd,
//   TraceBackBestPath), with LogLikelihood(frame, tid) = loglikes[frame][tid2pdf[tid]]
//   (nnet3/decodable-online-looped.cc:249-256).
//
// Design (not a port of the pointer-chasing CPU decoder, nor of Kaldi's cudadecoder):
//   * a persistent CTA per lane walks the frames of one utterance; the frontier lives in compact
//     (state, cost) arrays, the next frontier in an open-addressing table keyed by state whose
//     64-bit values pack (ordered cost, arc id) so that one atomicMin implements "keep the
//     cheapest token for this state and remember the arc it came through";
//   * emitting arcs are expanded one thread per arc: a block-wide prefix sum of out-degrees of the
//     surviving tokens + binary search maps a flat arc index to (token, arc);
//   * epsilon closure is a frontier-relaxation loop to the fix point;
//   * every surviving token appends one (previous token, arc) record to a per-lane arena; the best
//     path is traced back on the device and only word ids leave the GPU.
// Token costs use the reference's float expression order without FMA contraction, so costs, cutoffs
// and therefore the surviving token sets are bit-identical wherever the reference itself is
// order-independent ("decoder semantics").
//
// Where the reference IS order-dependent -- "safe frame" rules.  ProcessEmitting admits an arc against a transient
// next_cutoff (:780-787), so its token list also holds "extras": tokens whose cost lies between the frame's final
// cutoff and the cutoff seeded from the best token (:744-759), WHICH of them depends on the order of its token hash.
// This kernel keeps the tokens inside the final cutoff (the set A, order-independent) and inserts every arc inside
// the seed cutoff, so the entries it drops at compaction are a superset E of the reference's extras.  Extras are
// never expanded by ProcessNonemitting (cost >= cutoff) and die after one frame; they can matter only on the next
// frame, through GetCutoff's token count / nth_element or by being expanded.  With n = |A|, ne = |E|, me = min cost in E:
//   * --max-active binds on A (n > max_active, its cutoff below the beam cutoff): the cutoff is the same with any
//     subset of E added (extras cost more than every token of A), and it lies below me: safe;
//   * otherwise unsafe if n <= max_active < n + ne and me < beam cutoff (extras could make max-active bind),
//     or n <= min_active (the count test of :691, or an infinite cutoff under which extras would be expanded);
//   * unsafe if me <= the frame's cutoff (an extra would be expanded);
//   * last frame: unsafe if an extra with its final cost could be the best final token (or, n-best, lie inside the
//     lattice beam of it).
// An utterance with an unsafe frame carries status bit 4 (16) and is decoded again by strict_decode.cc; by induction
// over the frames an unflagged utterance has exactly the reference's tokens inside the final cutoffs, its costs, its
// cutoffs and its best path.  (Small graphs never get here: decode_small.cu reproduces the order itself.)
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "decode_common.cuh"
#include "engine.h"
#include "smem_attr.h"

namespace rs {

// CTA size is a launch-time choice (128 .. 512 threads): the frontier of a grammar graph is a few
// hundred tokens, where a small CTA pays far less per __syncthreads than a large one
constexpr int kMaxNT = 512;
constexpr int kMaxNW = kMaxNT / 32;
#define NT ((int)blockDim.x)
#define NW ((int)(blockDim.x >> 5))
constexpr int kEmptyKey = -1;

static int g_decode_threads = 0;
int DecodeCtaThreads() {
  if (g_decode_threads == 0) {
    const char *e = getenv("RS_B200_DECODE_THREADS");
    int v = e ? atoi(e) : 512;
    if (v != 32 && v != 64 && v != 128 && v != 256 && v != 512) v = 512;
    g_decode_threads = v;
  }
  return g_decode_threads;
}

__device__ __forceinline__ unsigned hash_state(int s) { return (unsigned)s * 2654435761u; }

struct Shared {
  unsigned warp_sums[kMaxNW + 1];
  float red_v[kMaxNW];
  int red_i[kMaxNW];
  unsigned hist[256];
  unsigned sel_prefix, sel_mask;
  int sel_k;
  unsigned nc_ord;
  int n_ins[2];
  int frontier_n[2];
  int n_next;
  int overflow;
  int utt;
  float best_cost;
  int best_idx;
  int any_final;
  int n_links;       // lattice mode: links recorded so far for the utterance
  int lat_overflow;  // lattice mode: link capacity exceeded
  unsigned seed_ord;             // next_cutoff as seeded from the best token's arcs (:744-759)
  unsigned min_extra_ord;        // cheapest entry beyond the final cutoff of the frame being finalised (ord())
  unsigned min_extra_final_ord;  // last frame: cheapest such entry with its final cost added
};

__device__ __forceinline__ unsigned block_excl_scan(unsigned v, Shared &S, unsigned *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) S.warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < NW ? S.warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    if (lane < NW) S.warp_sums[lane] = w;
  }
  __syncthreads();
  unsigned base = warp ? S.warp_sums[warp - 1] : 0;
  *total = S.warp_sums[NW - 1];
  __syncthreads();
  return base + x - v;
}

// min over the block of (v, i) with ties broken towards the smaller i; result broadcast via smem
__device__ __forceinline__ void block_min(float v, int i, Shared &S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov < v || (ov == v && oi < i)) {
      v = ov;
      i = oi;
    }
  }
  if (lane == 0) {
    S.red_v[warp] = v;
    S.red_i[warp] = i;
  }
  __syncthreads();
  if (warp == 0) {  // second level: one shuffle reduction over the warp results
    float bv = lane < NW ? S.red_v[lane] : __int_as_float(0x7f800000);
    int bi = lane < NW ? S.red_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      S.best_cost = bv;
      S.best_idx = bi;
    }
  }
  __syncthreads();
}

// k-th smallest (0-based) of a[0..n): 4-pass radix select on the order-preserving key;
// value-exact, the same value std::nth_element leaves at position k (GetCutoff :680-700).
__device__ float block_select(const float *a, int n, int k, Shared &S) {
  if (threadIdx.x == 0) {
    S.sel_prefix = 0;
    S.sel_mask = 0;
    S.sel_k = k;
  }
  for (int pass = 3; pass >= 0; pass--) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += NT) S.hist[i] = 0;
    __syncthreads();
    const unsigned prefix = S.sel_prefix, mask = S.sel_mask;
    for (int i = threadIdx.x; i < n; i += NT) {
      unsigned key = ord(a[i]);
      if ((key & mask) == prefix) atomicAdd(&S.hist[(key >> shift) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // warp 0 locates the bin that holds the k-th key: 8 bins per lane, shuffle scan over the lane sums
      // (a serial scan of the 256 bins by one thread cost ~7 k cycles per pass, 4 passes per call)
      const int lane = threadIdx.x;
      unsigned h[8], sum = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        h[i] = S.hist[lane * 8 + i];
        sum += h[i];
      }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const unsigned excl = incl - sum, kk = (unsigned)S.sel_k;
      __syncwarp();  // every lane has read sel_k before the owning lane overwrites it (racecheck: WAR inside the warp)
      const bool mine = kk >= excl && kk < incl;  // exactly one lane (k < n)
      if (mine) {
        unsigned rem = kk - excl, b = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (b == (unsigned)i && rem >= h[i]) {
            rem -= h[i];
            b = i + 1;
          }
        }
        S.sel_k = (int)rem;
        S.sel_prefix = prefix | ((unsigned)(lane * 8 + b) << shift);
        S.sel_mask = mask | (255u << shift);
      }
    }
    __syncthreads();
  }
  return unord(S.sel_prefix);
}

// volatile (uncached) loads of table words that other threads update with atomics; the tables live
// in global memory or -- for small graphs -- in shared memory, so the address is generic
__device__ __forceinline__ int ldv(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ unsigned long long ldv(const unsigned long long *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}

struct Table {
  int *hkey;
  unsigned long long *hval;
  int *hidx;
  int *ins_list;
  int *n_ins;  // shared memory counter
};

__device__ __forceinline__ int find_slot(const int *hkey, int state, unsigned mask, bool identity) {
  unsigned s = identity ? (unsigned)state : (hash_state(state) & mask);
  while (true) {
    int k = ldv(hkey + s);
    if (k == state) return (int)s;
    if (k == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
}

__device__ __forceinline__ int insert_slot(const Table &t, int state, unsigned mask, bool identity, int cap, int *overflow) {
  unsigned s = identity ? (unsigned)state : (hash_state(state) & mask);
  while (true) {
    int k = ldv(t.hkey + s);
    if (k == state) return (int)s;
    if (k == kEmptyKey) {
      int prev = atomicCAS(t.hkey + s, kEmptyKey, state);
      if (prev == kEmptyKey) {
        int idx = atomicAdd(t.n_ins, 1);
        if (idx < cap)
          t.ins_list[idx] = (int)s;
        else
          *overflow = 1;
        return (int)s;
      }
      if (prev == state) return (int)s;
    }
    s = (s + 1) & mask;
  }
}

// kLat = true additionally records the state-level lattice (every token with its forward cost, every
// forward link the reference would hold after ProcessEmitting / ProcessNonemitting) for the n-best tail.
template <bool kLat>
__global__ void __launch_bounds__(kMaxNT, 2) decode_kernel(const __grid_constant__ DecodeParams P) {
  __shared__ Shared S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  LaneWorkspace ws = P.lanes[blockIdx.x];
  const DevGraph &g = P.g;
  const DecodeConfig &cfg = P.cfg;
  // Small graphs (a grammar HCLG has a few hundred states): the per-frame state tables -- keys,
  // packed (cost, arc) values, token lists, epsilon frontier, prefix sums -- are carved out of shared
  // memory and addressed by state id, so token deduplication is a shared-memory atomicMin instead of
  // an L2 round trip.  Only the traceback arena stays in HBM.
  extern __shared__ __align__(16) unsigned char dsm[];
  const int smem_slots = cfg.smem_slots;
  const int tok_cap = smem_slots ? smem_slots : P.cfg.tok_cap;
  const int hash_size = smem_slots ? smem_slots : P.cfg.hash_size;
  if (smem_slots) {
    unsigned char *q = dsm;
    const size_t H = (size_t)smem_slots;
    for (int k = 0; k < 2; k++) {
      ws.hval[k] = reinterpret_cast<unsigned long long *>(q);
      q += 8 * H;
    }
    auto take = [&](size_t n) {
      int *r = reinterpret_cast<int *>(q);
      q += 4 * n;
      return r;
    };
    for (int k = 0; k < 2; k++) {
      ws.hkey[k] = take(H);
      ws.hidx[k] = take(H);
      ws.ins_list[k] = take(H);
      ws.tok_state[k] = take(H);
      ws.tok_cost[k] = reinterpret_cast<float *>(take(H));
      ws.frontier[k] = take(H);
    }
    ws.inq = take(H);
    ws.tok_slot = take(H);
    ws.pfx = reinterpret_cast<unsigned *>(take(H + 4));
    for (int i = threadIdx.x; i < smem_slots; i += NT) {
      ws.hkey[0][i] = ws.hkey[1][i] = kEmptyKey;
      ws.hval[0][i] = ws.hval[1][i] = kEmptyVal;
      ws.hidx[0][i] = ws.hidx[1][i] = -1;
      ws.inq[i] = 0;
    }
    __syncthreads();
  }
  const unsigned mask = (unsigned)hash_size - 1u;
  const bool identity = g.num_states <= hash_size;
  const unsigned NE = g.num_earcs;
  const float kInf = __int_as_float(0x7f800000);

  while (true) {
    if (tid == 0) S.utt = atomicAdd(P.next_utt, 1);
    __syncthreads();
    const int u = S.utt;
    __syncthreads();
    if (u >= P.n_utts) break;
    const int n_frames = P.n_frames[u];
    if (n_frames <= 0) {
      if (tid == 0) {
        P.n_words[u] = -1;
        P.status[u] = 0;
        P.cost[2 * u] = 0.f;
        P.cost[2 * u + 1] = 0.f;
        for (int c = 0; c < 4; c++) P.counters[4 * (size_t)u + c] = 0ULL;
      }
      continue;
    }
    if (tid == 0) {
      S.n_ins[0] = S.n_ins[1] = 0;
      S.frontier_n[0] = S.frontier_n[1] = 0;
      S.overflow = 0;
      S.n_links = 0;
      S.lat_overflow = 0;
    }
    __syncthreads();
    // lattice slices of this utterance (LatticeBuf)
    int2 *ltok = nullptr;
    int4 *llink = nullptr;
    int *ltb = nullptr, *lpos = nullptr;
    float *loff = nullptr;
    const int arena_cap = cfg.arena_cap;
    // lat_dead (block-uniform): the lattice of this utterance outgrew its slice; the search goes on and returns
    // the best path, the lattice is marked unusable (tok_base[n_frames + 1] = -1 -> rs_result.status bit 5)
    bool lat_dead = false;
    if constexpr (kLat) {
      ltok = P.lat.tok + (size_t)u * P.lat.tok_cap;
      llink = P.lat.link + (size_t)u * P.lat.link_cap;
      ltb = P.lat.tok_base + (size_t)u * (P.lat.max_t + 2);
      lpos = P.lat.link_pos + (size_t)u * (2 * P.lat.max_t + 4);
      loff = P.lat.cost_offset + (size_t)u * (P.lat.max_t + 1);
    }
    // slack = (cost through the link) - (destination's cost), the bracket of PruneForwardLinks' link_extra_cost
    // (:330-332) in its float order; stored with the link so that the pruning sweep touches no token or arc
    auto add_link = [&](int src, int dst, unsigned arc, float slack) {
      const int li = atomicAdd(&S.n_links, 1);
      if (li < P.lat.link_cap)
        llink[li] = make_int4(src, dst, (int)arc, __float_as_int(slack));
      else
        S.lat_overflow = 1;
    };
    unsigned long long cnt_tokens = 0, cnt_arcs = 0, cnt_created = 0;  // thread 0 / per-thread partials
    // optional phase timing (cfg.profile): thread 0 accumulates SM clocks between phase boundaries
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ph_last = clock64();
    auto tick = [&](int k) {
      if (cfg.profile && tid == 0) {
        const long long now = clock64();
        ph[k] += now - ph_last;
        ph_last = now;
      }
    };
    int status = 0;
    int info = 0;  // bit 16: --max-active decided the beam on some frame (see rs_result.status)
    int n_extra = 0;       // entries the last compaction dropped: superset of the reference's extras on the current frame
    float min_extra = kInf;
    int cur = 0;
    int n_cur = 0;        // alive tokens of the current frame
    int base_cur = 0;     // arena index of the current frame's first token
    int arena_n = 0;

    // One pass of: epsilon closure of table `tb` under `cutoff`, compaction of the alive entries
    // into tok_state/tok_cost[tb], traceback records.  `tprev` is the table of the previous frame.
    auto close_and_finalize = [&](int tb, int tprev, float cutoff, int base_prev, int base_new, bool last) -> int {
      Table T{ws.hkey[tb], ws.hval[tb], ws.hidx[tb], ws.ins_list[tb], &S.n_ins[tb]};
      // ---- ProcessNonemitting (:820-887): frontier = alive tokens whose state has epsilon arcs
      int fcur = 0;
      if (tid == 0) {
        S.frontier_n[0] = S.frontier_n[1] = 0;
        S.min_extra_ord = S.min_extra_final_ord = 0xffffffffu;
      }
      __syncthreads();
      {
        const int n_ins = S.n_ins[tb];
        for (int i = tid; i < n_ins; i += NT) {
          int s = T.ins_list[i];
          int st = ldv(T.hkey + s);
          float c = unord((unsigned)(ldv(T.hval + s) >> 32));
          if (c < cutoff && g.p_begin[st + 1] > g.p_begin[st]) {
            ws.inq[s] = 1;
            int idx = atomicAdd(&S.frontier_n[0], 1);
            ws.frontier[0][idx] = s;
          }
        }
      }
      __syncthreads();
      while (true) {
        const int nf = S.frontier_n[fcur];
        __syncthreads();
        if (nf == 0) break;
        if (tid == 0) S.frontier_n[fcur ^ 1] = 0;
        __syncthreads();
        for (int i = tid; i < nf; i += NT) {
          int s = ws.frontier[fcur][i];
          atomicExch(ws.inq + s, 0);
          __threadfence_block();
          int st = ldv(T.hkey + s);
          float c = unord((unsigned)(ldv(T.hval + s) >> 32));
          if (c >= cutoff) continue;
          for (unsigned a = g.p_begin[st]; a < g.p_begin[st + 1]; a++) {
            int4 arc = g.parc[a];
            float tot = __fadd_rn(c, __int_as_float(arc.z));
            cnt_arcs++;
            if (tot < cutoff) {
              if (*(volatile int *)&S.overflow) break;
              int s2 = insert_slot(T, arc.x, mask, identity, tok_cap, &S.overflow);
              unsigned long long pv = pack(tot, NE + a);
              unsigned long long old = atomicMin(T.hval + s2, pv);
              if (pv < old && g.p_begin[arc.x + 1] > g.p_begin[arc.x]) {
                if (atomicExch(ws.inq + s2, 1) == 0) {
                  int idx = atomicAdd(&S.frontier_n[fcur ^ 1], 1);
                  if (idx < tok_cap)
                    ws.frontier[fcur ^ 1][idx] = s2;
                  else
                    S.overflow = 1;
                }
              }
            }
          }
        }
        __syncthreads();
        if (S.overflow) break;
        fcur ^= 1;
      }
      __syncthreads();
      if (S.overflow) return -1;
      tick(3);
      // ---- compact the alive entries (cost < cutoff)
      const int n_ins = S.n_ins[tb];
      unsigned running = 0;
      for (int b0 = 0; b0 < n_ins; b0 += NT) {
        int i = b0 + tid;
        int s = -1, st = 0;
        float c = 0.f;
        unsigned alive = 0;
        if (i < n_ins) {
          s = T.ins_list[i];
          c = unord((unsigned)(ldv(T.hval + s) >> 32));
          alive = c < cutoff ? 1u : 0u;
          st = ldv(T.hkey + s);
          if (!alive) {  // a possible extra of the reference (safe-frame rules)
            atomicMin(&S.min_extra_ord, ord(c));
            if (last) {
              const float f = g.final_cost[st];
              if (f != kInf) atomicMin(&S.min_extra_final_ord, ord(__fadd_rn(c, f)));
            }
          }
        }
        unsigned total;
        unsigned pos = running + block_excl_scan(alive, S, &total);
        if (alive) {
          ws.tok_state[tb][pos] = st;
          ws.tok_cost[tb][pos] = c;
          ws.tok_slot[pos] = s;
          T.hidx[s] = (int)pos;
        } else if (s >= 0) {
          T.hidx[s] = -1;
        }
        running += total;
      }
      __syncthreads();
      const int n_new = (int)running;
      n_extra = n_ins - n_new;
      min_extra = unord(S.min_extra_ord);
      if (base_new + n_new > arena_cap) return -2;
      tick(4);
      // ---- traceback records: the arc stored with the winning cost names the predecessor state
      for (int pos = tid; pos < n_new; pos += NT) {
        int s = ws.tok_slot[pos];
        unsigned arc = (unsigned)(ldv(T.hval + s) & 0xffffffffULL);
        int prev = -1;
        if (arc == kArcNone) {
          prev = -1;
        } else if (arc < NE) {
          int ss = find_slot(ws.hkey[tprev], g.e_src[arc], mask, identity);
          prev = ss >= 0 ? base_prev + ws.hidx[tprev][ss] : -1;
        } else {
          int ss = find_slot(T.hkey, g.p_src[arc - NE], mask, identity);
          prev = ss >= 0 ? base_new + T.hidx[ss] : -1;
        }
        ws.arena[base_new + pos] = make_int2(prev, (int)arc);
        if constexpr (kLat)
          if (!lat_dead && base_new + n_new <= P.lat.tok_cap)
            ltok[base_new + pos] = make_int2(ws.tok_state[tb][pos], __float_as_int(ws.tok_cost[tb][pos]));
      }
      if constexpr (kLat)
        if (base_new + n_new > P.lat.tok_cap) lat_dead = true;
      __syncthreads();
      tick(5);
      return n_new;
    };
    // Lattice mode: the epsilon links of the finalised time (ProcessNonemitting :858-884 recreates the links
    // of a token from its final cost: every epsilon arc with tot_cost < cutoff).
    // A link whose cost exceeds its destination token's by more than lattice_beam is dropped here already:
    // PruneForwardLinks excises it whatever happens later, because link_extra_cost = extra_cost[next] + (cost
    // through the link - next.tot_cost) with extra_cost >= 0 (:330-337; the same float expression, and x + y >= y
    // holds in round-to-nearest).  On a grammar graph that is 98 % of the links.
    auto eps_links = [&](int tb, float cutoff, int base_new, int n_new) {
      for (int pos = tid; pos < n_new; pos += NT) {
        const int st = ws.tok_state[tb][pos];
        const float c = ws.tok_cost[tb][pos];
        for (unsigned a = g.p_begin[st]; a < g.p_begin[st + 1]; a++) {
          const int4 arc = g.parc[a];
          const float tot = __fadd_rn(c, __int_as_float(arc.z));
          if (tot < cutoff) {
            const int ss = find_slot(ws.hkey[tb], arc.x, mask, identity);
            const int dpos = ss >= 0 ? ws.hidx[tb][ss] : -1;
            if (dpos >= 0) {
              const float slack = __fsub_rn(tot, ws.tok_cost[tb][dpos]);
              if (!(slack > cfg.lattice_beam)) add_link(base_new + pos, base_new + dpos, NE + a, slack);
            }
          }
        }
      }
      __syncthreads();
    };
    auto clear_table = [&](int tb) {
      const int n_ins = min(S.n_ins[tb], tok_cap);
      for (int i = tid; i < n_ins; i += NT) {
        int s = ws.ins_list[tb][i];
        ws.hkey[tb][s] = kEmptyKey;
        ws.hval[tb][s] = kEmptyVal;
        ws.inq[s] = 0;
      }
      __syncthreads();
      if (tid == 0) S.n_ins[tb] = 0;
      __syncthreads();
    };
    auto wipe_tables = [&]() {  // after an overflow: entries may exist that are not in the lists
      for (int tb = 0; tb < 2; tb++)
        for (int i = tid; i < hash_size; i += NT) {
          ws.hkey[tb][i] = kEmptyKey;
          ws.hval[tb][i] = kEmptyVal;
        }
      for (int i = tid; i < hash_size; i += NT) ws.inq[i] = 0;
      __syncthreads();
      if (tid == 0) {
        S.n_ins[0] = S.n_ins[1] = 0;
        S.overflow = 0;
      }
      __syncthreads();
    };

    // ---- InitDecoding (:56-73): start token, then epsilon closure under cutoff = beam
    if (tid == 0) {
      Table T{ws.hkey[0], ws.hval[0], ws.hidx[0], ws.ins_list[0], &S.n_ins[0]};
      int s = insert_slot(T, g.start, mask, identity, tok_cap, &S.overflow);
      atomicMin(T.hval + s, pack(0.f, kArcNone));
    }
    __syncthreads();
    {
      int r = close_and_finalize(0, 1, cfg.beam, 0, 0, false);
      if (r < 0) {
        status |= (r == -1 ? 1 : 2);
        n_cur = 0;
      } else {
        n_cur = r;
        arena_n = r;
        cnt_created += r;
        if constexpr (kLat) {
          if (tid == 0) ltb[0] = lpos[0] = lpos[1] = 0;
          if (!lat_dead) eps_links(0, cfg.beam, 0, r);
          if (tid == 0) lpos[2] = S.n_links;
        }
      }
    }

    int frame = 0;
    for (; frame < n_frames && status == 0; frame++) {
      if (n_cur == 0) {
        status |= 4;  // "no surviving tokens" (:835-840)
        break;
      }
      const int nxt = cur ^ 1;
      const float *cost = ws.tok_cost[cur];
      const int *state = ws.tok_state[cur];
      const float *ll = P.loglikes + (size_t)(P.ll_row0[u] + frame) * P.ld;
      // ---- GetCutoff (:644-711)
      {
        float bv = kInf;
        int bi = 0x7fffffff;
        for (int i = tid; i < n_cur; i += NT) {
          float c = cost[i];
          if (c < bv) {
            bv = c;
            bi = i;
          }
        }
        block_min(bv, bi, S);
      }
      const float best = S.best_cost;
      const int best_idx = S.best_idx;
      const float beam_cutoff = __fadd_rn(best, cfg.beam);
      float cur_cutoff, adaptive_beam;
      {
        float max_active_cutoff = kInf, min_active_cutoff = kInf;
        if (n_cur > cfg.max_active) max_active_cutoff = block_select(cost, n_cur, cfg.max_active, S);
        if (max_active_cutoff < beam_cutoff) {
          adaptive_beam = __fadd_rn(__fsub_rn(max_active_cutoff, best), cfg.beam_delta);
          cur_cutoff = max_active_cutoff;
        } else {
          // (bits 8-11 say which rule fired: diagnostic detail of bit 4)
          if (n_extra > 0 && n_cur <= cfg.min_active) info |= 16 | 256;
          if (n_extra > 0 && n_cur <= cfg.max_active && n_cur + n_extra > cfg.max_active && min_extra < beam_cutoff) info |= 16 | 512;
          if (n_cur > cfg.min_active) {
            if (cfg.min_active == 0) {
              min_active_cutoff = best;
            } else {
              // the min_active-th smallest cost exceeds the beam cutoff iff at most min_active
              // tokens lie inside the beam; only then is its exact value needed
              unsigned inside = 0, total;
              for (int i = tid; i < n_cur; i += NT) inside += cost[i] <= beam_cutoff ? 1u : 0u;
              block_excl_scan(inside, S, &total);
              if ((int)total > cfg.min_active)
                min_active_cutoff = beam_cutoff;
              else
                min_active_cutoff = block_select(cost, n_cur, cfg.min_active, S);
            }
          }
          if (min_active_cutoff > beam_cutoff) {
            adaptive_beam = __fadd_rn(__fsub_rn(min_active_cutoff, best), cfg.beam_delta);
            cur_cutoff = min_active_cutoff;
          } else {
            adaptive_beam = cfg.beam;
            cur_cutoff = beam_cutoff;
          }
        }
      }
      if (n_extra > 0 && min_extra <= cur_cutoff) info |= 16 | 1024;
      tick(0);
      // ---- ProcessEmitting (:714-804)
      const float cost_offset = -best;
      if (warp == 0) {  // seed next_cutoff from the best token's arcs (:744-759)
        int st = state[best_idx];
        float m = kInf;
        for (unsigned a = g.e_begin[st] + lane; a < g.e_begin[st + 1]; a += 32) {
          int4 arc = g.earc[a];
          float nw = __fadd_rn(__fsub_rn(__fadd_rn(__int_as_float(arc.z), cost_offset), ll[arc.y]), best);
          m = fminf(m, __fadd_rn(nw, adaptive_beam));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) S.nc_ord = S.seed_ord = ord(m);
      }
      // out-degree prefix over the tokens inside the cutoff
      unsigned n_arcs = 0;
      for (int b0 = 0; b0 < n_cur; b0 += NT) {
        int i = b0 + tid;
        unsigned deg = 0;
        if (i < n_cur && cost[i] <= cur_cutoff) {
          int st = state[i];
          deg = g.e_begin[st + 1] - g.e_begin[st];
        }
        unsigned total;
        unsigned ex = block_excl_scan(deg, S, &total);
        if (i < n_cur) ws.pfx[i] = n_arcs + ex;
        n_arcs += total;
      }
      if (tid == 0) ws.pfx[n_cur] = n_arcs;
      __syncthreads();
      tick(1);
      {
        Table T{ws.hkey[nxt], ws.hval[nxt], ws.hidx[nxt], ws.ins_list[nxt], &S.n_ins[nxt]};
        const float seed_cutoff = unord(S.seed_ord);  // every arc inside the SEED cutoff is inserted: a superset of what
                                                    // any visiting order admits, independent of thread timing
        for (unsigned a = tid; a < n_arcs; a += NT) {
          int lo = 0, hi = n_cur;
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (ws.pfx[mid] <= a) lo = mid; else hi = mid;
          }
          const int st = state[lo];
          const unsigned ai = g.e_begin[st] + (a - ws.pfx[lo]);
          const int4 arc = g.earc[ai];
          const float ac = __fsub_rn(cost_offset, ll[arc.y]);
          const float tot = __fadd_rn(__fadd_rn(cost[lo], ac), __int_as_float(arc.z));
          if (tot >= seed_cutoff) continue;
          const float cand = __fadd_rn(tot, adaptive_beam);
          if (cand < seed_cutoff) atomicMin(&S.nc_ord, ord(cand));
          if (*(volatile int *)&S.overflow) break;
          int s2 = insert_slot(T, arc.x, mask, identity, tok_cap, &S.overflow);
          // (a plain load of the slot to skip non-improving arcs before the 64-bit atomic was measured: the extra
          // round trip costs more than the atomics it saves -- ARPA graph 43.8 -> 51.3 ms)
          atomicMin(T.hval + s2, pack(tot, ai));
        }
      }
      __syncthreads();
      if (S.overflow) {
        status |= 1;
        break;
      }
      tick(2);
      const float next_cutoff = unord(S.nc_ord);
      cnt_tokens += n_cur;
      cnt_arcs += (tid == 0) ? n_arcs : 0;
      const int base_new = base_cur + n_cur;
      int r = close_and_finalize(nxt, cur, next_cutoff, base_cur, base_new, frame == n_frames - 1);
      if (r < 0) {
        status |= (r == -1 ? 1 : 2);
        break;
      }
      if constexpr (kLat) if (!lat_dead) {
        // emitting links time frame -> frame + 1: every arc whose cost passed the frame's final next_cutoff
        // (the reference's transient cutoff lets more through; those lie beyond the beam and never survive
        // the lattice-beam pruning), then the epsilon links of the new time
        for (unsigned a = tid; a < n_arcs; a += NT) {
          int lo = 0, hi = n_cur;
          while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (ws.pfx[mid] <= a) lo = mid; else hi = mid;
          }
          const int st = state[lo];
          const unsigned ai = g.e_begin[st] + (a - ws.pfx[lo]);
          const int4 arc = g.earc[ai];
          const float ac = __fsub_rn(cost_offset, ll[arc.y]);
          const float tot = __fadd_rn(__fadd_rn(cost[lo], ac), __int_as_float(arc.z));
          // tot < next_cutoff: a link in any visiting order.  next_cutoff <= tot < seed cutoff: the reference holds
          // this link only if it reached the arc before its transient cutoff had tightened -- recorded as a "maybe"
          // link (arc id with the sign bit set); lattice_prune_kernel keeps it out of the sweep and reports the
          // utterance as order-sensitive if the link would have survived the lattice beam
          if (tot < unord(S.seed_ord)) {
            const int ss = find_slot(ws.hkey[nxt], arc.x, mask, identity);
            const int dpos = ss >= 0 ? ws.hidx[nxt][ss] : -1;
            if (dpos >= 0) {
              const float slack = __fsub_rn(tot, ws.tok_cost[nxt][dpos]);
              if (!(slack > cfg.lattice_beam))
                add_link(base_cur + lo, base_new + dpos, tot < next_cutoff ? ai : (ai | 0x80000000u), slack);
            }
          }
        }
        __syncthreads();
        if (tid == 0) {
          lpos[2 * (frame + 1) + 1] = S.n_links;
          ltb[frame + 1] = base_new;
          loff[frame] = cost_offset;
        }
        __syncthreads();
        eps_links(nxt, next_cutoff, base_new, r);
        if (tid == 0) lpos[2 * (frame + 2)] = S.n_links;
        if (S.lat_overflow) lat_dead = true;  // every thread reads the flag behind the barrier of eps_links
      }
      clear_table(cur);
      tick(6);
      base_cur = base_new;
      n_cur = r;
      arena_n = base_new + r;
      cnt_created += r;
      cur = nxt;
    }

    if (cfg.profile && tid == 0 && blockIdx.x == 0)
      printf("decode phases (clocks, utt %d, %d frames): cutoff %lld seed+prefix %lld expand %lld epsilon %lld compact %lld records %lld clear %lld\n", u,
             n_frames, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6]);
    if constexpr (kLat) {
      if (tid == 0) ltb[n_frames + 1] = lat_dead ? -1 : arena_n;
    }
    // ---- best path (lattice-faster-online-decoder.cc:78-173)
    int n_words = -1;
    if (status == 0 && n_cur == 0) status |= 4;
    if (status == 0) {
      const float *cost = ws.tok_cost[cur];
      const int *state = ws.tok_state[cur];
      int anyf = 0;
      for (int i = tid; i < n_cur; i += NT) anyf |= g.final_cost[state[i]] != kInf;
      anyf = __syncthreads_or(anyf);
      float bv = kInf;
      int bi = 0x7fffffff;
      for (int i = tid; i < n_cur; i += NT) {
        float c = cost[i];
        if (anyf) {
          float f = g.final_cost[state[i]];
          c = f != kInf ? __fadd_rn(c, f) : kInf;
        }
        if (c < bv) {
          bv = c;
          bi = i;
        }
      }
      block_min(bv, bi, S);
      // safe-frame rule of the last frame: an entry beyond the final cutoff that is final and, with its final cost,
      // as cheap as the best final token (n-best: inside the lattice beam of it) -- or final when no kept token is
      if (S.min_extra_final_ord != 0xffffffffu &&
          (!anyf || unord(S.min_extra_final_ord) <= __fadd_rn(S.best_cost, kLat ? cfg.lattice_beam : 0.f)))
        info |= 16 | 2048;
      if (S.best_idx == 0x7fffffff) {
        status |= 4;
      } else if (tid == 0) {
        int gid = base_cur + S.best_idx;
        float graph = anyf ? g.final_cost[state[S.best_idx]] : 0.f, acoustic = 0.f;
        int f = n_frames - 1;
        int nw = 0;
        int *wout = P.words + (size_t)u * cfg.max_words;
        bool wovf = false;
        while (gid >= 0) {
          int2 rec = ws.arena[gid];
          unsigned arc = (unsigned)rec.y;
          if (arc != kArcNone) {
            int4 a = arc < NE ? g.earc[arc] : g.parc[arc - NE];
            graph += __int_as_float(a.z);
            if (arc < NE) {
              acoustic -= P.loglikes[(size_t)(P.ll_row0[u] + f) * P.ld + a.y];
              f--;
            }
            if (a.w != 0) {
              if (nw < cfg.max_words)
                wout[cfg.max_words - 1 - nw] = a.w;
              else
                wovf = true;
              nw++;
            }
          }
          gid = rec.x;
        }
        if (wovf) {
          status |= 8;
          nw = cfg.max_words;
        }
        for (int i = 0; i < nw; i++) wout[i] = wout[cfg.max_words - nw + i];
        S.n_next = nw;
        P.cost[2 * u] = graph;
        P.cost[2 * u + 1] = acoustic;
      }
      __syncthreads();
      if (!(status & 4)) n_words = S.n_next;
      status = __syncthreads_or(status);
    }
    // per-utterance counters
    {
      unsigned long long v = cnt_arcs;
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      __shared__ unsigned long long red[kMaxNW];
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (tid == 0) {
        unsigned long long arcs = 0;
        for (int w = 0; w < NW; w++) arcs += red[w];
        P.counters[4 * (size_t)u + 0] = cnt_tokens;
        P.counters[4 * (size_t)u + 1] = arcs;
        P.counters[4 * (size_t)u + 2] = cnt_created;
        P.counters[4 * (size_t)u + 3] = (unsigned long long)arena_n;
        P.n_words[u] = (status & ~8) ? -1 : n_words;
        P.status[u] = status | info;
        if (status & ~8) {
          P.cost[2 * u] = 0.f;
          P.cost[2 * u + 1] = 0.f;
        }
      }
      __syncthreads();
    }
    if (status & 1)
      wipe_tables();
    else {
      clear_table(0);
      clear_table(1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// a20: lattice-beam pruning of the recorded lattice, PruneForwardLinksFinal (:376-458) on the last time and
// PruneForwardLinks (:299-370) on every earlier one, back to front -- the state FinalizeDecoding (:625-640)
// leaves behind.  extra_cost[token] = min over its links of extra_cost[next] + ((tot_cost + acoustic + graph)
// - next.tot_cost), the reference's float expression; a link is excised when that exceeds lattice_beam, a
// token when no link (and no final cost) survives.  Extra costs are >= 0, so their bit patterns order as
// unsigned integers and one atomicMin per link replaces the reference's per-token loop; links inside one time
// (epsilon arcs) are relaxed to the fix point.  The periodic PruneActiveTokens (:506-533) of the reference only
// removes links this final pass removes as well (its extra costs are lower bounds of the final ones).
// Then the survivors are renumbered and written as compact arcs (GetRawLattice :106-189: olabel, graph cost,
// acoustic cost without the per-frame offset) behind a global cursor.
constexpr int kPruneWin = 4096;   // tokens of one time whose extra costs fit the shared-memory window
constexpr int kPrunePos = 2048;   // staged link-position entries (utterances up to 1022 decoded frames)
__global__ void __launch_bounds__(kMaxNT) lattice_prune_kernel(const __grid_constant__ DecodeParams P, float lattice_beam,
                                                               LatticeHeader *headers, LatticeArc *arcs, int arcs_cap,
                                                               int *cursor) {
  __shared__ Shared S;
  __shared__ int s_changed, s_base, s_nsurv, s_nfin, s_maybe;
  const int tid = threadIdx.x;
  const int u = blockIdx.x;
  if (tid == 0) s_nsurv = s_maybe = 0;
  const DevGraph &g = P.g;
  const int T = P.n_frames[u];
  const float kInf = __int_as_float(0x7f800000);
  const unsigned NE = g.num_earcs;
  if (T <= 0 || P.n_words[u] < 0 || P.lat.tok_base[(size_t)u * (P.lat.max_t + 2) + T + 1] < 0) {
    if (tid == 0) headers[u] = LatticeHeader{0, 0, 0, 0, 0, {0, 0, 0}};
    return;
  }
  const int2 *ltok = P.lat.tok + (size_t)u * P.lat.tok_cap;
  unsigned *extra = reinterpret_cast<unsigned *>(P.lat.extra + (size_t)u * P.lat.tok_cap);
  int *newid = P.lat.newid + (size_t)u * P.lat.tok_cap;
  int4 *llink = P.lat.link + (size_t)u * P.lat.link_cap;
  int4 *surv = P.lat.surv + (size_t)u * P.lat.surv_cap;
  const int *ltb = P.lat.tok_base + (size_t)u * (P.lat.max_t + 2);
  const int *lpos = P.lat.link_pos + (size_t)u * (2 * P.lat.max_t + 4);
  const float *loff = P.lat.cost_offset + (size_t)u * (P.lat.max_t + 1);
  const int ntok = ltb[T + 1];
  const int nlink = lpos[2 * (T + 1)];
  for (int i = tid; i < ntok; i += NT) extra[i] = 0x7f800000u;
  // The sweep is a chain of short, dependent steps (133 times x a few hundred links), so it is bound by latency,
  // not by bytes: the extra costs of the two times in flight live in shared memory windows (global memory when a
  // time holds more than kPruneWin tokens), the position tables are staged in shared memory, and every thread
  // fetches its links of the NEXT time step before it works on the current one.
  __shared__ unsigned win[2][kPruneWin];
  __shared__ int s_pos[kPrunePos], s_tb[kPrunePos / 2];
  const bool staged = 2 * T + 3 <= kPrunePos;
  if (staged) {
    for (int i = tid; i < 2 * T + 3; i += NT) s_pos[i] = lpos[i];
    for (int i = tid; i < T + 2; i += NT) s_tb[i] = ltb[i];
  }
  __syncthreads();
  const int *pos = staged ? s_pos : lpos;
  const int *tb = staged ? s_tb : ltb;
  // ---- last time: final costs (ComputeFinalCosts :536-577)
  const int f0 = tb[T], f1 = tb[T + 1];
  int anyf = 0;
  for (int i = f0 + tid; i < f1; i += NT) anyf |= g.final_cost[ltok[i].x] != kInf;
  anyf = __syncthreads_or(anyf);
  {
    float bv = kInf;
    int bi = 0x7fffffff;
    for (int i = f0 + tid; i < f1; i += NT) {
      const float fc = anyf ? g.final_cost[ltok[i].x] : 0.f;
      const float c = __fadd_rn(__int_as_float(ltok[i].y), fc);
      if (c < bv) {
        bv = c;
        bi = i;
      }
    }
    block_min(bv, bi, S);
  }
  const float final_best = S.best_cost;
  __syncthreads();
  int wsel = 0;
  unsigned *cur = (f1 - f0 <= kPruneWin) ? win[wsel] : extra + f0;  // extra costs of time t, indexed by token - tb[t]
  unsigned *nxt = nullptr;                                          // ... of time t + 1
  int b0 = f0, b1 = f1;
  for (int i = f0 + tid; i < f1; i += NT) {
    const float fc = anyf ? g.final_cost[ltok[i].x] : 0.f;
    float e = __fsub_rn(__fadd_rn(__int_as_float(ltok[i].y), fc), final_best);
    if (e < 0.f) e = 0.f;
    cur[i - f0] = (e > lattice_beam) ? 0x7f800000u : __float_as_uint(e);
  }
  // link_extra_cost = extra_cost[next] + slack (:330-332); the slack was stored by decode_kernel<true>
  auto ldx = [](const unsigned *p) { return __uint_as_float(*reinterpret_cast<const volatile unsigned *>(p)); };
  int4 pf_e = make_int4(0, 0, 0, 0), pf_p = make_int4(0, 0, 0, 0);
  {
    const int p0 = pos[2 * T + 1], p1 = pos[2 * T + 2];
    if (p0 + tid < p1) pf_p = llink[p0 + tid];
  }
  __syncthreads();
  for (int t = T; t >= 0; t--) {
    const int e0 = pos[2 * t + 2], e1 = t < T ? pos[2 * t + 3] : e0;  // emitting links t -> t + 1
    const int p0 = pos[2 * t + 1], p1 = pos[2 * t + 2];              // epsilon links inside t
    const int4 my_e = pf_e, my_p = pf_p;
    if (t > 0) {  // links of the next step, in flight while this one is processed
      const int ne0 = pos[2 * t], ne1 = pos[2 * t + 1], np0 = pos[2 * t - 1];
      if (ne0 + tid < ne1) pf_e = llink[ne0 + tid];
      if (np0 + tid < ne0) pf_p = llink[np0 + tid];
    }
    for (int i = e0 + tid; i < e1; i += NT) {
      const int4 l = i < e0 + NT ? my_e : llink[i];
      const float le = __fadd_rn(ldx(nxt + (l.y - b1)), __int_as_float(l.w));
      if (l.z >= 0 && !(le > lattice_beam)) atomicMin(cur + (l.x - b0), __float_as_uint(fmaxf(le, 0.f)));
    }
    if (tid == 0) s_changed = 0;
    __syncthreads();
    while (p1 > p0) {
      int ch = 0;
      for (int i = p0 + tid; i < p1; i += NT) {
        const int4 l = i < p0 + NT ? my_p : llink[i];
        const float le = __fadd_rn(ldx(cur + (l.y - b0)), __int_as_float(l.w));
        if (!(le > lattice_beam)) {
          const unsigned v = __float_as_uint(fmaxf(le, 0.f));
          if (v < atomicMin(cur + (l.x - b0), v)) ch = 1;
        }
      }
      if (ch) s_changed = 1;
      __syncthreads();
      const int again = s_changed;
      __syncthreads();
      if (!again) break;
      if (tid == 0) s_changed = 0;
      __syncthreads();
    }
    // the extras of time t and t + 1 are final: the links that stay inside the beam (2 % on a grammar graph) are
    // copied to the utterance's survivor list, straight from the registers they were relaxed from; nothing is
    // written back to the rest, and no later pass reads the full link array again
    for (int i = e0 + tid; i < e1; i += NT) {
      const int4 l = i < e0 + NT ? my_e : llink[i];
      if (!(__fadd_rn(ldx(nxt + (l.y - b1)), __int_as_float(l.w)) > lattice_beam)) {
        if (l.z < 0) {  // a "maybe" link that would survive: the lattice depends on the reference's visiting order
          s_maybe = 1;
          continue;
        }
        const int k = atomicAdd(&s_nsurv, 1);
        if (k < P.lat.surv_cap) surv[k] = make_int4(l.x, l.y, l.z, t);
      }
    }
    for (int i = p0 + tid; i < p1; i += NT) {
      const int4 l = i < p0 + NT ? my_p : llink[i];
      if (!(__fadd_rn(ldx(cur + (l.y - b0)), __int_as_float(l.w)) > lattice_beam)) {
        const int k = atomicAdd(&s_nsurv, 1);
        if (k < P.lat.surv_cap) surv[k] = make_int4(l.x, l.y, l.z, t);
      }
    }
    if (cur != extra + b0)
      for (int i = tid; i < b1 - b0; i += NT) extra[b0 + i] = cur[i];
    __syncthreads();  // every reader of the t + 1 window is done before it is recycled
    if (t > 0) {  // time t - 1 becomes the current one
      nxt = cur;
      b1 = b0;
      b0 = tb[t - 1];
      wsel ^= 1;
      cur = (b1 - b0 <= kPruneWin) ? win[wsel] : extra + b0;
      if (cur != extra + b0)
        for (int i = tid; i < b1 - b0; i += NT) cur[i] = 0x7f800000u;
    }
    __syncthreads();
  }
  // ---- renumber the surviving tokens
  unsigned n_nodes = 0;
  for (int b0 = 0; b0 < ntok; b0 += NT) {
    const int i = b0 + tid;
    const unsigned alive = (i < ntok && extra[i] != 0x7f800000u) ? 1u : 0u;
    unsigned total;
    const unsigned ex = block_excl_scan(alive, S, &total);
    if (i < ntok) newid[i] = alive ? (int)(n_nodes + ex) : -1;
    n_nodes += total;
  }
  __syncthreads();
  // ---- reserve and write: survivors, then the final weights of the last time
  const int n_surv = s_nsurv;
  if (tid == 0) s_nfin = 0;
  __syncthreads();
  for (int i = f0 + tid; i < f1; i += NT) {
    const float fc = anyf ? g.final_cost[ltok[i].x] : 0.f;
    if (newid[i] >= 0 && fc != kInf) atomicAdd(&s_nfin, 1);
  }
  __syncthreads();
  const int n_out = n_surv + s_nfin;
  if (tid == 0) s_base = n_surv <= P.lat.surv_cap ? atomicAdd(cursor, n_out) : arcs_cap;
  __syncthreads();
  const int base = s_base;
  if (n_surv > P.lat.surv_cap || base + n_out > arcs_cap) {  // survivor list or output buffer too small: best path only
    if (tid == 0) headers[u] = LatticeHeader{0, 0, 0, 0, 0, {0, 0, 0}};
    return;
  }
  if (tid == 0) s_nfin = 0;
  __syncthreads();
  for (int k = tid; k < n_surv; k += NT) {
    const int4 l = surv[k];
    const bool emitting = (unsigned)l.z < NE;
    const int4 a = emitting ? g.earc[l.z] : g.parc[(unsigned)l.z - NE];
    float acoustic = 0.f;
    if (emitting) {  // GetRawLattice's acoustic cost is (offset - loglike) - offset, rounding included (:158-165)
      const float off = loff[l.w];
      acoustic = __fsub_rn(__fsub_rn(off, P.loglikes[(size_t)(P.ll_row0[u] + l.w) * P.ld + a.y]), off);
    }
    arcs[base + k] = LatticeArc{newid[l.x], newid[l.y], a.w, __int_as_float(a.z), acoustic};
  }
  for (int i = f0 + tid; i < f1; i += NT) {
    const float fc = anyf ? g.final_cost[ltok[i].x] : 0.f;
    if (newid[i] >= 0 && fc != kInf) arcs[base + n_surv + atomicAdd(&s_nfin, 1)] = LatticeArc{newid[i], -1, 0, fc, 0.f};
  }
  if (tid == 0) headers[u] = LatticeHeader{base, n_out, (int)n_nodes, 1, nlink, {s_maybe, 0, 0}};
}

void LaunchLatticePrune(const DecodeParams &p, float lattice_beam, LatticeHeader *headers, LatticeArc *arcs,
                        int arcs_cap, int *cursor, cudaStream_t stream) {
  if (p.n_utts == 0) return;
  lattice_prune_kernel<<<p.n_utts, kMaxNT, 0, stream>>>(p, lattice_beam, headers, arcs, arcs_cap, cursor);
}

size_t DecodeSmemBytes(int slots) { return (size_t)slots * (2 * 8 + 15 * 4) + 64; }

void LaunchDecode(const DecodeParams &p, int n_lanes, cudaStream_t stream, bool lattice) {
  if (p.n_utts == 0) return;
  size_t smem = 0;
  if (p.cfg.smem_slots) {
    smem = DecodeSmemBytes(p.cfg.smem_slots);
    if (lattice)
      EnsureDynSmem(decode_kernel<true>, smem);
    else
      EnsureDynSmem(decode_kernel<false>, smem);
  }
  if (lattice)
    decode_kernel<true><<<n_lanes, DecodeCtaThreads(), smem, stream>>>(p);
  else
    decode_kernel<false><<<n_lanes, DecodeCtaThreads(), smem, stream>>>(p);
}

}  // namespace rs
