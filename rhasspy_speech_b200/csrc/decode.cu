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
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "engine.h"

namespace rs {

// CTA size is a launch-time choice (128 .. 512 threads): the frontier of a grammar graph is a few
// hundred tokens, where a small CTA pays far less per __syncthreads than a large one
constexpr int kMaxNT = 512;
constexpr int kMaxNW = kMaxNT / 32;
#define NT ((int)blockDim.x)
#define NW ((int)(blockDim.x >> 5))
constexpr unsigned long long kEmptyVal = ~0ULL;
constexpr int kEmptyKey = -1;
constexpr unsigned kArcNone = 0xffffffffu;

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

__device__ __forceinline__ unsigned ord(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unord(unsigned o) {
  unsigned b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long pack(float cost, unsigned arc) {
  return ((unsigned long long)ord(cost) << 32) | arc;
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
    }
    __syncthreads();
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
    int cur = 0;
    int n_cur = 0;        // alive tokens of the current frame
    int base_cur = 0;     // arena index of the current frame's first token
    int arena_n = 0;

    // One pass of: epsilon closure of table `tb` under `cutoff`, compaction of the alive entries
    // into tok_state/tok_cost[tb], traceback records.  `tprev` is the table of the previous frame.
    auto close_and_finalize = [&](int tb, int tprev, float cutoff, int base_prev, int base_new) -> int {
      Table T{ws.hkey[tb], ws.hval[tb], ws.hidx[tb], ws.ins_list[tb], &S.n_ins[tb]};
      // ---- ProcessNonemitting (:820-887): frontier = alive tokens whose state has epsilon arcs
      int fcur = 0;
      if (tid == 0) S.frontier_n[0] = S.frontier_n[1] = 0;
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
      if (base_new + n_new > cfg.arena_cap) return -2;
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
      }
      __syncthreads();
      tick(5);
      return n_new;
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
      int r = close_and_finalize(0, 1, cfg.beam, 0, 0);
      if (r < 0) {
        status |= (r == -1 ? 1 : 2);
        n_cur = 0;
      } else {
        n_cur = r;
        arena_n = r;
        cnt_created += r;
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
          info |= 16;
          adaptive_beam = __fadd_rn(__fsub_rn(max_active_cutoff, best), cfg.beam_delta);
          cur_cutoff = max_active_cutoff;
        } else {
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
        if (lane == 0) S.nc_ord = ord(m);
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
        volatile unsigned *nc = &S.nc_ord;
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
          const float ncv = unord(*nc);
          if (tot >= ncv) continue;
          const float cand = __fadd_rn(tot, adaptive_beam);
          if (cand < ncv) atomicMin(&S.nc_ord, ord(cand));
          if (*(volatile int *)&S.overflow) break;
          int s2 = insert_slot(T, arc.x, mask, identity, tok_cap, &S.overflow);
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
      int r = close_and_finalize(nxt, cur, next_cutoff, base_cur, base_new);
      if (r < 0) {
        status |= (r == -1 ? 1 : 2);
        break;
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

size_t DecodeSmemBytes(int slots) { return (size_t)slots * (2 * 8 + 15 * 4) + 64; }

void LaunchDecode(const DecodeParams &p, int n_lanes, cudaStream_t stream) {
  if (p.n_utts == 0) return;
  size_t smem = 0;
  if (p.cfg.smem_slots) {
    smem = DecodeSmemBytes(p.cfg.smem_slots);
    static size_t configured = 48 * 1024;
    if (smem > configured) {
      cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      configured = smem;
    }
  }
  decode_kernel<<<n_lanes, DecodeCtaThreads(), smem, stream>>>(p);
}

}  // namespace rs
