// Stage (iii) for small graphs (a grammar HCLG: up to 1000 states): token passing with the reference's token ORDER
// reproduced on the device, so that the result is LatticeFasterDecoder's unconditionally -- including the
// order-dependent parts of its pruning (kaldi/src/decoder/lattice-faster-decoder.cc):
//   * ProcessEmitting (:714-804) admits an arc when its cost is below the TRANSIENT next_cutoff, which tightens as
//     tokens are visited (:780-787).  Visited in list order, the cutoff an arc meets is
//         min(seed from the best token's arcs (:744-759), min over all EARLIER arcs of cost + adaptive_beam),
//     an exclusive prefix-min over the flat (token, arc) enumeration -- a parallel scan, given the order.
//   * The order is the token hash's list order (kaldi/src/util/hash-list-inl.h:156-194): buckets by first occupation,
//     bucket = state % hash_size with hash_size >= 1000 (:44, PossiblyResizeHash :219-225 only grows it).  A graph of at
//     most 1000 states therefore has one state per bucket and the list order is the order of FIRST INSERTION:
//     emitting-phase tokens by the flat position of the first admitted arc that reached them, then the tokens
//     ProcessNonemitting (:820-887) creates, in the order its LIFO queue reaches them.
//   * FindOrAddToken (:252-293) keeps the first of equally cheap arrivals: an atomicMin over (cost, flat position).
//   * Tokens beyond the frame's final cutoff ("extras") stay in the list exactly as in the reference: they count in
//     GetCutoff (:644-711), are skipped by ProcessNonemitting, and may be final on the last frame.
// Every per-frame table is addressed by state id and lives in shared memory; one CTA walks one utterance.
// Larger graphs use decode.cu (same cutoff values, extras dropped, order-sensitive frames detected and flagged).
#include <cuda_pipeline.h>

#include <cfloat>
#include <cstdio>

#include "decode_common.cuh"
#include "engine.h"
#include "smem_attr.h"

namespace rs {

namespace {

constexpr int kNT = 256;
constexpr int kNW = kNT / 32;
constexpr int kSlots = 1024;         // >= kSmallMaxStates
constexpr int kItems = kSlots / kNT; // list items per thread in the blocked scans
constexpr unsigned kEpsBase = 0x20000000u;   // first-insertion keys of the epsilon phase start here
constexpr unsigned kEpsTag = 0x40000000u;    // low word of a packed value: epsilon arc id | kEpsTag
constexpr unsigned kNoFirst = 0xffffffffu;

struct SmallShared {
  unsigned long long nval[kSlots];  // next frontier by state: (ordered cost, flat position | epsilon arc tag)
  unsigned first[kSlots];           // by state: key of the first insertion (flat position, or kEpsBase + flat epsilon arc)
  int lstate[2][kSlots];            // token lists (current / next) in the reference's list order
  float lcost[2][kSlots];
  unsigned pfx[kSlots + 1];         // emitting out-degree prefix over the current list
  unsigned pfx2[kSlots + 1];        // epsilon out-degree prefix over the emitting-phase tokens, newest first
  int nidx[kSlots];                 // by state: index in the next list
  unsigned adm[kSmallMaxEarcs / 32];   // admitted emitting arcs of the frame, by flat position
  unsigned adm2[kSmallMaxParcs / 32];  // admitted epsilon arcs
  unsigned wsum[2][kNW];
  float wmin[2][kNW];
  int wmin_i[2][kNW];
  float seed, kth;
  int n_links, lat_overflow, flag, n_new;
};

struct Ctx {
  SmallShared &S;
  int parity = 0;
  __device__ explicit Ctx(SmallShared &s) : S(s) {}
  // exclusive prefix sum over the block in thread order; one barrier (the scratch alternates between two buffers)
  __device__ unsigned scan_sum(unsigned v, unsigned *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    unsigned *ws = S.wsum[parity];
    parity ^= 1;
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kNW; w++) {
      const unsigned t = ws[w];
      if (w < warp) base += t;
      tot += t;
    }
    *total = tot;
    return base + x - v;
  }
  // exclusive prefix min over the block in thread order (identity +inf), and the block minimum
  __device__ float scan_min(float v, float *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float kInf = __int_as_float(0x7f800000);
    float x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x = fminf(x, y);
    }
    float *ws = S.wmin[parity];
    parity ^= 1;
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    float base = kInf, tot = kInf;
#pragma unroll
    for (int w = 0; w < kNW; w++) {
      const float t = ws[w];
      if (w < warp) base = fminf(base, t);
      tot = fminf(tot, t);
    }
    *total = tot;
    float ex = __shfl_up_sync(0xffffffffu, x, 1);  // inclusive -> exclusive inside the warp
    if (lane == 0) ex = kInf;
    return fminf(base, ex);
  }
  // block minimum of (v, i), ties towards the smaller i
  __device__ void min_idx(float v, int i, float *bv, int *bi) {
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
    float *wv = S.wmin[parity];
    int *wi = S.wmin_i[parity];
    parity ^= 1;
    if (lane == 0) {
      wv[warp] = v;
      wi[warp] = i;
    }
    __syncthreads();
    float rv = wv[0];
    int ri = wi[0];
#pragma unroll
    for (int w = 1; w < kNW; w++) {
      const float ov = wv[w];
      const int oi = wi[w];
      if (ov < rv || (ov == rv && oi < ri)) {
        rv = ov;
        ri = oi;
      }
    }
    *bv = rv;
    *bi = ri;
  }
};

// largest i in [0, n) with pfx[i] <= a  (pfx ascending, pfx[0] = 0, a < pfx[n])
__device__ __forceinline__ int locate(const unsigned *pfx, int n, unsigned a) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pfx[mid] <= a) lo = mid; else hi = mid;
  }
  return lo;
}

template <bool kLat>
__global__ void __launch_bounds__(kNT) decode_small_kernel(const __grid_constant__ DecodeParams P) {
  __shared__ SmallShared S;
  Ctx X(S);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const DevGraph &g = P.g;
  const DecodeConfig &cfg = P.cfg;
  const unsigned NE = g.num_earcs;
  const float kInf = __int_as_float(0x7f800000);
  const int u = blockIdx.x;
  const int n_frames = P.n_frames[u];
  if (n_frames <= 0) {
    if (tid == 0) {
      P.n_words[u] = -1;
      P.status[u] = 0;
      P.cost[2 * u] = P.cost[2 * u + 1] = 0.f;
      for (int c = 0; c < 4; c++) P.counters[4 * (size_t)u + c] = 0ULL;
    }
    return;
  }
  for (int i = tid; i < kSlots; i += kNT) {
    S.nval[i] = kEmptyVal;
    S.first[i] = kNoFirst;
  }
  // Dynamic shared memory (layout fixed by SmallSmemPlan on the host): the graph's CSR offsets, its arcs when they fit,
  // the destination state of every flat arc of the frame, and two staged log-likelihood rows (the row of frame f + 1 is
  // copied asynchronously while frame f is processed).  Every per-arc access of the frame loop is then shared memory.
  extern __shared__ __align__(16) unsigned char dsm[];
  const int S1 = g.num_states + 1;
  unsigned *ebeg = reinterpret_cast<unsigned *>(dsm), *pbeg = ebeg + S1;
  unsigned char *q = dsm + (((size_t)2 * S1 * 4 + 15) & ~(size_t)15);
  const int4 *earc = g.earc, *parc = g.parc;
  const int *psrc = g.p_src;
  for (int i = tid; i < S1; i += kNT) {
    ebeg[i] = g.e_begin[i];
    pbeg[i] = g.p_begin[i];
  }
  if (cfg.small_cache_arcs) {
    int4 *e_s = reinterpret_cast<int4 *>(q);
    q += (size_t)NE * 16;
    int4 *p_s = reinterpret_cast<int4 *>(q);
    q += (size_t)g.num_parcs * 16;
    int *ps_s = reinterpret_cast<int *>(q);
    q += ((size_t)g.num_parcs * 4 + 15) & ~(size_t)15;
    for (unsigned i = tid; i < NE; i += kNT) e_s[i] = g.earc[i];
    for (unsigned i = tid; i < g.num_parcs; i += kNT) {
      p_s[i] = g.parc[i];
      ps_s[i] = g.p_src[i];
    }
    earc = e_s;
    parc = p_s;
    psrc = ps_s;
  }
  unsigned short *adst = reinterpret_cast<unsigned short *>(q);
  q += ((size_t)NE * 2 + 15) & ~(size_t)15;
  float *llbuf[2] = {nullptr, nullptr};
  if (cfg.small_ll_stage) {
    llbuf[0] = reinterpret_cast<float *>(q);
    llbuf[1] = llbuf[0] + P.ld;
  }
  const float *ll_rows = P.loglikes + (size_t)P.ll_row0[u] * P.ld;
  auto stage_row = [&](int frame) {  // asynchronous copy of one log-likelihood row (ld is a multiple of 4 floats)
    if (!cfg.small_ll_stage || frame >= n_frames) return;
    const float4 *src = reinterpret_cast<const float4 *>(ll_rows + (size_t)frame * P.ld);
    float4 *dst = reinterpret_cast<float4 *>(llbuf[frame & 1]);
    for (int i = tid; i < P.ld / 4; i += kNT) __pipeline_memcpy_async(dst + i, src + i, 16);
    __pipeline_commit();
  };
  stage_row(0);
  if (tid == 0) {
    S.n_links = 0;
    S.lat_overflow = 0;
    S.flag = 0;
  }
  int2 *arena = P.small_arena + P.small_arena_off[u];
  const long long arena_cap = P.small_arena_off[u + 1] - P.small_arena_off[u];
  int2 *ltok = nullptr;
  int4 *llink = nullptr;
  int *ltb = nullptr, *lpos = nullptr;
  float *loff = nullptr;
  bool lat_dead = false;
  if constexpr (kLat) {
    ltok = P.lat.tok + (size_t)u * P.lat.tok_cap;
    llink = P.lat.link + (size_t)u * P.lat.link_cap;
    ltb = P.lat.tok_base + (size_t)u * (P.lat.max_t + 2);
    lpos = P.lat.link_pos + (size_t)u * (2 * P.lat.max_t + 4);
    loff = P.lat.cost_offset + (size_t)u * (P.lat.max_t + 1);
  }
  auto add_link = [&](int src, int dst, unsigned arc, float slack) {
    const int li = atomicAdd(&S.n_links, 1);
    if (li < P.lat.link_cap)
      llink[li] = make_int4(src, dst, (int)arc, __float_as_int(slack));
    else
      S.lat_overflow = 1;
  };
  unsigned long long cnt_tokens = 0, cnt_arcs = 0, cnt_created = 0;  // block-uniform
  // optional phase timing (RS_B200_DECODE_PROFILE=1): thread 0 of block 0 accumulates SM clocks per phase
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ph_last = clock64();
  auto tick = [&](int k) {
    if (cfg.profile && tid == 0 && blockIdx.x == 0) {
      const long long now = clock64();
      ph[k] += now - ph_last;
      ph_last = now;
    }
  };
  int status = 0;
  int cur = 0;          // which list is the current one
  int n_cur = 0;
  int base_cur = 0;     // arena index of the current time's first token
  int arena_n = 0;
  __syncthreads();

  // ProcessNonemitting (:820-887) over the emitting-phase tokens lstate[nx][0 .. n_emit) under `cutoff`, then the new
  // time's list is final: costs, index by state, traceback records.  Returns the length of the list, < 0 on error.
  auto close_and_finalize = [&](int nx, int n_emit, float cutoff, int base_new) -> int {
    int *ls = S.lstate[nx];
    int n_new = n_emit;
    unsigned n_eps = 0;
    if (g.num_parcs) {
      if (g.eps_flat) {
        // No epsilon arc leads to a state with epsilon arcs: the queue holds the emitting-phase tokens with epsilon
        // arcs, popped newest first, and nothing is ever pushed again.  The creation order of the new tokens is the
        // flat order of the (token newest-first, arc) enumeration: the same two passes as the emitting phase.
        unsigned deg[kItems], sum = 0;
#pragma unroll
        for (int k = 0; k < kItems; k++) {
          const int r = tid * kItems + k;  // r-th newest
          deg[k] = 0;
          if (r < n_emit) {
            const int st = ls[n_emit - 1 - r];
            const float c = unord((unsigned)(S.nval[st] >> 32));
            if (c < cutoff) deg[k] = pbeg[st + 1] - pbeg[st];
          }
          sum += deg[k];
        }
        unsigned ex = X.scan_sum(sum, &n_eps);
#pragma unroll
        for (int k = 0; k < kItems; k++) {
          const int r = tid * kItems + k;
          if (r < n_emit) S.pfx2[r] = ex;
          ex += deg[k];
        }
        if (tid == 0) S.pfx2[n_emit] = n_eps;
        __syncthreads();
        if (n_eps > (unsigned)kSmallMaxParcs) n_eps = 0, status |= 1;  // cannot happen: one token per state
        for (unsigned e0 = 0; e0 < n_eps; e0 += kNT) {
          const unsigned e = e0 + tid;
          bool ok = false;
          if (e < n_eps) {
            const int r = locate(S.pfx2, n_emit, e);
            const int st = ls[n_emit - 1 - r];
            const unsigned pa = pbeg[st] + (e - S.pfx2[r]);
            const int4 arc = parc[pa];
            const float c = unord((unsigned)(S.nval[st] >> 32));
            const float tot = __fadd_rn(c, __int_as_float(arc.z));
            if (tot < cutoff) {
              ok = true;
              atomicMin(&S.nval[arc.x], pack(tot, kEpsTag | pa));
              atomicMin(&S.first[arc.x], kEpsBase + e);
            }
          }
          const unsigned b = __ballot_sync(0xffffffffu, ok);
          if (lane == 0) S.adm2[(e0 >> 5) + warp] = b;
        }
        __syncthreads();
        for (unsigned e0 = 0; e0 < n_eps; e0 += kNT) {
          const unsigned e = e0 + tid;
          int ns = -1;
          if (e < n_eps && ((S.adm2[e >> 5] >> (e & 31)) & 1u)) {
            const int r = locate(S.pfx2, n_emit, e);
            const int st = ls[n_emit - 1 - r];
            const int4 arc = parc[pbeg[st] + (e - S.pfx2[r])];
            if (S.first[arc.x] == kEpsBase + e) ns = arc.x;
          }
          unsigned total;
          const unsigned pos = X.scan_sum(ns >= 0 ? 1u : 0u, &total);
          if (ns >= 0) ls[n_new + pos] = ns;
          n_new += (int)total;
        }
        __syncthreads();
      } else {
        // General epsilon structure: the reference's LIFO queue, replayed by one thread (queue in pfx2).
        if (tid == 0) {
          int *q = reinterpret_cast<int *>(S.pfx2);
          const int qcap = kSlots + 1;
          int nq = 0, nn = n_emit;
          for (int j = 0; j < n_emit; j++)
            if (pbeg[ls[j] + 1] > pbeg[ls[j]]) q[nq++] = ls[j];
          unsigned visited = 0;
          while (nq > 0) {
            const int st = q[--nq];
            const float c = unord((unsigned)(S.nval[st] >> 32));
            if (c >= cutoff) continue;
            for (unsigned pa = pbeg[st]; pa < pbeg[st + 1]; pa++) {
              const int4 arc = parc[pa];
              visited++;
              const float tot = __fadd_rn(c, __int_as_float(arc.z));
              if (tot < cutoff) {
                const unsigned long long old = S.nval[arc.x];
                bool changed = false;
                if (old == kEmptyVal) {
                  S.nval[arc.x] = pack(tot, kEpsTag | pa);
                  S.first[arc.x] = kEpsBase;
                  ls[nn++] = arc.x;
                  changed = true;
                } else if (unord((unsigned)(old >> 32)) > tot) {
                  S.nval[arc.x] = pack(tot, kEpsTag | pa);
                  changed = true;
                }
                if (changed && pbeg[arc.x + 1] > pbeg[arc.x]) {
                  if (nq < qcap) q[nq++] = arc.x; else S.flag = 1;
                }
              }
            }
          }
          S.n_new = nn;
          S.kth = __uint_as_float(visited);
        }
        __syncthreads();
        n_new = S.n_new;
        n_eps = __float_as_uint(S.kth);
        __syncthreads();
      }
    }
    cnt_arcs += n_eps;
    tick(4);
    if ((long long)base_new + n_new > arena_cap) return -2;
    for (int i = tid; i < n_new; i += kNT) S.nidx[ls[i]] = i;
    __syncthreads();
    const int *lp = S.lstate[nx ^ 1];  // the previous time's list (the sources of the emitting arcs)
    for (int i = tid; i < n_new; i += kNT) {
      const int st = ls[i];
      const unsigned long long v = S.nval[st];
      const float c = unord((unsigned)(v >> 32));
      const unsigned low = (unsigned)(v & 0xffffffffULL);
      S.lcost[nx][i] = c;
      int prev = -1;
      unsigned arc = kArcNone;
      if (low == kArcNone) {
      } else if (low & kEpsTag) {
        const unsigned pa = low & ~kEpsTag;
        prev = base_new + S.nidx[psrc[pa]];
        arc = NE + pa;
      } else {
        const int lo = locate(S.pfx, n_cur, low);
        prev = base_cur + lo;
        arc = ebeg[lp[lo]] + (low - S.pfx[lo]);
      }
      arena[base_new + i] = make_int2(prev, (int)arc);
      if constexpr (kLat)
        if (!lat_dead && base_new + n_new <= P.lat.tok_cap) ltok[base_new + i] = make_int2(st, __float_as_int(c));
      S.nval[st] = kEmptyVal;
      S.first[st] = kNoFirst;
    }
    if constexpr (kLat)
      if (base_new + n_new > P.lat.tok_cap) lat_dead = true;
    __syncthreads();
    return n_new;
  };
  // lattice mode: the epsilon links of the finalised time -- every epsilon arc below the cutoff out of a token below
  // the cutoff (ProcessNonemitting regenerates a token's links whenever its cost changes, :858-884)
  auto eps_links = [&](int nx, float cutoff, int base_new, int n_new) {
    for (int i = tid; i < n_new; i += kNT) {
      const int st = S.lstate[nx][i];
      const float c = S.lcost[nx][i];
      if (!(c < cutoff)) continue;
      for (unsigned a = pbeg[st]; a < pbeg[st + 1]; a++) {
        const int4 arc = parc[a];
        const float tot = __fadd_rn(c, __int_as_float(arc.z));
        if (tot < cutoff) {
          const int d = S.nidx[arc.x];
          const float slack = __fsub_rn(tot, S.lcost[nx][d]);
          if (!(slack > cfg.lattice_beam)) add_link(base_new + i, base_new + d, NE + a, slack);
        }
      }
    }
    __syncthreads();
  };

  // ---- InitDecoding (:56-73)
  if (tid == 0) {
    S.lstate[0][0] = g.start;
    S.nval[g.start] = pack(0.f, kArcNone);
    S.first[g.start] = 0;
  }
  __syncthreads();
  {
    const int r = close_and_finalize(0, 1, cfg.beam, 0);
    if (r < 0) {
      status |= 2;
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

  for (int frame = 0; frame < n_frames && status == 0; frame++) {
    if (n_cur == 0) {
      status |= 4;
      break;
    }
    const int nx = cur ^ 1;
    const int *state = S.lstate[cur];
    const float *cost = S.lcost[cur];
    // this frame's row was staged during the previous frame (or before the loop); start on the next one
    const float *ll = ll_rows + (size_t)frame * P.ld;
    if (cfg.small_ll_stage) {
      __pipeline_wait_prior(0);
      __syncthreads();
      ll = llbuf[frame & 1];
      stage_row(frame + 1);
    }
    // ---- GetCutoff (:644-711) over the whole list, extras included
    float best;
    int best_idx;
    {
      float bv = kInf;
      int bi = 0x7fffffff;
      for (int i = tid; i < n_cur; i += kNT) {
        const float c = cost[i];
        if (c < bv) {
          bv = c;
          bi = i;
        }
      }
      X.min_idx(bv, bi, &best, &best_idx);
    }
    // value nth_element leaves at index k: the token with exactly k others before it in (cost, index) order
    auto kth = [&](int k) -> float {
      for (int i = tid; i < n_cur; i += kNT) {
        const float c = cost[i];
        int r = 0;
        for (int j = 0; j < n_cur; j++) {
          const float o = cost[j];
          r += (o < c || (o == c && j < i)) ? 1 : 0;
        }
        if (r == k) S.kth = c;
      }
      __syncthreads();
      const float v = S.kth;
      __syncthreads();
      return v;
    };
    const float beam_cutoff = __fadd_rn(best, cfg.beam);
    float cur_cutoff, adaptive_beam;
    {
      float max_active_cutoff = kInf, min_active_cutoff = kInf;
      if (n_cur > cfg.max_active) max_active_cutoff = kth(cfg.max_active);
      if (max_active_cutoff < beam_cutoff) {
        adaptive_beam = __fadd_rn(__fsub_rn(max_active_cutoff, best), cfg.beam_delta);
        cur_cutoff = max_active_cutoff;
      } else {
        if (n_cur > cfg.min_active) {
          if (cfg.min_active == 0) {
            min_active_cutoff = best;
          } else {
            unsigned inside = 0, total;
            for (int i = tid; i < n_cur; i += kNT) inside += cost[i] <= beam_cutoff ? 1u : 0u;
            X.scan_sum(inside, &total);
            min_active_cutoff = (int)total > cfg.min_active ? beam_cutoff : kth(cfg.min_active);
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
    if (warp == 0) {  // next_cutoff seeded from the best token's arcs (:744-759)
      const int st = state[best_idx];
      float m = kInf;
      for (unsigned a = ebeg[st] + lane; a < ebeg[st + 1]; a += 32) {
        const int4 arc = earc[a];
        const float nw = __fadd_rn(__fsub_rn(__fadd_rn(__int_as_float(arc.z), cost_offset), ll[arc.y]), best);
        m = fminf(m, __fadd_rn(nw, adaptive_beam));
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) S.seed = m;
    }
    unsigned n_arcs;
    {
      unsigned deg[kItems], sum = 0;
#pragma unroll
      for (int k = 0; k < kItems; k++) {
        const int i = tid * kItems + k;
        deg[k] = 0;
        if (i < n_cur && cost[i] <= cur_cutoff) {
          const int st = state[i];
          deg[k] = ebeg[st + 1] - ebeg[st];
        }
        sum += deg[k];
      }
      unsigned ex = X.scan_sum(sum, &n_arcs);
#pragma unroll
      for (int k = 0; k < kItems; k++) {
        const int i = tid * kItems + k;
        if (i < n_cur) S.pfx[i] = ex;
        ex += deg[k];
      }
      if (tid == 0) S.pfx[n_cur] = n_arcs;
    }
    __syncthreads();
    tick(1);
    // pass 1: every flat arc against the cutoff it meets in list order; winners by (cost, flat position)
    float run = S.seed;
    for (unsigned a0 = 0; a0 < n_arcs; a0 += kNT) {
      const unsigned a = a0 + tid;
      float tot = kInf, cand = kInf;
      int ns = 0;
      if (a < n_arcs) {
        const int lo = locate(S.pfx, n_cur, a);
        const int4 arc = earc[ebeg[state[lo]] + (a - S.pfx[lo])];
        const float ac = __fsub_rn(cost_offset, ll[arc.y]);
        tot = __fadd_rn(__fadd_rn(cost[lo], ac), __int_as_float(arc.z));
        cand = __fadd_rn(tot, adaptive_beam);
        ns = arc.x;
        adst[a] = (unsigned short)ns;
      }
      float chunk_min;
      const float met = fminf(run, X.scan_min(cand, &chunk_min));
      const bool ok = a < n_arcs && tot < met;
      if (ok) {
        atomicMin(&S.nval[ns], pack(tot, a));
        atomicMin(&S.first[ns], a);
      }
      const unsigned b = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) S.adm[(a0 >> 5) + warp] = b;
      run = fminf(run, chunk_min);
    }
    const float next_cutoff = run;
    __syncthreads();
    tick(2);
    // pass 2: the states in the order of their first insertion
    int n_emit = 0;
    for (unsigned a0 = 0; a0 < n_arcs; a0 += kNT) {
      const unsigned a = a0 + tid;
      int ns = -1;
      if (a < n_arcs && ((S.adm[a >> 5] >> (a & 31)) & 1u)) {
        const int d = adst[a];
        if (S.first[d] == a) ns = d;
      }
      unsigned total;
      const unsigned pos = X.scan_sum(ns >= 0 ? 1u : 0u, &total);
      if (ns >= 0) S.lstate[nx][n_emit + pos] = ns;
      n_emit += (int)total;
    }
    __syncthreads();
    tick(3);
    cnt_tokens += n_cur;
    cnt_arcs += n_arcs;
    const int base_new = base_cur + n_cur;
    const int r = close_and_finalize(nx, n_emit, next_cutoff, base_new);
    tick(5);
    if (r < 0) {
      status |= 2;
      break;
    }
    if constexpr (kLat) if (!lat_dead) {
      // forward links time frame -> frame + 1: every admitted arc (the reference links each arc it admits, also into a
      // token another arc reaches more cheaply); links beyond the lattice beam of their destination are dropped here
      // already (PruneForwardLinks excises them whatever happens later: link_extra_cost >= the slack, :330-337)
      for (unsigned a = tid; a < n_arcs; a += kNT) {
        if (!((S.adm[a >> 5] >> (a & 31)) & 1u)) continue;
        const int lo = locate(S.pfx, n_cur, a);
        const unsigned ai = ebeg[state[lo]] + (a - S.pfx[lo]);
        const int4 arc = earc[ai];
        const float ac = __fsub_rn(cost_offset, ll[arc.y]);
        const float tot = __fadd_rn(__fadd_rn(cost[lo], ac), __int_as_float(arc.z));
        const int d = S.nidx[arc.x];
        const float slack = __fsub_rn(tot, S.lcost[nx][d]);
        if (!(slack > cfg.lattice_beam)) add_link(base_cur + lo, base_new + d, ai, slack);
      }
      __syncthreads();
      if (tid == 0) {
        lpos[2 * (frame + 1) + 1] = S.n_links;
        ltb[frame + 1] = base_new;
        loff[frame] = cost_offset;
      }
      __syncthreads();
      eps_links(nx, next_cutoff, base_new, r);
      if (tid == 0) lpos[2 * (frame + 2)] = S.n_links;
      if (S.lat_overflow) lat_dead = true;
    }
    base_cur = base_new;
    n_cur = r;
    arena_n = base_new + r;
    cnt_created += r;
    cur = nx;
  }
  if (cfg.small_ll_stage) __pipeline_wait_prior(0);
  if constexpr (kLat) {
    if (tid == 0) ltb[n_frames + 1] = lat_dead ? -1 : arena_n;
  }
  if (cfg.profile && tid == 0 && blockIdx.x == 0)
    printf("decode_small phases (clocks, utt %d, %d frames): cutoff %lld seed+prefix %lld pass1 %lld pass2 %lld epsilon %lld finalize %lld\n", u,
           n_frames, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5]);
  // ---- best path (lattice-faster-online-decoder.cc:78-173)
  int n_words = -1;
  if (status == 0 && n_cur == 0) status |= 4;
  if (status == 0) {
    const float *cost = S.lcost[cur];
    const int *state = S.lstate[cur];
    int anyf = 0;
    for (int i = tid; i < n_cur; i += kNT) anyf |= g.final_cost[state[i]] != kInf;
    anyf = __syncthreads_or(anyf);
    float bv = kInf;
    int bi = 0x7fffffff;
    for (int i = tid; i < n_cur; i += kNT) {
      float c = cost[i];
      if (anyf) {
        const float f = g.final_cost[state[i]];
        c = f != kInf ? __fadd_rn(c, f) : kInf;
      }
      if (c < bv) {
        bv = c;
        bi = i;
      }
    }
    float best;
    int best_idx;
    X.min_idx(bv, bi, &best, &best_idx);
    if (best_idx == 0x7fffffff) {
      status |= 4;
    } else {
      if (tid == 0) {
        int gid = base_cur + best_idx;
        float graph = anyf ? g.final_cost[state[best_idx]] : 0.f, acoustic = 0.f;
        int f = n_frames - 1;
        int nw = 0;
        int *wout = P.words + (size_t)u * cfg.max_words;
        bool wovf = false;
        while (gid >= 0) {
          const int2 rec = arena[gid];
          const unsigned arc = (unsigned)rec.y;
          if (arc != kArcNone) {
            const int4 a = arc < NE ? earc[arc] : parc[arc - NE];
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
        if (wovf) nw = cfg.max_words;
        for (int i = 0; i < nw; i++) wout[i] = wout[cfg.max_words - nw + i];
        S.n_new = nw;
        S.flag |= wovf ? 8 : 0;
        P.cost[2 * u] = graph;
        P.cost[2 * u + 1] = acoustic;
      }
      __syncthreads();
      n_words = S.n_new;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // S.flag bit 0: the epsilon queue of the one-thread replay overflowed -> not exact, let the host decide (bit 16)
    const int info = (S.flag & 1) ? 16 : 0;
    status |= S.flag & 8;
    P.counters[4 * (size_t)u + 0] = cnt_tokens;
    P.counters[4 * (size_t)u + 1] = cnt_arcs;
    P.counters[4 * (size_t)u + 2] = cnt_created;
    P.counters[4 * (size_t)u + 3] = (unsigned long long)arena_n;
    P.n_words[u] = (status & ~8) ? -1 : n_words;
    P.status[u] = status | info;
    if (status & ~8) P.cost[2 * u] = P.cost[2 * u + 1] = 0.f;
  }
}

}  // namespace

bool DecodeSmallSupports(const DevGraph &g) {
  return g.num_states <= kSmallMaxStates && g.num_earcs <= (unsigned)kSmallMaxEarcs && g.num_parcs <= (unsigned)kSmallMaxParcs;
}

// dynamic shared memory of decode_small_kernel: CSR offsets | [arcs] | flat-arc destinations | [two log-likelihood rows];
// the optional parts are dropped (arcs first) when they do not fit next to the static tables
static size_t SmallSmemPlan(const DecodeParams &p, int *cache_arcs, int *ll_stage) {
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t base = up16((size_t)2 * (p.g.num_states + 1) * 4) + up16((size_t)p.g.num_earcs * 2);
  const size_t arcs = (size_t)p.g.num_earcs * 16 + (size_t)p.g.num_parcs * 16 + up16((size_t)p.g.num_parcs * 4);
  const size_t rows = (size_t)2 * p.ld * 4;
  const size_t budget = 150 * 1024;  // next to ~43 KB of static tables; leaves room for a second CTA on small graphs
  *ll_stage = (p.ld % 4 == 0 && base + rows <= budget) ? 1 : 0;
  *cache_arcs = base + (*ll_stage ? rows : 0) + arcs <= budget ? 1 : 0;
  return base + (*ll_stage ? rows : 0) + (*cache_arcs ? arcs : 0) + 16;
}

void LaunchDecodeSmall(const DecodeParams &p_in, cudaStream_t stream, bool lattice) {
  if (p_in.n_utts == 0) return;
  DecodeParams p = p_in;
  const size_t smem = SmallSmemPlan(p, &p.cfg.small_cache_arcs, &p.cfg.small_ll_stage);
  if (lattice)
    EnsureDynSmem(decode_small_kernel<true>, smem);
  else
    EnsureDynSmem(decode_small_kernel<false>, smem);
  if (lattice)
    decode_small_kernel<true><<<p.n_utts, kNT, smem, stream>>>(p);
  else
    decode_small_kernel<false><<<p.n_utts, kNT, smem, stream>>>(p);
}

}  // namespace rs
