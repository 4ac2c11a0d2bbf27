"""CPU restatement of the reference's lattice decoder for the parity tests (rows a18-a22 of SURVEY 8).

TEST INFRASTRUCTURE ONLY -- nothing here is on the product path.  Only tests/, scripts/debug_*.py,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.  Pure-Python loops in float32: meant for
the tiny fixtures (a few hundred tokens per frame), where it finishes in seconds.

Follows kaldi/src/decoder/lattice-faster-decoder.cc step by step, in the order the reference visits tokens where
that order is observable (the transient next_cutoff of ProcessEmitting, the LIFO queue of ProcessNonemitting):
  InitDecoding :56-73, GetCutoff :644-711, ProcessEmitting :714-804, ProcessNonemitting :820-887,
  FindOrAddToken :252-293, PruneForwardLinksFinal :376-458, PruneForwardLinks :299-370, FinalizeDecoding :625-640,
  ComputeFinalCosts :536-577, GetRawLattice :106-189.
Pinned against the reference: tests/test_decoder_oracle.py compares the pruned state-level lattice with
`latgen-faster-mapped --determinize-lattice=false` (tests/golden/nbest_golden.npz, made from oracle/_ref).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

F = np.float32
INF = F(np.inf)


class ConstFst:
    """HCLG.fst as ConstFst<StdArc> (kaldi/openfst/src/include/fst/const-fst.h:192-232, lib/fst.cc:58-82)."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            b = f.read()
        p = 0

        def i32():
            nonlocal p
            v = struct.unpack_from("<i", b, p)[0]
            p += 4
            return v

        def s():
            nonlocal p
            n = i32()
            v = b[p:p + n].decode()
            p += n
            return v
        assert i32() == 2125659606, "not an OpenFst file"
        assert s() == "const", "only ConstFst is restated here"
        s()
        version, flags = i32(), i32()
        p += 8
        self.start, ns, na = struct.unpack_from("<qqq", b, p)
        p += 24
        assert not (flags & 3), "symbol tables inside the FST are not restated here"
        if version == 1:
            p += (-p) % 16
        st = np.frombuffer(b, dtype=[("final", "<f4"), ("pos", "<u4"), ("narcs", "<u4"), ("nieps", "<u4"), ("noeps", "<u4")], count=ns, offset=p)
        p += 20 * ns
        if version == 1:
            p += (-p) % 16
        arcs = np.frombuffer(b, dtype=[("ilabel", "<i4"), ("olabel", "<i4"), ("weight", "<f4"), ("nextstate", "<i4")], count=na, offset=p)
        self.final = st["final"].astype(np.float32)
        self.emit: List[List[Tuple[int, int, np.float32, int]]] = []
        self.eps: List[List[Tuple[int, int, np.float32, int]]] = []
        for i in range(ns):
            a = arcs[st["pos"][i]:st["pos"][i] + st["narcs"][i]]
            rows = [(int(x["ilabel"]), int(x["olabel"]), F(x["weight"]), int(x["nextstate"])) for x in a]
            self.emit.append([r for r in rows if r[0] != 0])
            self.eps.append([r for r in rows if r[0] == 0])


class _Tok:
    __slots__ = ("tot", "links", "extra", "state")

    def __init__(self, tot, state):
        self.tot = tot
        self.links: list = []      # (next token, ilabel, olabel, graph cost, acoustic cost)
        self.extra = F(0.0)
        self.state = state


def _get_cutoff(costs: np.ndarray, beam, max_active, min_active, beam_delta):
    """GetCutoff :644-711 -> (cur_cutoff, adaptive_beam, best index)."""
    best_i = int(np.argmin(costs))
    best = costs[best_i]
    beam_cutoff = F(best + beam)
    min_c, max_c = INF, INF
    tmp = costs
    if len(tmp) > max_active:
        max_c = np.partition(tmp, max_active)[max_active]
    if max_c < beam_cutoff:
        return max_c, F(F(max_c - best) + beam_delta), best_i
    if len(tmp) > min_active:
        if min_active == 0:
            min_c = best
        else:
            head = np.partition(tmp, max_active)[:max_active] if len(tmp) > max_active else tmp
            min_c = np.partition(head, min_active)[min_active]
    if min_c > beam_cutoff:
        return min_c, F(F(min_c - best) + beam_delta), best_i
    return beam_cutoff, F(beam), best_i


def decode(fst: ConstFst, loglikes: np.ndarray, tid2pdf: np.ndarray, beam: float = 24.0, max_active: int = 7000,
           min_active: int = 200, beam_delta: float = 0.5, lattice_beam: float = 8.0, prune_interval: int = 25,
           prune_scale: float = 0.1) -> Optional[dict]:
    """One utterance.  Returns None when no token survives, else
    {words, graph_cost, acoustic_cost, lattice: {src, dst, olabel, graph, acoustic, n_states}, tokens_per_frame}."""
    beam, beam_delta, lattice_beam = F(beam), F(beam_delta), F(lattice_beam)
    ll = np.asarray(loglikes, dtype=np.float32)
    T = ll.shape[0]
    frames: List[Dict[int, _Tok]] = [dict()]
    cost_offsets: List[np.float32] = []

    def nonemitting(toks: Dict[int, _Tok], cutoff):
        queue = [s for s in toks if fst.eps[s]]
        while queue:
            s = queue.pop()
            tok = toks[s]
            if tok.tot >= cutoff:
                continue
            tok.links = []
            for il, ol, w, ns in fst.eps[s]:
                tot = F(tok.tot + w)
                if tot < cutoff:
                    nt = toks.get(ns)
                    changed = False
                    if nt is None:
                        nt = toks[ns] = _Tok(tot, ns)
                        changed = True
                    elif nt.tot > tot:
                        nt.tot = tot
                        changed = True
                    tok.links.append((nt, 0, ol, w, F(0.0)))
                    if changed and fst.eps[ns]:
                        queue.append(ns)

    def prune_forward_links(toks: Dict[int, _Tok], delta) -> Tuple[bool, bool]:
        """PruneForwardLinks :299-370 -> (extra_costs_changed, links_pruned)."""
        any_changed = links_pruned = False
        changed = True
        while changed:
            changed = False
            for tok in toks.values():
                e = INF
                keep = []
                for link in tok.links:
                    nt = link[0]
                    le = F(nt.extra + F(F(F(tok.tot + link[4]) + link[3]) - nt.tot))
                    if le > lattice_beam:
                        links_pruned = True
                        continue
                    if le < 0:
                        le = F(0.0)
                    if le < e:
                        e = le
                    keep.append(link)
                tok.links = keep
                with np.errstate(invalid="ignore"):
                    if abs(F(e - tok.extra)) > delta:
                        changed = True
                tok.extra = e
            if changed:
                any_changed = True
        return any_changed, links_pruned

    must_links: List[bool] = [True]
    must_toks: List[bool] = [True]

    def prune_active_tokens(delta):
        """PruneActiveTokens :506-533 (called every prune_interval frames, delta = lattice_beam * prune_scale)."""
        cur = len(frames) - 1
        for f in range(cur - 1, -1, -1):
            if must_links[f]:
                ch, pr = prune_forward_links(frames[f], delta)
                if ch and f > 0:
                    must_links[f - 1] = True
                if pr:
                    must_toks[f] = True
                must_links[f] = False
            if f + 1 < cur and must_toks[f + 1]:
                frames[f + 1] = {s: tok for s, tok in frames[f + 1].items() if tok.extra != INF}
                must_toks[f + 1] = False

    frames[0][fst.start] = _Tok(F(0.0), fst.start)
    nonemitting(frames[0], beam)
    for t in range(T):
        if prune_interval > 0 and t % prune_interval == 0:
            prune_active_tokens(F(lattice_beam * F(prune_scale)))
        cur = frames[t]
        if not cur:
            return None
        states = list(cur)
        costs = np.array([cur[s].tot for s in states], dtype=np.float32)
        cur_cutoff, adaptive_beam, best_i = _get_cutoff(costs, beam, max_active, min_active, beam_delta)
        best_tok = cur[states[best_i]]
        cost_offset = F(-best_tok.tot)
        next_cutoff = INF
        for il, ol, w, ns in fst.emit[states[best_i]]:
            nw = F(F(F(w + cost_offset) - ll[t, tid2pdf[il]]) + best_tok.tot)
            if F(nw + adaptive_beam) < next_cutoff:
                next_cutoff = F(nw + adaptive_beam)
        cost_offsets.append(cost_offset)
        nxt: Dict[int, _Tok] = dict()
        frames.append(nxt)
        must_links.append(True)
        must_toks.append(True)
        for s in states:
            tok = cur[s]
            if tok.tot <= cur_cutoff:
                for il, ol, w, ns in fst.emit[s]:
                    ac = F(cost_offset - ll[t, tid2pdf[il]])
                    tot = F(F(tok.tot + ac) + w)
                    if tot >= next_cutoff:
                        continue
                    if F(tot + adaptive_beam) < next_cutoff:
                        next_cutoff = F(tot + adaptive_beam)
                    nt = nxt.get(ns)
                    if nt is None:
                        nt = nxt[ns] = _Tok(tot, ns)
                    elif nt.tot > tot:
                        nt.tot = tot
                    tok.links.append((nt, il, ol, w, ac))
        nonemitting(nxt, next_cutoff)
    last = frames[T]
    if not last:
        return None
    # ---- FinalizeDecoding: PruneForwardLinksFinal, then PruneForwardLinks + PruneTokensForFrame back to front
    anyf = any(fst.final[s] != INF for s in last)
    fin = {s: (fst.final[s] if anyf else F(0.0)) for s in last}
    final_best = min(F(tok.tot + fin[s]) for s, tok in last.items())

    def prune_links(toks: Dict[int, _Tok], final: bool):
        changed = True
        while changed:
            changed = False
            for s, tok in toks.items():
                e = F(F(tok.tot + fin[s]) - final_best) if final else INF
                keep = []
                for link in tok.links:
                    nt = link[0]
                    le = F(nt.extra + F(F(F(tok.tot + link[4]) + link[3]) - nt.tot))
                    if le > lattice_beam:
                        continue
                    if le < 0:
                        le = F(0.0)
                    if le < e:
                        e = le
                    keep.append(link)
                tok.links = keep
                if final and e > lattice_beam:
                    e = INF
                if e != tok.extra and not (np.isinf(e) and np.isinf(tok.extra)):
                    changed = True
                tok.extra = e
    prune_links(last, True)
    for t in range(T - 1, -1, -1):
        prune_links(frames[t], False)
    alive = [{s: tok for s, tok in fr.items() if tok.extra != INF} for fr in frames]
    # ---- GetRawLattice: states time-major, acoustic cost minus the frame's offset
    ids = {}
    for fr in alive:
        for tok in fr.values():
            ids[id(tok)] = len(ids)
    src, dst, olab, gr, ac, ilab = [], [], [], [], [], []
    for t, fr in enumerate(alive):
        for s, tok in fr.items():
            for nt, il, ol, w, a in tok.links:
                if id(nt) not in ids:
                    continue
                src.append(ids[id(tok)])
                dst.append(ids[id(nt)])
                olab.append(ol)
                ilab.append(il)
                gr.append(w)
                ac.append(F(a - cost_offsets[t]) if il != 0 else F(0.0))
            if t == T and fin[s] != INF:
                src.append(ids[id(tok)])
                dst.append(-1)
                olab.append(0)
                ilab.append(0)
                gr.append(fin[s])
                ac.append(F(0.0))
    lattice = dict(src=np.array(src, np.int32), dst=np.array(dst, np.int32), olabel=np.array(olab, np.int32),
                   graph=np.array(gr, np.float32), acoustic=np.array(ac, np.float32), n_states=len(ids),
                   ilabel=np.array(ilab, np.int32),
                   hclg_state=np.array([tok.state for fr in alive for tok in fr.values()], np.int32),
                   time=np.array([t for t, fr in enumerate(alive) for _ in fr], np.int32))
    return dict(lattice=lattice, tokens_per_frame=[len(fr) for fr in frames])
