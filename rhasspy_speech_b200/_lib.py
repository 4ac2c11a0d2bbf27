"""ctypes binding of librs_b200.so (C ABI in include/rs_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
Python or CPU fallback: if the library is missing, or CUDA is unavailable when a model is loaded,
the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Union, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librs_b200.so")

ERRLEN = 2048


class DecoderOpts(C.Structure):
    _fields_ = [("beam", C.c_float), ("max_active", C.c_int32), ("min_active", C.c_int32), ("lattice_beam", C.c_float),
                ("acoustic_scale", C.c_float), ("beam_delta", C.c_float), ("max_tokens_per_frame", C.c_int32),
                ("max_tokens_per_utt", C.c_int32), ("max_words", C.c_int32), ("num_lanes", C.c_int32),
                ("dither_seed", C.c_uint32), ("strict_fallback", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("n_utts", C.c_int32), ("n_hyp", C.POINTER(C.c_int32)), ("word_offset", C.POINTER(C.c_int32)),
                ("word_ids", C.POINTER(C.c_int32)), ("graph_cost", C.POINTER(C.c_float)),
                ("acoustic_cost", C.POINTER(C.c_float)), ("num_frames", C.POINTER(C.c_int32)),
                ("status", C.POINTER(C.c_int32)), ("hyp_offset", C.POINTER(C.c_int32)),
                ("hyp_word_offset", C.POINTER(C.c_int32)), ("hyp_word_ids", C.POINTER(C.c_int32)),
                ("hyp_graph_cost", C.POINTER(C.c_float)), ("hyp_acoustic_cost", C.POINTER(C.c_float))]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("feature_ms", C.c_float), ("nnet_ms", C.c_float), ("decode_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("audio_seconds", C.c_double),
                ("frames_decoded", C.c_uint64), ("tokens_expanded", C.c_uint64), ("arcs_visited", C.c_uint64),
                ("tokens_created", C.c_uint64), ("records_written", C.c_uint64), ("nnet_flops", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_int32),
                ("nnet_bytes", C.c_uint64), ("lattice_states", C.c_uint64), ("lattice_arcs", C.c_uint64),
                ("lattice_links_recorded", C.c_uint64), ("strict_utts", C.c_int32), ("strict_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/rs_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_ERR = [C.c_char_p, C.c_size_t]
SYMBOLS = [
    ("rs_decoder_opts_default", None, [C.POINTER(DecoderOpts)]),
    ("rs_device_count", C.c_int, []),
    ("rs_model_load", _P, [C.c_char_p, C.c_char_p, C.c_int] + _ERR),
    ("rs_model_free", None, [_P]),
    ("rs_model_info", C.c_int, [_P] + [C.POINTER(C.c_int32)] * 6),
    ("rs_graph_load", _P, [C.c_char_p, C.c_char_p, C.c_int] + _ERR),
    ("rs_graph_free", None, [_P]),
    ("rs_graph_info", C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    ("rs_graph_word", C.c_char_p, [_P, C.c_int32]),
    ("rs_decoder_create", _P, [_P, _P, C.POINTER(DecoderOpts)] + _ERR),
    ("rs_decoder_free", None, [_P]),
    ("rs_host_alloc", _P, [C.c_size_t] + _ERR),
    ("rs_host_free", None, [_P]),
    ("rs_fuzzy_load", _P, [C.c_char_p, C.c_char_p] + _ERR),
    ("rs_fuzzy_free", None, [_P]),
    ("rs_fuzzy_match", C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_float)] + _ERR),
    ("rs_fuzzy_word", C.c_char_p, [_P, C.c_int32]),
    ("rs_decoder_set_graph", C.c_int, [_P, _P] + _ERR),
    ("rs_decoder_set_nbest", C.c_int, [_P, C.c_int32, C.c_float] + _ERR),
    ("rs_decoder_set_staging_overlap", C.c_int, [_P, C.c_int32] + _ERR),
    ("rs_debug_lattice_nbest", C.c_int, [_P] * 5 + [C.c_int32, C.c_int32, C.c_int32, C.c_float, _P, _P, C.c_int32, _P]),
    ("rs_debug_read_matrix", C.c_int, [C.c_char_p, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)] + _ERR),
    ("rs_debug_strict_decode", C.c_int, [C.c_char_p, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.POINTER(DecoderOpts), C.c_int32,
                                        C.c_float, _P, _P, C.c_int32, _P, _P] + _ERR),
    ("rs_decode_pcm", C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.POINTER(Result))] + _ERR),
    ("rs_decode_wavs", C.c_int, [_P, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.POINTER(Result))] + _ERR),
    ("rs_decode_loglikes", C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.POINTER(Result))] + _ERR),
    ("rs_result_free", None, [C.POINTER(Result)]),
    ("rs_stream_open", _P, [_P] + _ERR),
    ("rs_stream_accept", C.c_int, [_P, C.c_void_p, C.c_int32] + _ERR),
    ("rs_stream_finish", C.c_int, [_P, C.POINTER(C.POINTER(Result))] + _ERR),
    ("rs_stream_close", None, [_P]),
    ("rs_streams_finish", C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.POINTER(Result))] + _ERR),
    ("rs_decoder_timings", C.c_int, [_P, C.POINTER(Timings)]),
    ("rs_debug_fetch", C.c_int, [_P, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)] + _ERR),
    ("rs_debug_gemm", C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_float)] + _ERR),
    ("rs_model_plan", C.c_char_p, [_P]),
    ("rs_model_check", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t] + _ERR),
    ("rs_graph_check", C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int64)] + _ERR),
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class RsError(RuntimeError):
    pass


def _check(ok: bool, err):
    if not ok:
        raise RsError(err.value.decode(errors="replace"))


def device_count() -> int:
    return int(load_library().rs_device_count())


def model_check(final_mdl: str, online_conf: str) -> str:
    """Host-only parse of the model artefacts; returns a summary and the compiled plan (no GPU needed)."""
    lib = load_library()
    out = C.create_string_buffer(1 << 16)
    err = C.create_string_buffer(ERRLEN)
    rc = lib.rs_model_check(os.fsencode(final_mdl), os.fsencode(online_conf), out, len(out), err, ERRLEN)
    _check(rc == 0, err)
    return out.value.decode()


def graph_check(hclg_fst: str, words_txt: Optional[str]) -> dict:
    lib = load_library()
    counts = (C.c_int64 * 6)()
    err = C.create_string_buffer(ERRLEN)
    rc = lib.rs_graph_check(os.fsencode(hclg_fst), os.fsencode(words_txt) if words_txt else None, counts, err, ERRLEN)
    _check(rc == 0, err)
    return dict(zip(("states", "emitting_arcs", "epsilon_arcs", "start", "final_states", "words"), [int(x) for x in counts]))


def debug_gemm(src: np.ndarray, w: np.ndarray, offsets: Sequence[int] = (0,), stride: int = 1,
               bias: Optional[np.ndarray] = None, relu: bool = False, path: int = 1, iters: int = 0, device: int = 0):
    """One TDNN-style layer through the affine-layer kernels (rs_debug_gemm); returns (out, ms_per_launch)."""
    lib = load_library()
    src = np.ascontiguousarray(src, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = src.shape
    n = w.shape[0]
    assert w.shape[1] == k * len(offsets)
    offs = (C.c_int * len(offsets))(*[int(o) for o in offsets])
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    m = max(rows // stride, 1)
    out = np.zeros((m, n), dtype=np.float32)
    ms = C.c_float()
    err = C.create_string_buffer(ERRLEN)
    rc = lib.rs_debug_gemm(device, src.ctypes.data, rows, k, offs, len(offsets), stride, w.ctypes.data, n,
                           None if b is None else b.ctypes.data, int(relu), path, iters, out.ctypes.data, C.byref(ms),
                           err, ERRLEN)
    _check(rc == 0, err)
    return out, ms.value


def lattice_nbest(src, dst, olabel, graph, acoustic, n_nodes: int, n: int, acoustic_scale: float = 1.0):
    """Host half of the n-best tail on a caller-provided state-level lattice (rs_debug_lattice_nbest):
    returns [(word ids, graph cost, acoustic cost), ...], best first."""
    lib = load_library()
    src, dst, olabel = [np.ascontiguousarray(a, dtype=np.int32) for a in (src, dst, olabel)]
    graph, acoustic = [np.ascontiguousarray(a, dtype=np.float32) for a in (graph, acoustic)]
    max_words = 4096 * n
    woff = np.zeros(n + 1, np.int32)
    wid = np.zeros(max_words, np.int32)
    cost = np.zeros(2 * n, np.float32)
    k = lib.rs_debug_lattice_nbest(src.ctypes.data, dst.ctypes.data, olabel.ctypes.data, graph.ctypes.data,
                                   acoustic.ctypes.data, len(src), n_nodes, n, acoustic_scale, woff.ctypes.data,
                                   wid.ctypes.data, max_words, cost.ctypes.data)
    if k < 0:
        raise RsError("rs_debug_lattice_nbest failed (%d)" % k)
    return [([int(x) for x in wid[woff[h]:woff[h + 1]]], float(cost[2 * h]), float(cost[2 * h + 1])) for h in range(k)]


def read_matrix(path: str) -> np.ndarray:
    """A Kaldi Matrix<float> file (FM / DM / CM / CM2 / CM3 / text) through the library's reader (rs_debug_read_matrix)."""
    lib = load_library()
    r, c = C.c_int32(), C.c_int32()
    err = C.create_string_buffer(ERRLEN)
    _check(lib.rs_debug_read_matrix(os.fsencode(path), None, C.byref(r), C.byref(c), err, ERRLEN) == 0, err)
    out = np.zeros((r.value, c.value), np.float32)
    if out.size:
        _check(lib.rs_debug_read_matrix(os.fsencode(path), out.ctypes.data, C.byref(r), C.byref(c), err, ERRLEN) == 0, err)
    return out


def strict_decode(hclg_fst: str, tid2pdf: np.ndarray, loglikes: np.ndarray, nbest: int = 1, acoustic_scale: float = 1.0, **opts):
    """The strict-order host decoder on one log-likelihood matrix (rs_debug_strict_decode, no GPU):
    returns ([(word ids, graph cost, acoustic cost), ...] best first, (lattice states, lattice arcs))."""
    lib = load_library()
    o = DecoderOpts()
    lib.rs_decoder_opts_default(C.byref(o))
    for k, v in opts.items():
        if not hasattr(o, k):
            raise TypeError("unknown decoder option " + k)
        setattr(o, k, v)
    ll = np.ascontiguousarray(loglikes, dtype=np.float32)
    t2p = np.ascontiguousarray(tid2pdf, dtype=np.int32)
    max_words = 4096 * nbest
    woff = np.zeros(nbest + 1, np.int32)
    wid = np.zeros(max_words, np.int32)
    cost = np.zeros(2 * nbest, np.float32)
    lat = np.zeros(2, np.int32)
    err = C.create_string_buffer(ERRLEN)
    k = lib.rs_debug_strict_decode(os.fsencode(hclg_fst), t2p.ctypes.data, len(t2p), ll.ctypes.data, ll.shape[0], ll.shape[1],
                                   C.byref(o), nbest, acoustic_scale, woff.ctypes.data, wid.ctypes.data, max_words,
                                   cost.ctypes.data, lat.ctypes.data, err, ERRLEN)
    _check(k >= 0, err)
    hyps = [([int(x) for x in wid[woff[h]:woff[h + 1]]], float(cost[2 * h]), float(cost[2 * h + 1])) for h in range(k)]
    return hyps, (int(lat[0]), int(lat[1]))


class PinnedAudio:
    """A batch of utterances back to back in page-locked memory (rs_host_alloc): `views[i]` are int16 arrays the
    caller fills (or that were filled from `utterances`); Decoder.decode_pcm(views) then copies the batch to the
    device straight from this block, without the staging memcpy a pageable buffer needs."""

    def __init__(self, lengths: Sequence[int]):
        self.lib = load_library()
        total = int(sum(int(x) for x in lengths))
        err = C.create_string_buffer(ERRLEN)
        self.ptr = self.lib.rs_host_alloc(max(total, 1) * 2, err, ERRLEN)
        _check(bool(self.ptr), err)
        buf = (C.c_int16 * max(total, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.int16)
        self.views: List[np.ndarray] = []
        o = 0
        addrs = []
        for n in lengths:
            self.views.append(self.array[o:o + int(n)])
            addrs.append(self.ptr + 2 * o)
            o += int(n)
        # the argument arrays of rs_decode_pcm, built once: Decoder.decode_pcm(pinned_audio) passes them as they are
        k = len(addrs)
        self._ptrs = (C.c_void_p * max(k, 1))(*addrs)
        self._ns = (C.c_int32 * max(k, 1))(*[int(x) for x in lengths])
        self._n = k

    @classmethod
    def from_utterances(cls, utterances: Sequence[np.ndarray]) -> "PinnedAudio":
        pa = cls([len(u) for u in utterances])
        for v, u in zip(pa.views, utterances):
            v[:] = np.asarray(u, dtype=np.int16)
        return pa

    def close(self):
        if getattr(self, "ptr", None):
            self.views, self.array = [], None
            self.lib.rs_host_free(self.ptr)
            self.ptr = None

    __del__ = close


def _address(a: np.ndarray) -> int:
    """Address of a contiguous array's data (from_buffer is ~2x cheaper than ndarray.ctypes.data)."""
    if a.size and a.flags.writeable:
        return C.addressof(C.c_char.from_buffer(a))
    return a.ctypes.data


class Hypotheses:
    """Python copy of an rs_result."""

    def __init__(self, r: Result):
        n = r.n_utts
        self.n_utts = n

        def arr(ptr, count):
            dt = np.dtype(ptr._type_)
            if count <= 0 or not ptr:
                return np.zeros(0, dtype=dt)
            # one memcpy out of the library's buffer (np.ctypeslib.as_array costs ~0.1 ms per call)
            return np.frombuffer(C.string_at(ptr, count * dt.itemsize), dtype=dt)

        off = arr(r.word_offset, n + 1) if n else np.zeros(1, np.int32)
        ids = arr(r.word_ids, int(off[-1]))
        self.n_hyp = arr(r.n_hyp, n)
        # Python lists once (tolist), then plain slices: ~3x cheaper per utterance than converting numpy slices
        ids_l, off_l, nh_l = ids.tolist(), off.tolist(), self.n_hyp.tolist()
        self.words: List[Optional[List[int]]] = [ids_l[off_l[u]:off_l[u + 1]] if nh_l[u] else None for u in range(n)]
        self.graph_cost = arr(r.graph_cost, n)
        self.acoustic_cost = arr(r.acoustic_cost, n)
        self.num_frames = arr(r.num_frames, n)
        self.status = arr(r.status, n)
        # every hypothesis, best first: nbest[u] = [(word ids, graph cost, acoustic cost), ...]; built on first use
        # (the single-best path of a 256-utterance batch should not pay for 256 tuples it never reads)
        self._nbest: Optional[List[List[tuple]]] = None
        self._hyp = None
        if n and self.n_hyp.max(initial=0) > 1:
            ho = arr(r.hyp_offset, n + 1)
            total = int(ho[-1])
            wo = arr(r.hyp_word_offset, total + 1)
            self._hyp = (ho, wo, arr(r.hyp_word_ids, int(wo[-1]) if total else 0), arr(r.hyp_graph_cost, total),
                         arr(r.hyp_acoustic_cost, total))

    @property
    def nbest(self) -> List[List[tuple]]:
        if self._nbest is None:
            out: List[List[tuple]] = [[] for _ in range(self.n_utts)]
            if self._hyp is not None:
                ho, wo, wid, gc, ac = [a.tolist() for a in self._hyp]
                for u in range(self.n_utts):
                    out[u] = [(wid[wo[h]:wo[h + 1]], gc[h], ac[h]) for h in range(ho[u], ho[u + 1])]
            else:
                gc, ac, nh = self.graph_cost.tolist(), self.acoustic_cost.tolist(), self.n_hyp.tolist()
                for u in range(self.n_utts):
                    if nh[u]:
                        out[u] = [(self.words[u], gc[u], ac[u])]
            self._nbest = out
        return self._nbest


class Fuzzy:
    """lang_dir/G.fuzzy.fst + words.txt resident on the host (rs_fuzzy_*): rhasspy's fuzzy matcher without OpenFst tools."""

    def __init__(self, g_fuzzy_fst: str, words_txt: str):
        self.lib = load_library()
        err = C.create_string_buffer(ERRLEN)
        self.h = self.lib.rs_fuzzy_load(os.fsencode(g_fuzzy_fst), os.fsencode(words_txt), err, ERRLEN)
        _check(bool(self.h), err)

    def match(self, hyps: Sequence[Sequence[int]]):
        """hyps: word ids per hypothesis, best first -> (output word ids, cost) or None when nothing matches."""
        flat = np.array([w for h in hyps for w in h], dtype=np.int32)
        off = np.zeros(len(hyps) + 1, np.int32)
        off[1:] = np.cumsum([len(h) for h in hyps])
        out = np.zeros(4096, np.int32)
        n, cost = C.c_int32(), C.c_float()
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_fuzzy_match(self.h, flat.ctypes.data, off.ctypes.data, len(hyps), out.ctypes.data, len(out),
                                     C.byref(n), C.byref(cost), err, ERRLEN)
        _check(rc >= 0, err)
        if rc == 1:
            return None
        return [int(x) for x in out[:n.value]], float(cost.value)

    def word(self, i: int) -> Optional[str]:
        w = self.lib.rs_fuzzy_word(self.h, i)
        return w.decode() if w is not None else None

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_fuzzy_free(self.h)
            self.h = None

    __del__ = close


class Model:
    def __init__(self, final_mdl: str, online_conf: str, device: int = 0):
        self.lib = load_library()
        err = C.create_string_buffer(ERRLEN)
        self.h = self.lib.rs_model_load(os.fsencode(final_mdl), os.fsencode(online_conf), device, err, ERRLEN)
        _check(bool(self.h), err)
        vals = [C.c_int32() for _ in range(6)]
        self.lib.rs_model_info(self.h, *[C.byref(v) for v in vals])
        (self.num_pdfs, self.frame_subsampling_factor, self.ivector_dim, self.feat_dim, self.left_context,
         self.right_context) = [v.value for v in vals]
        self.device = device

    def plan(self) -> str:
        return self.lib.rs_model_plan(self.h).decode()

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_model_free(self.h)
            self.h = None

    __del__ = close


class Graph:
    def __init__(self, hclg_fst: str, words_txt: Optional[str], device: int = 0):
        self.lib = load_library()
        err = C.create_string_buffer(ERRLEN)
        self.h = self.lib.rs_graph_load(os.fsencode(hclg_fst), os.fsencode(words_txt) if words_txt else None, device, err, ERRLEN)
        _check(bool(self.h), err)
        ns, na, nw = C.c_int32(), C.c_int64(), C.c_int32()
        self.lib.rs_graph_info(self.h, C.byref(ns), C.byref(na), C.byref(nw))
        self.num_states, self.num_arcs, self.num_words = ns.value, na.value, nw.value

    def word(self, i: int) -> Optional[str]:
        # symbols seen before come out of a dict (one ctypes call + decode per word is ~1.5 us; a transcript repeats
        # the same few hundred words); the table of a loaded graph never changes
        try:
            return self._word_cache[i]
        except (KeyError, AttributeError):
            w = self.lib.rs_graph_word(self.h, i)
            w = w.decode() if w is not None else None
            if not hasattr(self, "_word_cache"):
                self._word_cache = {}
            self._word_cache[i] = w
            return w

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_graph_free(self.h)
            self.h = None

    __del__ = close


class Decoder:
    def __init__(self, model: Model, graph: Graph, **opts):
        self.lib = load_library()
        self.model, self.graph = model, graph
        o = DecoderOpts()
        self.lib.rs_decoder_opts_default(C.byref(o))
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError("unknown decoder option " + k)
            setattr(o, k, v)
        self.opts = o
        err = C.create_string_buffer(ERRLEN)
        self.h = self.lib.rs_decoder_create(model.h, graph.h, C.byref(o), err, ERRLEN)
        _check(bool(self.h), err)

    def set_graph(self, graph: "Graph"):
        """Bind the decoder to another HCLG (rs_decoder_set_graph); the device workspace is kept."""
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_decoder_set_graph(self.h, graph.h, err, ERRLEN)
        _check(rc == 0, err)
        self.graph = graph

    def set_nbest(self, nbest: int = 1, acoustic_scale: float = 1.0):
        """lattice-to-nbest --n / --acoustic-scale for every later decode call (rs_decoder_set_nbest)."""
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_decoder_set_nbest(self.h, int(nbest), float(acoustic_scale), err, ERRLEN)
        _check(rc == 0, err)

    def set_staging_overlap(self, on: bool = True):
        """Copy stream + per-item MFCC launches (default) or one stream with separate stage times (rs_decoder_set_staging_overlap)."""
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_decoder_set_staging_overlap(self.h, 1 if on else 0, err, ERRLEN)
        _check(rc == 0, err)

    def _take(self, rc: int, res, err) -> Hypotheses:
        _check(rc == 0, err)
        try:
            return Hypotheses(res.contents)
        finally:
            self.lib.rs_result_free(res)

    def decode_pcm(self, pcm: "Union[Sequence[np.ndarray], PinnedAudio]") -> Hypotheses:
        if isinstance(pcm, PinnedAudio):        # the whole block, no per-utterance marshalling
            res = C.POINTER(Result)()
            err = C.create_string_buffer(ERRLEN)
            rc = self.lib.rs_decode_pcm(self.h, pcm._ptrs, pcm._ns, pcm._n, C.byref(res), err, ERRLEN)
            return self._take(rc, res, err)
        arrs = [np.ascontiguousarray(p, dtype=np.int16) for p in pcm]
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[_address(a) for a in arrs])
        ns = (C.c_int32 * max(n, 1))(*[a.size for a in arrs])
        res = C.POINTER(Result)()
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_decode_pcm(self.h, ptrs, ns, n, C.byref(res), err, ERRLEN)
        return self._take(rc, res, err)

    def decode_wavs(self, paths: Sequence[str]) -> Hypotheses:
        n = len(paths)
        arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
        res = C.POINTER(Result)()
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_decode_wavs(self.h, arr, n, C.byref(res), err, ERRLEN)
        return self._take(rc, res, err)

    def decode_loglikes(self, loglikes: Sequence[np.ndarray]) -> Hypotheses:
        arrs = [np.ascontiguousarray(m, dtype=np.float32) for m in loglikes]
        for a in arrs:
            if a.ndim != 2 or a.shape[1] != self.model.num_pdfs:
                raise ValueError("log-likelihood matrices must be [frames x %d]" % self.model.num_pdfs)
        n = len(arrs)
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
        ns = (C.c_int32 * max(n, 1))(*[a.shape[0] for a in arrs])
        res = C.POINTER(Result)()
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_decode_loglikes(self.h, ptrs, ns, n, C.byref(res), err, ERRLEN)
        return self._take(rc, res, err)

    def timings(self) -> dict:
        t = Timings()
        self.lib.rs_decoder_timings(self.h, C.byref(t))
        return t.as_dict()

    def fetch(self, what: int, utt: int) -> np.ndarray:
        """0 MFCC, 1 iVector, 2 log-likelihoods, 3 normalised MFCC, 4 LDA features of the last call."""
        r, c = C.c_int32(), C.c_int32()
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_debug_fetch(self.h, what, utt, None, C.byref(r), C.byref(c), err, ERRLEN)
        _check(rc == 0, err)
        out = np.zeros((r.value, c.value), dtype=np.float32)
        if out.size:
            rc = self.lib.rs_debug_fetch(self.h, what, utt, out.ctypes.data, C.byref(r), C.byref(c), err, ERRLEN)
            _check(rc == 0, err)
        return out

    def open_stream(self) -> "Stream":
        return Stream(self)

    def finish_streams(self, streams: Sequence["Stream"]) -> Hypotheses:
        n = len(streams)
        if any(not getattr(s, "h", None) for s in streams):
            raise RsError("finish_streams: a stream of the batch is closed")
        arr = (C.c_void_p * max(n, 1))(*[s.h for s in streams])
        res = C.POINTER(Result)()
        err = C.create_string_buffer(ERRLEN)
        rc = self.lib.rs_streams_finish(arr, n, C.byref(res), err, ERRLEN)
        return self._take(rc, res, err)

    def close(self):
        if getattr(self, "h", None):
            self.lib.rs_decoder_free(self.h)
            self.h = None

    __del__ = close


class Stream:
    def __init__(self, dec: Decoder):
        self.dec = dec
        err = C.create_string_buffer(ERRLEN)
        self.h = dec.lib.rs_stream_open(dec.h, err, ERRLEN)
        _check(bool(self.h), err)
        self._err = err         # reused by accept(): one call per 80 ms chunk and stream

    def accept(self, chunk: bytes):
        """Raw s16le bytes (any even length); the library copies them, so the bytes object is passed as is."""
        if not chunk:
            return
        if not isinstance(chunk, bytes):
            chunk = bytes(chunk)
        err = self._err
        rc = self.dec.lib.rs_stream_accept(self.h, chunk, len(chunk) // 2, err, ERRLEN)
        _check(rc == 0, err)

    def finish(self) -> Hypotheses:
        res = C.POINTER(Result)()
        err = C.create_string_buffer(ERRLEN)
        rc = self.dec.lib.rs_stream_finish(self.h, C.byref(res), err, ERRLEN)
        return self.dec._take(rc, res, err)

    def close(self):
        if getattr(self, "h", None):
            self.dec.lib.rs_stream_close(self.h)
            self.h = None

    __del__ = close
