"""Seeded synthetic model / graph fixtures in Kaldi's on-disk formats.

The real ``en_US-zamia`` artefacts are a download and are not available offline
(SURVEY.md, fact 3), so every parity test and the benchmark run on artefacts written by
this module.  Both sides -- the reference Kaldi binaries in ``oracle/_ref`` and the
B200 library -- consume *the same files*, laid out exactly as rhasspy-speech lays out a
trained model (reference ``rhasspy_speech/transcribe_wav.py:43-57``,
``kaldi/egs/wsj/s5/steps/online/nnet3/prepare_online_decoding.sh:122-209``):

    <model_dir>/model/model/final.mdl                      TransitionModel + nnet3 AmNnetSimple
    <model_dir>/model/online/conf/{online,mfcc,ivector_extractor,splice,online_cmvn}.conf
    <model_dir>/model/online/ivector_extractor/final.{mat,dubm,ie}, global_cmvn.stats
    <graph_dir>/HCLG.fst  (OpenFst ConstFst<StdArc>)       <graph_dir>/words.txt

The network follows the shape of the in-tree TDNN-F recipe
(``kaldi/egs/wsj/s5/local/chain/tuning/run_tdnn_1g.sh:182-208``): fixed-affine "lda" over
Append(-1,0,1,ReplaceIndex(ivector,t,0)), a relu-batchnorm layer, tdnnf layers
(linear TdnnComponent -> affine TdnnComponent -> ReLU -> BatchNorm -> dropout -> NoOp with a
Sum(Scale(0.66, prev), .) bypass), prefinal + output.  Graphs are built in the
``--reorder=true`` layout of mkgraph (forward transition into a state that carries the
self-loop), with optional silence, epsilon hops and ARPA-style back-off arcs.

This is a data generator only: nothing here is on the decode path.
"""
from __future__ import annotations

import math
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# Kaldi stream writer (binary "\0B" or text), see kaldi/src/base/io-funcs{,-inl}.h


class KaldiWriter:
    def __init__(self, binary: bool = True):
        self.binary = binary
        self.parts: List[bytes] = [b"\0B"] if binary else []

    def tok(self, t: str):
        self.parts.append(t.encode() + b" ")

    def nl(self):
        if not self.binary:
            self.parts.append(b"\n")

    def i32(self, v: int):
        if self.binary:
            self.parts.append(b"\x04" + struct.pack("<i", int(v)))
        else:
            self.parts.append(b"%d " % int(v))

    def f32(self, v: float):
        if self.binary:
            self.parts.append(b"\x04" + struct.pack("<f", float(v)))
        else:
            self.parts.append(repr(float(np.float32(v))).encode() + b" ")

    def f64(self, v: float):
        if self.binary:
            self.parts.append(b"\x08" + struct.pack("<d", float(v)))
        else:
            self.parts.append(repr(float(v)).encode() + b" ")

    def boolean(self, v: bool):
        self.parts.append((b"T" if v else b"F") + (b"" if self.binary else b" "))

    def intvec(self, v: Sequence[int]):
        if self.binary:
            a = np.asarray(v, dtype="<i4")
            self.parts.append(b"\x04" + struct.pack("<i", a.size) + a.tobytes())
        else:
            self.parts.append(b"[ " + b" ".join(b"%d" % int(x) for x in v) + b" ]\n")

    def vec(self, v: np.ndarray, double: bool = False):
        v = np.asarray(v, dtype="<f8" if double else "<f4").reshape(-1)
        if self.binary:
            self.parts.append((b"DV " if double else b"FV ") + b"\x04" + struct.pack("<i", v.size) + v.tobytes())
        else:
            self.parts.append(b" [ " + b" ".join(repr(float(x)).encode() for x in v) + b" ]\n")

    def mat(self, m: np.ndarray, double: bool = False):
        m = np.asarray(m, dtype="<f8" if double else "<f4")
        if m.size == 0:
            m = m.reshape(0, 0)
        assert m.ndim == 2
        if self.binary:
            self.parts.append((b"DM " if double else b"FM ") + b"\x04" + struct.pack("<i", m.shape[0])
                              + b"\x04" + struct.pack("<i", m.shape[1]) + np.ascontiguousarray(m).tobytes())
        else:
            if m.shape[0] == 0:
                self.parts.append(b" [ ]\n")
                return
            rows = [b"  " + b" ".join(repr(float(x)).encode() for x in r) for r in m]
            self.parts.append(b" [\n" + b"\n".join(rows) + b" ]\n")

    def spmat(self, m: np.ndarray, double: bool = True):
        """Packed lower triangle, row by row (kaldi/src/matrix/packed-matrix.cc)."""
        m = np.asarray(m)
        n = m.shape[0]
        il = np.tril_indices(n)
        packed = np.asarray(m[il], dtype="<f8" if double else "<f4")
        if self.binary:
            self.parts.append((b"DP " if double else b"FP ") + b"\x04" + struct.pack("<i", n) + packed.tobytes())
        else:
            out = [b" ["]
            k = 0
            for r in range(n):
                out.append(b"\n" + b" ".join(repr(float(x)).encode() for x in packed[k:k + r + 1]))
                k += r + 1
            out.append(b" ]\n")
            self.parts.append(b"".join(out))

    def raw(self, b: bytes):
        self.parts.append(b)

    def save(self, path: str):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "wb") as f:
            f.write(b"".join(self.parts))


# --------------------------------------------------------------------------------------


@dataclass
class SynthSpec:
    """Dimensions of one synthetic model + graph."""
    name: str = "tiny"
    seed: int = 1
    binary: bool = True
    # MFCC (kaldi/egs/wsj/s5/conf/mfcc_hires.conf)
    num_ceps: int = 40
    num_mel_bins: int = 40
    low_freq: float = 20.0
    high_freq: float = -400.0
    dither: float = 0.0                 # parity runs use dither=0 (SURVEY "hard parts")
    # iVector extractor
    ivector_dim: int = 30
    num_gauss: int = 32
    lda_dim: int = 40
    lda_bias: bool = False              # final.mat with the extra offset column
    splice_left: int = 3
    splice_right: int = 3
    num_gselect: int = 5
    min_post: float = 0.025
    posterior_scale: float = 0.1
    max_count: float = 100.0
    ivector_period: int = 10
    nnet_cmvn: bool = False             # --cmvn-config on the nnet input as well
    # acoustic model
    chain: bool = True                  # 1-state chain topology + frame_subsampling_factor 3
    frame_subsampling_factor: int = 3
    hidden_dim: int = 64
    bottleneck_dim: int = 16
    tdnnf_strides: Tuple[int, ...] = (1, 0, 3, 3)
    prefinal_small: int = 24
    num_phones: int = 20                # non-silence phones; phone 1 is SIL
    variants_per_phone: int = 3         # context-dependent variants -> pdfs
    output_scale: float = 1.5           # spread of the pseudo log-likelihoods
    priors: bool = False                # non-empty <Priors> (subtracts log prior)
    log_softmax: bool = False
    # graph
    graph: str = "grammar"              # "grammar" | "arpa"
    sentences: Optional[List[str]] = None
    vocab_size: int = 300               # arpa
    bigrams_per_word: int = 8           # arpa
    eps_hops: int = 1                   # extra epsilon hops on back-off arcs
    min_phones: int = 2
    max_phones: int = 6


TINY = SynthSpec()
ZAMIA_LIKE = SynthSpec(name="zamia_like", seed=7, ivector_dim=100, num_gauss=512, hidden_dim=1024,
                       bottleneck_dim=128, tdnnf_strides=(1, 1, 1, 0, 3, 3, 3, 3, 3, 3, 3, 3),
                       prefinal_small=192, num_phones=42, variants_per_phone=36, output_scale=1.5)

# sentences of the reference's tests/en_US-zamia fixture (file stems, "_" -> " ")
EN_US_SENTENCES = [
    "how cold is it", "how hot is it", "is it cold", "is it hot", "is the bedroom light on",
    "is the garage door closed", "is the garage door open", "is the garage light on",
    "is the living room lamp on", "make the bedroom light blue", "make the bedroom light green",
    "make the bedroom light red", "make the garage light blue", "make the garage light green",
    "make the garage light red", "make the living room lamp blue", "make the living room lamp green",
    "make the living room lamp red", "set the bedroom light to blue", "set the bedroom light to green",
    "set the bedroom light to red", "set the garage light to blue", "set the garage light to green",
    "set the garage light to red", "set the living room lamp to blue", "set the living room lamp to green",
    "set the living room lamp to red", "tell me the time", "turn off the bedroom light",
    "turn off the garage light", "turn off the living room lamp", "turn on the bedroom light",
    "turn on the garage light", "turn on the living room lamp", "what is the temperature",
    "what time is it", "whats the temperature", "whats the time",
]


# measured on the reference's tests/en_US-zamia WAVs with mfcc_hires (int16-scale samples)
_MFCC_STD = [36, 28, 19, 24, 24, 21, 24, 23, 20, 20, 17, 17, 16, 14, 16, 9.5, 9.4, 8, 5.4, 4.7, 3.2, 2.4, 0.8, 0.5,
             1.6, 2.2, 2.9, 3.6, 4.4, 4, 4.6, 4.2, 4.3, 4.4, 3.9, 4.1, 3.4, 3.1, 3.2, 2.6]


def mfcc_stats(dim: int) -> Tuple[np.ndarray, np.ndarray]:
    """(mean, std) profile used to scale the synthetic LDA / CMVN so activations are O(1)."""
    std = np.interp(np.linspace(0, len(_MFCC_STD) - 1, dim), np.arange(len(_MFCC_STD)), _MFCC_STD)
    mean = np.zeros(dim)
    mean[0] = 94.0
    return mean, std


# --------------------------------------------------------------------------------------
# transition model


@dataclass
class HmmInfo:
    """Everything the graph builder needs to know about the transition model."""
    # per (phone, variant): list over hmm states of (self_tid, fwd_tid, self_cost, fwd_cost)
    states: Dict[Tuple[int, int], List[Tuple[int, int, float, float]]] = field(default_factory=dict)
    num_pdfs: int = 0
    num_tids: int = 0
    tid2pdf: Optional[np.ndarray] = None


def _write_transition_model(w: KaldiWriter, spec: SynthSpec, rng: np.random.Generator) -> HmmInfo:
    nph = spec.num_phones + 1  # + SIL (phone 1)
    phones = list(range(1, nph + 1))
    if spec.chain:
        # kaldi/src/hmm/hmm-topology.h: chain "1-state" topology with separate forward/self-loop pdfs
        entry = [dict(fwd=0, slf=1, trans=[(0, 0.5), (1, 0.5)]), dict(fwd=-1, slf=-1, trans=[])]
        nstates = 1
    else:
        entry = [dict(fwd=k, slf=k, trans=[(k, 0.75), (k + 1, 0.25)]) for k in range(3)] + \
                [dict(fwd=-1, slf=-1, trans=[])]
        nstates = 3
    w.tok("<TransitionModel>"); w.nl()
    w.tok("<Topology>")
    if w.binary:
        w.intvec(phones)
        w.intvec([-1] + [0] * nph)
        if spec.chain:
            w.i32(-1)
        w.i32(1)
        w.i32(len(entry))
        for st in entry:
            w.i32(st["fwd"])
            if spec.chain:
                w.i32(st["slf"])
            w.i32(len(st["trans"]))
            for d, p in st["trans"]:
                w.i32(d); w.f32(p)
    else:
        w.nl(); w.tok("<TopologyEntry>"); w.nl(); w.tok("<ForPhones>"); w.nl()
        w.raw((" ".join(str(p) for p in phones) + "\n").encode())
        w.tok("</ForPhones>"); w.nl()
        for j, st in enumerate(entry):
            w.tok("<State>"); w.i32(j)
            if st["fwd"] != -1:
                if spec.chain:
                    w.tok("<ForwardPdfClass>"); w.i32(st["fwd"]); w.tok("<SelfLoopPdfClass>"); w.i32(st["slf"])
                else:
                    w.tok("<PdfClass>"); w.i32(st["fwd"])
            for d, p in st["trans"]:
                w.tok("<Transition>"); w.i32(d); w.f32(p)
            w.tok("</State>"); w.nl()
        w.tok("</TopologyEntry>"); w.nl()
    w.tok("</Topology>"); w.nl()

    # tuples, sorted (phone, hmm_state, forward_pdf, self_loop_pdf); SIL has one variant
    tuples = []
    pdf = 0
    variant_pdfs: Dict[Tuple[int, int], List[Tuple[int, int]]] = {}
    for ph in phones:
        nvar = 1 if ph == 1 else spec.variants_per_phone
        for v in range(nvar):
            per_state = []
            for s in range(nstates):
                if spec.chain:
                    per_state.append((pdf, pdf + 1)); pdf += 2
                else:
                    per_state.append((pdf, pdf)); pdf += 1
            variant_pdfs[(ph, v)] = per_state
    for ph in phones:
        nvar = 1 if ph == 1 else spec.variants_per_phone
        for s in range(nstates):
            for v in range(nvar):
                f, sl = variant_pdfs[(ph, v)][s]
                tuples.append((ph, s, f, sl, v))
    tuples.sort(key=lambda t: t[:4])
    w.tok("<Tuples>" if spec.chain else "<Triples>"); w.i32(len(tuples)); w.nl()
    for ph, s, f, sl, _ in tuples:
        w.i32(ph); w.i32(s); w.i32(f)
        if spec.chain:
            w.i32(sl)
        w.nl()
    w.tok("</Tuples>" if spec.chain else "</Triples>"); w.nl()

    # transition ids are 1-based, one per topology transition per tuple
    # (kaldi/src/hmm/transition-model.cc:144-177)
    info = HmmInfo()
    log_probs = [0.0]
    tid2pdf = [0]
    tid = 1
    for ph, s, f, sl, v in tuples:
        p_self = float(rng.uniform(0.55, 0.9)) if not spec.chain else float(rng.uniform(0.35, 0.65))
        p_fwd = 1.0 - p_self
        self_tid, fwd_tid = tid, tid + 1
        tid += 2
        log_probs += [math.log(p_self), math.log(p_fwd)]
        tid2pdf += [sl, f]
        info.states.setdefault((ph, v), [None] * nstates)[s] = (
            self_tid, fwd_tid, float(np.float32(-math.log(p_self))), float(np.float32(-math.log(p_fwd))))
    info.num_pdfs = pdf
    info.num_tids = tid - 1
    info.tid2pdf = np.asarray(tid2pdf, dtype=np.int32)
    w.tok("<LogProbs>"); w.nl()
    w.vec(np.asarray(log_probs, dtype=np.float32))
    w.tok("</LogProbs>"); w.nl()
    w.tok("</TransitionModel>"); w.nl()
    return info


# --------------------------------------------------------------------------------------
# nnet3: parameters are drawn first, the BatchNorm statistics are then *calibrated* by running
# speech-like features through the layers (as training would leave them), and only then is
# the model written.  `nnet_forward` is the float64 numpy forward of exactly this architecture.


def _calib_mfcc(pcm: np.ndarray, spec: "SynthSpec") -> np.ndarray:
    """Plain numpy MFCC (np.fft; NOT bit-faithful to Kaldi) -- calibration data only."""
    x = np.asarray(pcm, dtype=np.float64)
    T = 1 + (len(x) - 400) // 160
    fr = x[np.arange(T)[:, None] * 160 + np.arange(400)[None, :]]
    fr = fr - fr.mean(axis=1, keepdims=True)
    fr = np.concatenate([fr[:, :1] * 0.03, fr[:, 1:] - 0.97 * fr[:, :-1]], axis=1)
    fr = fr * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 399)) ** 0.85
    pw = np.abs(np.fft.rfft(fr, 512, axis=1)) ** 2
    nb = spec.num_mel_bins
    mel = lambda f: 1127.0 * np.log(1.0 + f / 700.0)
    hi = spec.high_freq if spec.high_freq > 0 else 8000.0 + spec.high_freq
    edges = np.linspace(mel(spec.low_freq), mel(hi), nb + 2)
    m = mel(np.arange(257) * (16000.0 / 512))
    fb = np.zeros((nb, 257))
    for b in range(nb):
        l, c, r = edges[b], edges[b + 1], edges[b + 2]
        up = (m - l) / (c - l)
        dn = (r - m) / (r - c)
        fb[b] = np.where((m > l) & (m < r), np.where(m <= c, up, dn), 0.0)
    lm = np.log(np.maximum(pw @ fb.T, np.finfo(np.float32).eps))
    dct = np.sqrt(2.0 / nb) * np.cos(np.pi / nb * (np.arange(nb)[None, :] + 0.5) * np.arange(spec.num_ceps)[:, None])
    dct[0] = np.sqrt(1.0 / nb)
    return (lm @ dct.T) * (1.0 + 11.0 * np.sin(np.pi * np.arange(spec.num_ceps) / 22.0))


def _tdnn_apply(x: np.ndarray, t0: int, W: np.ndarray, b, offsets) -> Tuple[np.ndarray, int]:
    """x holds times t0..t0+len-1 (dense); returns the valid output range and its first time."""
    lo, hi = min(offsets), max(offsets)
    n = x.shape[0] - (hi - lo)
    k = x.shape[1]
    y = np.zeros((n, W.shape[0]))
    for i, o in enumerate(offsets):
        y += x[o - lo:o - lo + n] @ W[:, i * k:(i + 1) * k].T
    if b is not None and len(b):
        y += b
    return y, t0 - lo


def _bn(x, bn):
    mean, var, eps = bn
    scale = 1.0 / np.sqrt(np.maximum(var, 0.0) + eps)
    return x * scale + (-mean * scale)


def nnet_forward(P: dict, feats: np.ndarray, ivector: np.ndarray, calibrate: bool = False, branch: str = "chain"):
    """float64 forward of the synthetic TDNN-F (pseudo log-likelihoods before priors), for every
    input frame t in [0, T) with the reference's edge handling: input frames outside [0, T) are
    copies of the first / last frame (nnet3/decodable-online-looped.cc:150-161)."""
    T = feats.shape[0]
    L, R = P["left_context"], P["right_context"]
    idx = np.clip(np.arange(-L, T + R), 0, T - 1)
    x = feats[idx].astype(np.float64)
    t0 = -L

    def calib(name, v):
        if calibrate:
            P[name] = (v.mean(axis=0), v.var(axis=0) + 1e-3, 1e-3)

    n = x.shape[0] - 2
    sp = np.concatenate([x[0:n], x[1:n + 1], x[2:n + 2], np.tile(ivector.astype(np.float64), (n, 1))], axis=1)
    t0 += 1
    h = sp @ P["lda.W"].T + P["lda.b"]
    h = np.maximum(h @ P["tdnn1.W"].T + P["tdnn1.b"], 0.0)
    calib("tdnn1.bn", h)
    h = _bn(h, P["tdnn1.bn"])
    for li, (off1, off2) in enumerate(P["tdnnf_offsets"]):
        nme = "tdnnf%d" % (li + 2)
        y, ty = _tdnn_apply(h, t0, P[nme + ".lin.W"], None, off1)
        y, ty = _tdnn_apply(y, ty, P[nme + ".aff.W"], P[nme + ".aff.b"], off2)
        y = np.maximum(y, 0.0)
        calib(nme + ".bn", y)
        y = _bn(y, P[nme + ".bn"])
        h = 0.66 * h[ty - t0:ty - t0 + y.shape[0]] + y
        t0 = ty
    s = h @ P["prefinal-l.W"].T
    p = "prefinal-" + branch
    a = np.maximum(s @ P[p + ".aff.W"].T + P[p + ".aff.b"], 0.0)
    calib(p + ".bn1", a)
    a = _bn(a, P[p + ".bn1"])
    a = a @ P[p + ".lin.W"].T
    calib(p + ".bn2", a)
    a = _bn(a, P[p + ".bn2"])
    o = "output" if branch == "chain" else "output-xent"
    out = a @ P[o + ".W"].T + P[o + ".b"]
    assert t0 == 0 and out.shape[0] == T, (t0, out.shape, T)
    return out


def make_nnet_params(spec: "SynthSpec", rng: np.random.Generator, num_pdfs: int) -> dict:
    H, B, D, IV, S = spec.hidden_dim, spec.bottleneck_dim, spec.num_ceps, spec.ivector_dim, spec.prefinal_small
    P: dict = {}
    lda_in = 3 * D + IV
    m1, s1 = mfcc_stats(D)
    in_mean = np.concatenate([m1, m1, m1, np.zeros(IV)])
    in_std = np.concatenate([s1, s1, s1, np.ones(IV)])
    W = rng.standard_normal((lda_in, lda_in)) / math.sqrt(lda_in) / in_std[None, :]
    P["lda.W"] = W.astype(np.float32)
    P["lda.b"] = (rng.standard_normal(lda_in) * 0.1 - W @ in_mean).astype(np.float32)

    def mat(o, i, scale=1.0):
        return (rng.standard_normal((o, i)) * (scale / math.sqrt(i))).astype(np.float32)

    def bn(d):
        return (np.zeros(d), np.ones(d), 1e-3)

    P["tdnn1.W"], P["tdnn1.b"], P["tdnn1.bn"] = mat(H, lda_in), (rng.standard_normal(H) * 0.1).astype(np.float32), bn(H)
    offs = []
    for li, stride in enumerate(spec.tdnnf_strides):
        nme = "tdnnf%d" % (li + 2)
        off1 = [-stride, 0] if stride else [0]
        off2 = [0, stride] if stride else [0]
        offs.append((off1, off2))
        P[nme + ".lin.W"] = mat(B, H * len(off1))
        P[nme + ".aff.W"] = mat(H, B * len(off2), 2.0)
        P[nme + ".aff.b"] = (rng.standard_normal(H) * 0.1).astype(np.float32)
        P[nme + ".bn"] = bn(H)
    P["tdnnf_offsets"] = offs
    P["prefinal-l.W"] = mat(S, H)
    for branch in ("chain", "xent"):
        p = "prefinal-" + branch
        P[p + ".aff.W"], P[p + ".aff.b"], P[p + ".bn1"] = mat(H, S, 2.0), (rng.standard_normal(H) * 0.1).astype(np.float32), bn(H)
        P[p + ".lin.W"], P[p + ".bn2"] = mat(S, H), bn(S)
        o = "output" if branch == "chain" else "output-xent"
        P[o + ".W"] = mat(num_pdfs, S, spec.output_scale)
        P[o + ".b"] = (rng.standard_normal(num_pdfs) * 0.1 * spec.output_scale).astype(np.float32)
    P["left_context"] = 1 + sum(s for s in spec.tdnnf_strides)
    P["right_context"] = 1 + sum(s for s in spec.tdnnf_strides)
    # calibrate the BatchNorm statistics on speech-like features, one branch after the other
    feats = np.concatenate([_calib_mfcc(synth_speech(2.0, 9000 + i), spec) for i in range(3)])
    iv = rng.standard_normal(IV) * 0.7
    for branch in ("chain", "xent"):
        nnet_forward(P, feats, iv, calibrate=True, branch=branch)
    for k in list(P):
        if k.endswith(".bn") or k.endswith(".bn1") or k.endswith(".bn2"):
            mean, var, eps = P[k]
            P[k] = (mean.astype(np.float32).astype(np.float64), var.astype(np.float32).astype(np.float64), eps)
    return P


def _updatable_prefix(w: KaldiWriter, typ: str):
    w.tok("<%s>" % typ)
    w.tok("<MaxChange>"); w.f32(0.75)
    w.tok("<L2Regularize>"); w.f32(0.01)
    w.tok("<LearningRate>"); w.f32(0.001)


def _comp_tdnn(w, W, b, offsets):
    _updatable_prefix(w, "TdnnComponent")
    w.tok("<TimeOffsets>"); w.intvec(offsets)
    w.tok("<LinearParams>"); w.mat(W)
    w.tok("<BiasParams>"); w.vec(b if b is not None else np.zeros(0, np.float32))
    w.tok("<OrthonormalConstraint>"); w.f32(-1.0 if b is None else 0.0)
    w.tok("<UseNaturalGradient>"); w.boolean(True)
    w.tok("<NumSamplesHistory>"); w.f32(2000.0)
    w.tok("<AlphaInOut>"); w.f32(4.0); w.f32(4.0)
    w.tok("<RankInOut>"); w.i32(20); w.i32(20)
    w.tok("</TdnnComponent>"); w.nl()


def _comp_ngaffine(w, W, b):
    _updatable_prefix(w, "NaturalGradientAffineComponent")
    w.tok("<LinearParams>"); w.mat(W)
    w.tok("<BiasParams>"); w.vec(b)
    w.tok("<RankIn>"); w.i32(20)
    w.tok("<RankOut>"); w.i32(80)
    w.tok("<UpdatePeriod>"); w.i32(4)
    w.tok("<NumSamplesHistory>"); w.f32(2000.0)
    w.tok("<Alpha>"); w.f32(4.0)
    w.tok("</NaturalGradientAffineComponent>"); w.nl()


def _comp_linear(w, W):
    _updatable_prefix(w, "LinearComponent")
    w.tok("<Params>"); w.mat(W)
    w.tok("<OrthonormalConstraint>"); w.f32(-1.0)
    w.tok("<UseNaturalGradient>"); w.boolean(True)
    w.tok("<RankInOut>"); w.i32(20); w.i32(80)
    w.tok("<Alpha>"); w.f32(4.0)
    w.tok("<NumSamplesHistory>"); w.f32(2000.0)
    w.tok("<UpdatePeriod>"); w.i32(4)
    w.tok("</LinearComponent>"); w.nl()


def _comp_fixed_affine(w, W, b):
    w.tok("<FixedAffineComponent>")
    w.tok("<LinearParams>"); w.mat(W)
    w.tok("<BiasParams>"); w.vec(b)
    w.tok("</FixedAffineComponent>"); w.nl()


def _comp_nonlin(w, typ, dim):
    w.tok("<%s>" % typ)
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<ValueAvg>"); w.vec(np.zeros(0, np.float32))
    w.tok("<DerivAvg>"); w.vec(np.zeros(0, np.float32))
    w.tok("<Count>"); w.f64(0.0)
    w.tok("<NumDimsSelfRepaired>"); w.f64(0.0)
    w.tok("<NumDimsProcessed>"); w.f64(0.0)
    w.tok("</%s>" % typ); w.nl()


def _comp_batchnorm(w, bn):
    mean, var, eps = bn
    dim = len(mean)
    w.tok("<BatchNormComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<BlockDim>"); w.i32(dim)
    w.tok("<Epsilon>"); w.f32(eps)
    w.tok("<TargetRms>"); w.f32(1.0)
    w.tok("<TestMode>"); w.boolean(False)
    w.tok("<Count>"); w.f64(10000.0)
    w.tok("<StatsMean>"); w.vec(mean.astype(np.float32))
    w.tok("<StatsVar>"); w.vec(var.astype(np.float32))
    w.tok("</BatchNormComponent>"); w.nl()


def _comp_dropout(w, dim):
    w.tok("<GeneralDropoutComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<BlockDim>"); w.i32(dim)
    w.tok("<TimePeriod>"); w.i32(0)
    w.tok("<DropoutProportion>"); w.f32(0.0)
    w.tok("<Continuous>")
    w.tok("</GeneralDropoutComponent>"); w.nl()


def _comp_noop(w, dim):
    w.tok("<NoOpComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<BackpropScale>"); w.f32(1.0)
    w.tok("</NoOpComponent>"); w.nl()


def _write_nnet3(w: KaldiWriter, spec: SynthSpec, rng: np.random.Generator, num_pdfs: int) -> dict:
    H, B, D, IV, S = spec.hidden_dim, spec.bottleneck_dim, spec.num_ceps, spec.ivector_dim, spec.prefinal_small
    P = make_nnet_params(spec, rng, num_pdfs)
    lines = ["input-node name=ivector dim=%d" % IV, "input-node name=input dim=%d" % D]
    comps = []  # (name, writer-callable)

    def node(name, inp):
        lines.append("component-node name=%s component=%s input=%s" % (name, name, inp))

    comps.append(("lda", lambda: _comp_fixed_affine(w, P["lda.W"], P["lda.b"])))
    node("lda", "Append(Offset(input, -1), input, Offset(input, 1), ReplaceIndex(ivector, t, 0))")
    comps.append(("tdnn1.affine", lambda: _comp_ngaffine(w, P["tdnn1.W"], P["tdnn1.b"])))
    node("tdnn1.affine", "lda")
    comps.append(("tdnn1.relu", lambda: _comp_nonlin(w, "RectifiedLinearComponent", H)))
    node("tdnn1.relu", "tdnn1.affine")
    comps.append(("tdnn1.batchnorm", lambda: _comp_batchnorm(w, P["tdnn1.bn"])))
    node("tdnn1.batchnorm", "tdnn1.relu")
    comps.append(("tdnn1.dropout", lambda: _comp_dropout(w, H)))
    node("tdnn1.dropout", "tdnn1.batchnorm")
    prev = "tdnn1.dropout"
    for li, (off1, off2) in enumerate(P["tdnnf_offsets"]):
        n = "tdnnf%d" % (li + 2)
        comps.append((n + ".linear", lambda n=n, o=off1: _comp_tdnn(w, P[n + ".lin.W"], None, o)))
        node(n + ".linear", prev)
        comps.append((n + ".affine", lambda n=n, o=off2: _comp_tdnn(w, P[n + ".aff.W"], P[n + ".aff.b"], o)))
        node(n + ".affine", n + ".linear")
        comps.append((n + ".relu", lambda: _comp_nonlin(w, "RectifiedLinearComponent", H)))
        node(n + ".relu", n + ".affine")
        comps.append((n + ".batchnorm", lambda n=n: _comp_batchnorm(w, P[n + ".bn"])))
        node(n + ".batchnorm", n + ".relu")
        comps.append((n + ".dropout", lambda: _comp_dropout(w, H)))
        node(n + ".dropout", n + ".batchnorm")
        comps.append((n + ".noop", lambda: _comp_noop(w, H)))
        node(n + ".noop", "Sum(Scale(0.66, %s), %s.dropout)" % (prev, n))
        prev = n + ".noop"
    comps.append(("prefinal-l", lambda: _comp_linear(w, P["prefinal-l.W"])))
    node("prefinal-l", prev)
    for branch in ("chain", "xent"):
        p = "prefinal-" + branch
        comps.append((p + ".affine", lambda p=p: _comp_ngaffine(w, P[p + ".aff.W"], P[p + ".aff.b"])))
        node(p + ".affine", "prefinal-l")
        comps.append((p + ".relu", lambda: _comp_nonlin(w, "RectifiedLinearComponent", H)))
        node(p + ".relu", p + ".affine")
        comps.append((p + ".batchnorm1", lambda p=p: _comp_batchnorm(w, P[p + ".bn1"])))
        node(p + ".batchnorm1", p + ".relu")
        comps.append((p + ".linear", lambda p=p: _comp_linear(w, P[p + ".lin.W"])))
        node(p + ".linear", p + ".batchnorm1")
        comps.append((p + ".batchnorm2", lambda p=p: _comp_batchnorm(w, P[p + ".bn2"])))
        node(p + ".batchnorm2", p + ".linear")
        out = "output" if branch == "chain" else "output-xent"
        comps.append((out + ".affine", lambda out=out: _comp_ngaffine(w, P[out + ".W"], P[out + ".b"])))
        node(out + ".affine", p + ".batchnorm2")
        if branch == "xent" or spec.log_softmax:
            comps.append((out + ".log-softmax", lambda: _comp_nonlin(w, "LogSoftmaxComponent", num_pdfs)))
            node(out + ".log-softmax", out + ".affine")
            lines.append("output-node name=%s input=%s.log-softmax objective=linear" % (out, out))
        else:
            lines.append("output-node name=%s input=%s.affine objective=linear" % (out, out))
    w.tok("<Nnet3>")
    w.raw(b"\n")
    for ln in lines:
        w.raw(ln.encode() + b"\n")
    w.raw(b"\n")
    w.tok("<NumComponents>"); w.i32(len(comps)); w.nl()
    for name, fn in comps:
        w.tok("<ComponentName>"); w.tok(name)
        fn()
    w.tok("</Nnet3>"); w.nl()
    w.tok("<LeftContext>"); w.i32(0)
    w.tok("<RightContext>"); w.i32(0)
    w.tok("<Priors>")
    if spec.priors:
        p = rng.dirichlet(np.full(num_pdfs, 5.0)).astype(np.float32)
        w.vec(p)
        P["priors"] = p
    else:
        w.vec(np.zeros(0, np.float32))
    return P


# --------------------------------------------------------------------------------------
# graph


class GraphBuilder:
    """Arc-list builder for an HCLG-shaped WFST (tropical weights, ilabel = transition-id)."""

    def __init__(self):
        self.arcs: List[List[Tuple[int, int, float, int]]] = []
        self.final: List[float] = []

    def add_state(self) -> int:
        self.arcs.append([])
        self.final.append(float("inf"))
        return len(self.arcs) - 1

    def add_arc(self, s, ilabel, olabel, weight, nextstate):
        self.arcs[s].append((int(ilabel), int(olabel), float(weight), int(nextstate)))

    def to_arrays(self):
        ns = len(self.arcs)
        narcs = np.fromiter((len(a) for a in self.arcs), dtype=np.int64, count=ns)
        pos = np.zeros(ns + 1, dtype=np.int64)
        np.cumsum(narcs, out=pos[1:])
        flat = [x for a in self.arcs for x in a]
        arc = np.zeros(len(flat), dtype=[("ilabel", "<i4"), ("olabel", "<i4"), ("weight", "<f4"), ("nextstate", "<i4")])
        if flat:
            il, ol, wt, nx = zip(*flat)
            arc["ilabel"], arc["olabel"], arc["weight"], arc["nextstate"] = il, ol, wt, nx
        return pos, arc, np.asarray(self.final, dtype=np.float32)


def write_const_fst(path: str, start: int, pos: np.ndarray, arc: np.ndarray, final: np.ndarray, aligned: bool = False):
    """ConstFst<StdArc> binary (kaldi/openfst/src/lib/fst.cc:58-82, include/fst/const-fst.h:192-232)."""
    ns, na = final.shape[0], arc.shape[0]

    def s(x: str) -> bytes:
        return struct.pack("<i", len(x)) + x.encode()

    version = 1 if aligned else 2
    hdr = struct.pack("<i", 2125659606) + s("const") + s("standard") + struct.pack("<i", version) + \
        struct.pack("<i", 0) + struct.pack("<Q", 0x1) + struct.pack("<q", start) + struct.pack("<q", ns) + struct.pack("<q", na)
    st = np.zeros(ns, dtype=[("final", "<f4"), ("pos", "<u4"), ("narcs", "<u4"), ("nieps", "<u4"), ("noeps", "<u4")])
    st["final"] = final
    st["pos"] = pos[:-1]
    st["narcs"] = np.diff(pos)
    # per-state epsilon counts; the reference trusts them (ProcessNonemitting only queues states with
    # NumInputEpsilons() != 0, lattice-faster-decoder.cc:846-850), so they must be exact
    owner = np.repeat(np.arange(ns), np.diff(pos).astype(np.int64))
    st["nieps"] = np.bincount(owner[arc["ilabel"] == 0], minlength=ns)
    st["noeps"] = np.bincount(owner[arc["olabel"] == 0], minlength=ns)
    with open(path, "wb") as f:
        f.write(hdr)
        if aligned:
            f.write(b"\0" * ((-f.tell()) % 16))
        f.write(st.tobytes())
        if aligned:
            f.write(b"\0" * ((-f.tell()) % 16))
        f.write(arc.tobytes())


def _lexicon(words: Sequence[str], spec: SynthSpec, rng: np.random.Generator):
    lex = {}
    for wd in words:
        n = int(rng.integers(spec.min_phones, spec.max_phones + 1))
        lex[wd] = [(int(rng.integers(2, spec.num_phones + 2)), int(rng.integers(0, spec.variants_per_phone)))
                   for _ in range(n)]
    return lex


def _hmm_seq(hmm: HmmInfo, phones: Sequence[Tuple[int, int]]):
    out = []
    for pv in phones:
        out.extend(hmm.states[pv])
    return out


def _enter(g: GraphBuilder, src: int, st, olabel: int, cost: float) -> int:
    """Forward arc of hmm state `st` from `src` into a fresh node that carries its self-loop."""
    self_tid, fwd_tid, self_cost, fwd_cost = st
    n = g.add_state()
    g.add_arc(src, fwd_tid, olabel, cost + fwd_cost, n)
    g.add_arc(n, self_tid, 0, self_cost, n)
    return n


def _build_grammar(spec: SynthSpec, hmm: HmmInfo, rng: np.random.Generator):
    sentences = spec.sentences or EN_US_SENTENCES
    words = sorted({wd for s in sentences for wd in s.split()})
    wid = {wd: i + 1 for i, wd in enumerate(words)}
    lex = _lexicon(words, spec, rng)
    sil = hmm.states[(1, 0)]
    g = GraphBuilder()
    start = g.add_state()
    # prefix tree over word sequences
    tree: Dict[Tuple[str, ...], Dict[str, bool]] = {}
    ends = set()
    for s in sentences:
        ws = tuple(s.split())
        for i in range(len(ws)):
            tree.setdefault(ws[:i], {})[ws[i]] = True
        ends.add(ws)

    def junctions(node: int) -> List[Tuple[int, float]]:
        """`node` plus an optional-silence detour, both usable as sources of the next word."""
        s_node = node
        for st in sil:
            s_node = _enter(g, s_node, st, 0, 0.693)
        # one epsilon hop back to an (otherwise empty) junction exercises ProcessNonemitting
        j = g.add_state()
        g.add_arc(s_node, 0, 0, 0.0, j)
        return [(node, 0.693), (j, 0.0)]

    def expand(prefix: Tuple[str, ...], sources: List[Tuple[int, float]]):
        nxt = sorted(tree.get(prefix, {}))
        branch_cost = math.log(len(nxt)) if nxt else 0.0
        for wd in nxt:
            seq = _hmm_seq(hmm, lex[wd])
            first = None
            for src, c in sources:
                if first is None:
                    first = _enter(g, src, seq[0], wid[wd], c + branch_cost)
                else:
                    g.add_arc(src, seq[0][1], wid[wd], c + branch_cost + seq[0][3], first)
            n = first
            for st in seq[1:]:
                n = _enter(g, n, st, 0, 0.0)
            here = prefix + (wd,)
            srcs = junctions(n)
            if here in ends:
                for s_, c_ in srcs:
                    g.final[s_] = min(g.final[s_], c_)
            expand(here, srcs)

    expand((), junctions(start))
    return g, start, words


def _build_arpa(spec: SynthSpec, hmm: HmmInfo, rng: np.random.Generator):
    V = spec.vocab_size
    words = ["w%05d" % i for i in range(V)]
    lex = _lexicon(words, spec, rng)
    seqs = [_hmm_seq(hmm, lex[wd]) for wd in words]
    sil = hmm.states[(1, 0)]
    g = GraphBuilder()
    start = g.add_state()
    uni = g.add_state()               # unigram / back-off hub
    hist = [g.add_state() for _ in range(V)]   # bigram history node of word w (carries w's last self-loop)
    for wi in range(V):
        g.add_arc(hist[wi], seqs[wi][-1][0], 0, seqs[wi][-1][2], hist[wi])

    uni_cost = rng.gamma(2.0, 1.0, V) + math.log(V) * 0.5

    def add_tree(src: int, items: List[Tuple[int, float]]):
        """Lexicon prefix tree from `src`; word label and LM cost sit on the last arc."""
        children: Dict[Tuple[int, Tuple], int] = {}
        for wi, cost in items:
            node = src
            seq = seqs[wi]
            for k, st in enumerate(seq[:-1]):
                key = (node, st[:2])
                if key not in children:
                    children[key] = _enter(g, node, st, 0, 0.0)
                node = children[key]
            last = seq[-1]
            g.add_arc(node, last[1], wi + 1, cost + last[3], hist[wi])

    g.add_arc(start, 0, 0, 0.0, uni)
    add_tree(uni, [(wi, float(uni_cost[wi])) for wi in range(V)])
    # optional silence at the hub
    s_node = uni
    for st in sil:
        s_node = _enter(g, s_node, st, 0, 1.0)
    g.add_arc(s_node, 0, 0, 0.0, uni)
    for wi in range(V):
        succ = rng.choice(V, size=min(spec.bigrams_per_word, V), replace=False)
        add_tree(hist[wi], [(int(s), float(rng.gamma(2.0, 0.7))) for s in succ])
        # back-off: epsilon hop(s) to the unigram hub
        node = hist[wi]
        for _ in range(max(spec.eps_hops - 1, 0)):
            nn = g.add_state()
            g.add_arc(node, 0, 0, 0.1, nn)
            node = nn
        g.add_arc(node, 0, 0, float(rng.gamma(2.0, 0.5)), uni)
        g.final[hist[wi]] = float(rng.gamma(2.0, 1.0))
    return g, start, words


# --------------------------------------------------------------------------------------


def _write_ivector_extractor(ie_dir: str, spec: SynthSpec, rng: np.random.Generator, binary: bool):
    D, G, R, L = spec.num_ceps, spec.num_gauss, spec.ivector_dim, spec.lda_dim
    nsplice = spec.splice_left + 1 + spec.splice_right
    # final.mat
    w = KaldiWriter(binary)
    m1, s1 = mfcc_stats(D)
    lda = (rng.standard_normal((L, D * nsplice)) / np.tile(s1, nsplice)[None, :] / math.sqrt(D * nsplice) * 2.0).astype(np.float32)
    if spec.lda_bias:
        lda = np.concatenate([lda, rng.standard_normal((L, 1)).astype(np.float32) * 0.1], axis=1)
    w.mat(lda)
    w.save(os.path.join(ie_dir, "final.mat"))
    # global_cmvn.stats: [2 x (D+1)] double, sums / sums of squares / count
    cnt = 50000.0
    mean = m1 + rng.standard_normal(D) * 0.5
    var = (s1 * rng.uniform(0.8, 1.2, D)) ** 2
    stats = np.zeros((2, D + 1))
    stats[0, :D] = mean * cnt
    stats[0, D] = cnt
    stats[1, :D] = (var + mean * mean) * cnt
    w = KaldiWriter(binary)
    w.mat(stats, double=True)
    w.save(os.path.join(ie_dir, "global_cmvn.stats"))
    # final.dubm
    means = rng.standard_normal((G, L)).astype(np.float32) * 1.5
    inv_vars = (1.0 / rng.uniform(0.5, 2.0, (G, L))).astype(np.float32)
    weights = rng.dirichlet(np.full(G, 4.0)).astype(np.float32)
    w = KaldiWriter(binary)
    w.tok("<DiagGMM>"); w.nl()
    w.tok("<WEIGHTS>"); w.vec(weights)
    w.tok("<MEANS_INVVARS>"); w.mat(means * inv_vars)
    w.tok("<INV_VARS>"); w.mat(inv_vars)
    w.tok("</DiagGMM>"); w.nl()
    w.save(os.path.join(ie_dir, "final.dubm"))
    # final.ie
    w = KaldiWriter(binary)
    w.tok("<IvectorExtractor>")
    w.tok("<w>"); w.mat(np.zeros((0, 0)), double=True)
    w.tok("<w_vec>"); w.vec(weights.astype(np.float64), double=True)
    w.tok("<M>"); w.i32(G)
    prior_offset = 100.0
    for gi in range(G):
        M = rng.standard_normal((L, R)) * 0.3
        M[:, 0] = means[gi].astype(np.float64) / prior_offset
        w.mat(M, double=True)
    w.tok("<SigmaInv>")
    for gi in range(G):
        A = rng.standard_normal((L, L)) * 0.05
        S = np.diag(inv_vars[gi].astype(np.float64)) + A @ A.T
        w.spmat(S, double=True)
    w.tok("<IvectorOffset>"); w.f64(prior_offset)
    w.tok("</IvectorExtractor>")
    w.save(os.path.join(ie_dir, "final.ie"))


@dataclass
class SynthPaths:
    model_dir: str
    graph_dir: str
    final_mdl: str
    online_conf: str
    hclg: str
    words_txt: str
    words: List[str]
    num_pdfs: int
    num_states: int
    num_arcs: int
    nnet_params: Optional[dict] = None
    tid2pdf: Optional[np.ndarray] = None


def write_graph(graph_dir: str, spec: SynthSpec, hmm: HmmInfo, aligned: bool = False):
    rng = np.random.default_rng(spec.seed + 1000003)
    os.makedirs(graph_dir, exist_ok=True)
    if spec.graph == "grammar":
        g, start, words = _build_grammar(spec, hmm, rng)
    elif spec.graph == "arpa":
        g, start, words = _build_arpa(spec, hmm, rng)
    else:
        raise ValueError(spec.graph)
    pos, arc, final = g.to_arrays()
    write_const_fst(os.path.join(graph_dir, "HCLG.fst"), start, pos, arc, final, aligned=aligned)
    with open(os.path.join(graph_dir, "words.txt"), "w") as f:
        f.write("<eps> 0\n")
        for i, wd in enumerate(words):
            f.write("%s %d\n" % (wd, i + 1))
        f.write("#0 %d\n" % (len(words) + 1))
    return words, final.shape[0], arc.shape[0]


def write_model(root: str, spec: SynthSpec = TINY, aligned_fst: bool = False) -> SynthPaths:
    """Write model + graph for `spec` under `root` and return the paths."""
    rng = np.random.default_rng(spec.seed)
    root = os.path.abspath(root)
    model_dir = os.path.join(root, "model_dir")
    graph_dir = os.path.join(root, "graph")
    mdl_dir = os.path.join(model_dir, "model", "model")
    online = os.path.join(model_dir, "model", "online")
    conf = os.path.join(online, "conf")
    ie_dir = os.path.join(online, "ivector_extractor")
    for d in (mdl_dir, conf, ie_dir, graph_dir):
        os.makedirs(d, exist_ok=True)

    w = KaldiWriter(spec.binary)
    hmm = _write_transition_model(w, spec, rng)
    nnet_params = _write_nnet3(w, spec, rng, hmm.num_pdfs)
    final_mdl = os.path.join(mdl_dir, "final.mdl")
    w.save(final_mdl)
    _write_ivector_extractor(ie_dir, spec, rng, spec.binary)

    with open(os.path.join(conf, "mfcc.conf"), "w") as f:
        f.write("# synthetic hires MFCC config\n--use-energy=false   # comment\n--num-mel-bins=%d\n--num-ceps=%d\n"
                "--low-freq=%g\n--high-freq=%g\n--dither=%g\n" % (spec.num_mel_bins, spec.num_ceps, spec.low_freq,
                                                                    spec.high_freq, spec.dither))
    with open(os.path.join(conf, "splice.conf"), "w") as f:
        f.write("--left-context=%d\n--right-context=%d\n" % (spec.splice_left, spec.splice_right))
    with open(os.path.join(conf, "online_cmvn.conf"), "w") as f:
        f.write("# configuration file for apply-cmvn-online\n")
    with open(os.path.join(conf, "ivector_extractor.conf"), "w") as f:
        f.write("--splice-config=%s/splice.conf\n--cmvn-config=%s/online_cmvn.conf\n--lda-matrix=%s/final.mat\n"
                "--global-cmvn-stats=%s/global_cmvn.stats\n--diag-ubm=%s/final.dubm\n--ivector-extractor=%s/final.ie\n"
                "--num-gselect=%d\n--min-post=%g\n--posterior-scale=%g\n--max-remembered-frames=1000\n--max-count=%g\n"
                "--ivector-period=%d\n" % (conf, conf, ie_dir, ie_dir, ie_dir, ie_dir, spec.num_gselect, spec.min_post,
                                           spec.posterior_scale, spec.max_count, spec.ivector_period))
    online_conf = os.path.join(conf, "online.conf")
    with open(online_conf, "w") as f:
        f.write("--feature-type=mfcc\n--mfcc-config=%s/mfcc.conf\n--ivector-extraction-config=%s/ivector_extractor.conf\n"
                "--endpoint.silence-phones=1\n" % (conf, conf))
        if spec.frame_subsampling_factor != 1:
            f.write("--frame-subsampling-factor=%d\n" % spec.frame_subsampling_factor)
        if spec.nnet_cmvn:
            f.write("--cmvn-config=%s/online_cmvn.conf\n--global-cmvn-stats=%s/global_cmvn.stats\n" % (conf, ie_dir))
    words, ns, na = write_graph(graph_dir, spec, hmm, aligned=aligned_fst)
    return SynthPaths(model_dir, graph_dir, final_mdl, online_conf, os.path.join(graph_dir, "HCLG.fst"),
                      os.path.join(graph_dir, "words.txt"), words, hmm.num_pdfs, ns, na, nnet_params, hmm.tid2pdf)


# --------------------------------------------------------------------------------------
# audio


def write_wav(path: str, pcm: np.ndarray, rate: int = 16000):
    pcm = np.asarray(pcm, dtype="<i2")
    data = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " +
                struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16) + b"data" + struct.pack("<I", len(data)) + data)


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    with open(path, "rb") as f:
        b = f.read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    p = 12
    rate = 16000
    while p + 8 <= len(b):
        cid, sz = b[p:p + 4], struct.unpack("<I", b[p + 4:p + 8])[0]
        if cid == b"fmt ":
            rate = struct.unpack("<I", b[p + 12:p + 16])[0]
        if cid == b"data":
            return np.frombuffer(b[p + 8:p + 8 + sz], dtype="<i2").copy(), rate
        p += 8 + sz + (sz & 1)
    raise ValueError("no data chunk in " + path)


def load_pool(wav_dir: Optional[str] = None) -> List[np.ndarray]:
    """The reference's tests/en_US-zamia utterances (49 WAVs, 16 kHz mono, 1-2 s; committed under
    tests/golden/en_US-zamia): the pool BASELINE config 2 cuts its 3-5 s utterances from (make_utterances(pool=...))."""
    import glob
    if wav_dir is None:
        wav_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "en_US-zamia")
    paths = sorted(glob.glob(os.path.join(wav_dir, "*.wav")))
    if not paths:
        raise FileNotFoundError("no fixture WAVs under " + wav_dir)
    pool = []
    for f in paths:
        pcm, rate = read_wav(f)
        assert rate == 16000, (f, rate)
        pool.append(pcm)
    return pool


def synth_speech(seconds: float, seed: int, rate: int = 16000) -> np.ndarray:
    """Speech-like test audio: voiced harmonics with moving formants + noise bursts + pauses."""
    rng = np.random.default_rng(seed)
    n = int(seconds * rate)
    t = np.arange(n) / rate
    f0 = 110.0 + 40.0 * np.sin(2 * np.pi * 0.7 * t + rng.uniform(0, 6.28)) + rng.uniform(-20, 60)
    phase = 2 * np.pi * np.cumsum(f0) / rate
    sig = np.zeros(n)
    n_seg = max(int(seconds * 6), 1)
    bounds = np.sort(rng.integers(0, n, n_seg - 1))
    seg_id = np.searchsorted(bounds, np.arange(n))
    formants = rng.uniform(300, 3200, (n_seg, 3))
    voiced = rng.uniform(0, 1, n_seg) < 0.7
    silent = rng.uniform(0, 1, n_seg) < 0.15
    for h in range(1, 30):
        fh = f0 * h
        amp = np.zeros(n)
        for k in range(3):
            fc = formants[seg_id, k]
            amp += np.exp(-0.5 * ((fh - fc) / 180.0) ** 2)
        sig += amp * np.sin(h * phase) / h ** 0.5
    noise = rng.standard_normal(n)
    sig = np.where(voiced[seg_id], sig, 0.6 * noise)
    sig = np.where(silent[seg_id], 0.01 * noise, sig)
    env = np.convolve(np.abs(rng.standard_normal(n)), np.ones(800) / 800, mode="same")
    sig = sig * (0.5 + env)
    sig = sig / (np.max(np.abs(sig)) + 1e-9) * rng.uniform(3000, 12000)
    sig += rng.standard_normal(n) * 2.0
    return np.clip(np.round(sig), -32768, 32767).astype(np.int16)


def make_utterances(n: int, seed: int = 1234, min_s: float = 3.0, max_s: float = 5.0,
                    pool: Optional[List[np.ndarray]] = None, noise_seed: int = 5678) -> List[np.ndarray]:
    """BASELINE config-2 style inputs: 3-5 s utterances.

    With `pool` (fixture WAVs) utterances are concatenations of pool items cut at a
    uniformly drawn length, plus sigma=2 LSB Gaussian noise so that lanes differ
    (SURVEY.md section 8d); without it they are synthesised.
    """
    rng = np.random.default_rng(seed)
    nrng = np.random.default_rng(noise_seed)
    out = []
    for i in range(n):
        target = int(rng.uniform(min_s, max_s) * 16000)
        if pool:
            parts = []
            tot = 0
            while tot < target:
                p = pool[int(rng.integers(0, len(pool)))]
                parts.append(p)
                tot += len(p)
            x = np.concatenate(parts)[:target].astype(np.float64)
            x = x + nrng.standard_normal(target) * 2.0
            out.append(np.clip(np.round(x), -32768, 32767).astype(np.int16))
        else:
            out.append(synth_speech(target / 16000.0, seed * 100003 + i))
    return out
