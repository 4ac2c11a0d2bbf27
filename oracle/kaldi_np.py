"""CPU restatement (numpy) of the feature / iVector stages of the reference's hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product (rhasspy_speech_b200/*).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Parity status: PINNED against outputs of the reference itself, run in this container from
oracle/_ref (built by oracle/build_ref.py from /root/reference/kaldi): `compute-mfcc-feats` for
mfcc() and `ivector-extract-online2` for the periodic iVector schedule (tests/test_oracle.py,
fixtures in tests/golden made by tests/golden/make_golden.py).  The acoustic model forward and
the decoder are not restated here: their oracle is the reference binary itself
(oracle/ref_run.py: nnet3-compute, latgen-faster-mapped, online2-wav-nnet3-latgen-faster), which
travels to the GPU box inside oracle/_ref.

Each function cites the reference lines it follows (paths relative to /root/reference/kaldi/src).
float32 arithmetic is done op by op in the reference's order (x86-64 build without FMA), so mfcc()
is bit-exact up to the BLAS dot products (mel filterbank, DCT), whose summation order is
OpenBLAS-kernel specific.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
import os
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "logf", "expf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
for _n in ("cos", "sin", "log", "exp"):
    getattr(_libm, _n).restype = ctypes.c_double
    getattr(_libm, _n).argtypes = [ctypes.c_double]
_libm.pow.restype = ctypes.c_double
_libm.pow.argtypes = [ctypes.c_double, ctypes.c_double]
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]


def cosf(x):
    return f32(_libm.cosf(float(f32(x))))


def sinf(x):
    return f32(_libm.sinf(float(f32(x))))


def logf(x):
    return f32(_libm.logf(float(f32(x))))


# ----------------------------------------------------------------------------------------------
# Kaldi object files (binary), base/io-funcs-inl.h, matrix/kaldi-matrix.cc


class KReader:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.b = f.read()
        assert self.b[:2] == b"\0B", "oracle reader handles binary Kaldi files only: " + path
        self.p = 2

    def token(self) -> str:
        e = self.b.index(b" ", self.p)
        t = self.b[self.p:e].decode()
        self.p = e + 1
        return t

    def expect(self, t: str):
        g = self.token()
        assert g == t, (g, t)

    def i32(self) -> int:
        assert self.b[self.p] == 4
        v = struct.unpack_from("<i", self.b, self.p + 1)[0]
        self.p += 5
        return v

    def f64(self) -> float:
        sz = self.b[self.p]
        v = struct.unpack_from("<d" if sz == 8 else "<f", self.b, self.p + 1)[0]
        self.p += 1 + sz
        return v

    def vec(self) -> np.ndarray:
        t = self.token()
        n = self.i32()
        dt = "<f4" if t == "FV" else "<f8"
        v = np.frombuffer(self.b, dt, n, self.p).copy()
        self.p += v.nbytes
        return v

    def mat(self) -> np.ndarray:
        t = self.token()
        r, c = self.i32(), self.i32()
        dt = "<f4" if t == "FM" else "<f8"
        m = np.frombuffer(self.b, dt, r * c, self.p).reshape(r, c).copy()
        self.p += m.nbytes
        return m

    def spmat(self) -> np.ndarray:
        t = self.token()
        n = self.i32()
        dt = "<f4" if t == "FP" else "<f8"
        k = n * (n + 1) // 2
        packed = np.frombuffer(self.b, dt, k, self.p).copy()
        self.p += packed.nbytes
        m = np.zeros((n, n), dtype=packed.dtype)
        m[np.tril_indices(n)] = packed
        return m + np.tril(m, -1).T


def read_conf(path: str) -> Dict[str, str]:
    """util/parse-options.cc ReadConfigFile."""
    kv = {}
    with open(path) as f:
        for line in f:
            line = line.split("#")[0].strip()
            if not line:
                continue
            assert line.startswith("--"), line
            k, _, v = line[2:].partition("=")
            kv[k.replace("_", "-")] = v.strip() if _ else "true"
    return kv


# ----------------------------------------------------------------------------------------------
# MFCC


@dataclass
class MfccOpts:
    """feat/feature-window.h:53-105, mel-computations.h:56-74, feature-mfcc.h:51-81 defaults."""
    samp_freq: float = 16000.0
    frame_shift_ms: float = 10.0
    frame_length_ms: float = 25.0
    dither: float = 1.0
    preemph: float = 0.97
    remove_dc: bool = True
    window_type: str = "povey"
    num_bins: int = 23
    low_freq: float = 20.0
    high_freq: float = 0.0
    num_ceps: int = 13
    use_energy: bool = True
    raw_energy: bool = True
    energy_floor: float = 0.0
    cepstral_lifter: float = 22.0

    @staticmethod
    def from_conf(path: str) -> "MfccOpts":
        kv = read_conf(path)
        o = MfccOpts()
        names = {"sample-frequency": "samp_freq", "frame-shift": "frame_shift_ms", "frame-length": "frame_length_ms",
                 "dither": "dither", "preemphasis-coefficient": "preemph", "num-mel-bins": "num_bins", "low-freq": "low_freq",
                 "high-freq": "high_freq", "num-ceps": "num_ceps", "energy-floor": "energy_floor",
                 "cepstral-lifter": "cepstral_lifter"}
        for k, v in kv.items():
            if k in names:
                cur = getattr(o, names[k])
                setattr(o, names[k], type(cur)(float(v)) if not isinstance(cur, int) else int(v))
            elif k in ("use-energy", "raw-energy", "remove-dc-offset"):
                setattr(o, {"use-energy": "use_energy", "raw-energy": "raw_energy", "remove-dc-offset": "remove_dc"}[k], v == "true")
            elif k == "window-type":
                o.window_type = v
            else:
                raise ValueError("unsupported mfcc option " + k)
        return o


class SplitRadixRealFft:
    """matrix/srfft.cc: ComputeTables :77-118, ComputeRecursive :207-345, BitReversePermute :180-204,
    SplitRadixRealFft::Compute :356-440.  Vectorised over frames (rows); float32, same op order."""

    def __init__(self, n: int):
        assert n & (n - 1) == 0 and n >= 8
        self.N = n
        self.N_ = n // 2
        self.logn = int(math.log2(self.N_))
        self.tabs = {}
        for i in range(self.logn, 3, -1):
            m = 1 << i
            m4, m8 = m // 4, m // 8
            cn, spcn, smcn, c3n, spc3n, smc3n = [], [], [], [], [], []
            for nn in range(1, m4):
                if nn == m8:
                    continue
                ang = f32(nn * (2 * math.pi) / m)
                c, s = cosf(ang), sinf(ang)
                cn.append(c); spcn.append(f32(-(s + c))); smcn.append(f32(s - c))
                ang = f32(3 * nn * (2 * math.pi) / m)
                c, s = cosf(ang), sinf(ang)
                c3n.append(c); spc3n.append(f32(-(s + c))); smc3n.append(f32(s - c))
            self.tabs[i] = [np.asarray(t, dtype=f32) for t in (cn, spcn, smcn, c3n, spc3n, smc3n)]
        # bit-reverse permutation: run the reference's swap sequence on an index vector
        lg2 = self.logn >> 1
        if self.logn & 1:
            lg2 += 1
        brseed = [0] * (1 << lg2)
        brseed[1] = 1
        for j in range(2, lg2 + 1):
            imax = 1 << (j - 1)
            for i in range(imax):
                brseed[i] <<= 1
                brseed[i + imax] = brseed[i] + 1
        perm = list(range(self.N_))
        lg2 = self.logn >> 1
        n2 = 1 << lg2
        for off in range(1, n2):
            fj = n2 * brseed[off]
            i, j = off, fj
            perm[i], perm[j] = perm[j], perm[i]
            xp = i
            for gno in range(1, brseed[off]):
                xp += n2
                j = fj + brseed[gno]
                perm[xp], perm[j] = perm[j], perm[xp]
        self.perm = np.asarray(perm)
        # real-FFT twiddles: the reference iterates kN *= rootN in float (ComplexMul)
        rre, rim = cosf(f32(-(2 * math.pi) / n)), sinf(f32(-(2 * math.pi) / n))
        kre, kim = f32(1.0), f32(0.0)
        self.kn = []
        for k in range(1, self.N_ // 2 + 1):
            tre = f32(f32(kre * rre) - f32(kim * rim))
            kim = f32(f32(kre * rim) + f32(kim * rre))
            kre = tre
            self.kn.append((kre, kim))

    def _rec(self, xr, xi, o, logn):
        """xr/xi: [F, N_] float32 arrays modified in place; block at offset o of size 2**logn."""
        if logn < 3:
            if logn == 2:
                a, b = xr[:, o].copy(), xr[:, o + 2].copy()
                xr[:, o], xr[:, o + 2] = a + b, a - b
                a, b = xi[:, o].copy(), xi[:, o + 2].copy()
                xi[:, o], xi[:, o + 2] = a + b, a - b
                a, b = xr[:, o + 1].copy(), xr[:, o + 3].copy()
                xr[:, o + 1], xr[:, o + 3] = a + b, a - b
                a, b = xi[:, o + 1].copy(), xi[:, o + 3].copy()
                xi[:, o + 1], xi[:, o + 3] = a + b, a - b
                a, b = xr[:, o].copy(), xr[:, o + 1].copy()
                xr[:, o], xr[:, o + 1] = a + b, a - b
                a, b = xi[:, o].copy(), xi[:, o + 1].copy()
                xi[:, o], xi[:, o + 1] = a + b, a - b
                r1, i1 = xr[:, o + 2].copy(), xi[:, o + 2].copy()
                r2, i2 = xr[:, o + 3].copy(), xi[:, o + 3].copy()
                xi[:, o + 2] = i1 - r2
                xr[:, o + 3] = r1 - i2
                xr[:, o + 2] = r1 + i2
                xi[:, o + 3] = i1 + r2
            elif logn == 1:
                a, b = xr[:, o].copy(), xr[:, o + 1].copy()
                xr[:, o], xr[:, o + 1] = a + b, a - b
                a, b = xi[:, o].copy(), xi[:, o + 1].copy()
                xi[:, o], xi[:, o + 1] = a + b, a - b
            return
        m = 1 << logn
        m2, m4, m8 = m // 2, m // 4, m // 8
        # step 1
        for x in (xr, xi):
            a, b = x[:, o:o + m2].copy(), x[:, o + m2:o + m].copy()
            x[:, o:o + m2] = a + b
            x[:, o + m2:o + m] = a - b
        # step 2
        r1 = xr[:, o + m2:o + m2 + m4].copy(); r2 = xr[:, o + m2 + m4:o + m].copy()
        i1 = xi[:, o + m2:o + m2 + m4].copy(); i2 = xi[:, o + m2 + m4:o + m].copy()
        xi[:, o + m2:o + m2 + m4] = i1 - r2
        xr[:, o + m2 + m4:o + m] = r1 - i2
        xr[:, o + m2:o + m2 + m4] = r1 + i2
        xi[:, o + m2 + m4:o + m] = i1 + r2
        # steps 3 & 4
        sq = f32(math.sqrt(0.5))
        idx = [n for n in range(1, m4) if n != m8]
        if logn >= 4:
            cn, spcn, smcn, c3n, spc3n, smc3n = self.tabs[logn]
            ii = np.asarray(idx)
            r1 = xr[:, o + m2 + ii].copy(); i1 = xi[:, o + m2 + ii].copy()
            r2 = xr[:, o + m2 + m4 + ii].copy(); i2 = xi[:, o + m2 + m4 + ii].copy()
            t2 = cn * (r1 + i1)
            t1 = spcn * r1 + t2
            xr[:, o + m2 + ii] = smcn * i1 + t2
            xi[:, o + m2 + ii] = t1
            t2 = c3n * (r2 + i2)
            t1 = spc3n * r2 + t2
            xr[:, o + m2 + m4 + ii] = smc3n * i2 + t2
            xi[:, o + m2 + m4 + ii] = t1
        n = m8
        r1 = xr[:, o + m2 + n].copy(); i1 = xi[:, o + m2 + n].copy()
        r2 = xr[:, o + m2 + m4 + n].copy(); i2 = xi[:, o + m2 + m4 + n].copy()
        xi[:, o + m2 + n] = sq * (i1 - r1)
        xr[:, o + m2 + n] = sq * (r1 + i1)
        xi[:, o + m2 + m4 + n] = -sq * (r2 + i2)
        xr[:, o + m2 + m4 + n] = sq * (i2 - r2)
        self._rec(xr, xi, o, logn - 1)
        self._rec(xr, xi, o + m2, logn - 2)
        self._rec(xr, xi, o + 3 * (m // 4), logn - 2)

    def compute(self, data: np.ndarray) -> np.ndarray:
        """data: [F, N] float32 -> packed spectrum [re0, re(N/2), re1, im1, ...]."""
        data = np.ascontiguousarray(data, dtype=f32)
        N, N2 = self.N, self.N // 2
        xr = data[:, 0::2].copy()
        xi = data[:, 1::2].copy()
        self._rec(xr, xi, 0, self.logn)
        xr = xr[:, self.perm]
        xi = xi[:, self.perm]
        d = np.empty_like(data)
        d[:, 0::2] = xr
        d[:, 1::2] = xi
        half = f32(0.5)
        for k in range(1, N2 // 2 + 1):
            kre, kim = self.kn[k - 1]
            a_re, a_im = d[:, 2 * k].copy(), d[:, 2 * k + 1].copy()
            b_re, b_im = d[:, N - 2 * k].copy(), d[:, N - 2 * k + 1].copy()
            ck_re = half * (a_re + b_re)
            ck_im = half * (a_im - b_im)
            dk_re = half * (a_im + b_im)
            dk_im = -half * (a_re - b_re)
            # ComplexAddProduct(Dk, kN, &data[2k]): c += b*a with a = Dk, b = kN
            d[:, 2 * k] = ck_re + (kre * dk_re - kim * dk_im)
            d[:, 2 * k + 1] = ck_im + (kre * dk_im + kim * dk_re)
            kd = N2 - k
            if kd != k:
                ndk_im = -dk_im
                nkre = -kre
                d[:, 2 * kd] = ck_re + (nkre * dk_re - kim * ndk_im)
                d[:, 2 * kd + 1] = -ck_im + (nkre * ndk_im + kim * dk_re)
        z = d[:, 0] + d[:, 1]
        n2 = d[:, 0] - d[:, 1]
        d[:, 0], d[:, 1] = z, n2
        return d


class MfccComputer:
    def __init__(self, o: MfccOpts):
        self.o = o
        self.shift = int(o.samp_freq * 0.001 * o.frame_shift_ms)
        self.length = int(o.samp_freq * 0.001 * o.frame_length_ms)
        self.padded = 1
        while self.padded < self.length:
            self.padded <<= 1
        # feat/feature-window.cc:109-135 (double, stored float)
        a = (2 * math.pi) / (self.length - 1)
        assert o.window_type == "povey"
        self.window = np.asarray([_libm.pow(0.5 - 0.5 * _libm.cos(a * float(i)), 0.85) for i in range(self.length)], dtype=f32)
        self.fft = SplitRadixRealFft(self.padded)
        # feat/mel-computations.cc:33-142
        nyq = f32(0.5) * f32(o.samp_freq)
        low = f32(o.low_freq)
        high = f32(o.high_freq) if o.high_freq > 0 else f32(nyq + f32(o.high_freq))
        nfft = self.padded // 2
        bw = f32(f32(o.samp_freq) / f32(self.padded))
        mel = lambda fr: f32(f32(1127.0) * logf(f32(f32(1.0) + f32(f32(fr) / f32(700.0)))))
        mlow, mhigh = mel(low), mel(high)
        delta = f32(f32(mhigh - mlow) / f32(o.num_bins + 1))
        self.bins = []
        for b in range(o.num_bins):
            left = f32(mlow + f32(f32(b) * delta))
            center = f32(mlow + f32(f32(b + 1) * delta))
            right = f32(mlow + f32(f32(b + 2) * delta))
            w = np.zeros(nfft, dtype=f32)
            first = last = -1
            for i in range(nfft):
                m = mel(f32(bw * f32(i)))
                if m > left and m < right:
                    if m <= center:
                        w[i] = f32(f32(m - left) / f32(center - left))
                    else:
                        w[i] = f32(f32(right - m) / f32(right - center))
                    if first < 0:
                        first = i
                    last = i
            assert first >= 0
            self.bins.append((first, w[first:last + 1].copy()))
        # matrix/matrix-functions.cc:592-608 (rows 0..num_ceps-1), feat/mel-computations.cc:253-259
        N = o.num_bins
        dct = np.zeros((o.num_ceps, N), dtype=f32)
        dct[0, :] = f32(math.sqrt(1.0 / float(f32(N))))
        nz = float(f32(math.sqrt(2.0 / float(f32(N)))))
        for k in range(1, o.num_ceps):
            for n in range(N):
                dct[k, n] = f32(nz * _libm.cos(math.pi / N * (n + 0.5) * k))
        self.dct = dct
        Q = float(f32(o.cepstral_lifter))
        self.lifter = np.asarray([1.0 + 0.5 * Q * _libm.sin(math.pi * i / Q) for i in range(o.num_ceps)], dtype=f32) if Q != 0 else None

    def num_frames(self, nsamp: int) -> int:
        return 0 if nsamp < self.length else 1 + (nsamp - self.length) // self.shift

    def windows(self, pcm: np.ndarray) -> np.ndarray:
        """feature-window.cc:137-224 with dither == 0: [T, padded] float32."""
        x = np.asarray(pcm).astype(f32)
        T = self.num_frames(len(x))
        idx = np.arange(T)[:, None] * self.shift + np.arange(self.length)[None, :]
        w = x[idx]
        if self.o.remove_dc:
            # window->Add(-window->Sum() / frame_length): Sum() is cblas_sdot against a stride-0 one
            s = w.sum(axis=1, dtype=f32)
            w = w + (-s / f32(self.length))[:, None].astype(f32)
        self.log_energy = None
        if self.o.use_energy and self.o.raw_energy:
            e = np.maximum((w * w).sum(axis=1, dtype=f32), np.finfo(f32).eps)
            self.log_energy = np.log(e.astype(np.float64)).astype(f32)
        if self.o.preemph != 0.0:
            c = f32(self.o.preemph)
            w2 = w.copy()
            w2[:, 1:] = w[:, 1:] - c * w[:, :-1]
            w2[:, 0] = w[:, 0] - c * w[:, 0]
            w = w2
        w = w * self.window[None, :]
        out = np.zeros((T, self.padded), dtype=f32)
        out[:, :self.length] = w
        return out

    def compute(self, pcm: np.ndarray) -> np.ndarray:
        """feature-mfcc.cc:28-80."""
        win = self.windows(pcm)
        if win.shape[0] == 0:
            return np.zeros((0, self.o.num_ceps), dtype=f32)
        spec = self.fft.compute(win)
        half = self.padded // 2
        re, im = spec[:, 2::2], spec[:, 3::2]
        power = np.empty((spec.shape[0], half + 1), dtype=f32)
        power[:, 1:half] = re * re + im * im          # feature-functions.cc:29-51
        power[:, 0] = spec[:, 0] * spec[:, 0]
        power[:, half] = spec[:, 1] * spec[:, 1]
        mel = np.empty((spec.shape[0], self.o.num_bins), dtype=f32)
        for b, (off, w) in enumerate(self.bins):
            mel[:, b] = power[:, off:off + len(w)] @ w
        mel = np.log(np.maximum(mel, np.finfo(f32).eps).astype(f32))
        feat = (mel @ self.dct.T).astype(f32)
        if self.lifter is not None:
            feat = feat * self.lifter[None, :]
        if self.o.use_energy:
            le = self.log_energy
            if self.o.energy_floor > 0.0:
                le = np.maximum(le, f32(math.log(self.o.energy_floor)))
            feat[:, 0] = le
        return feat


# ----------------------------------------------------------------------------------------------
# online CMVN + splice + LDA + UBM posteriors + iVector


@dataclass
class CmvnOpts:
    cmn_window: int = 600
    speaker_frames: int = 600
    global_frames: int = 200
    normalize_mean: bool = True
    normalize_variance: bool = False


def online_cmvn(feats: np.ndarray, global_stats: np.ndarray, o: CmvnOpts = CmvnOpts()) -> np.ndarray:
    """feat/online-feature.cc:337-452 (no speaker stats: spk == utt in the reference's invocation),
    transform/cmvn.cc:64-115."""
    T, D = feats.shape
    out = np.empty_like(feats)
    run = np.zeros(D, dtype=np.float64)
    run2 = np.zeros(D, dtype=np.float64)
    cnt = 0.0
    f64 = feats.astype(np.float64)
    gcount = global_stats[0, D]
    for t in range(T):
        run += f64[t]
        if o.normalize_variance:
            run2 += f64[t] * f64[t]
        cnt += 1.0
        if t - o.cmn_window >= 0:
            run -= f64[t - o.cmn_window]
            if o.normalize_variance:
                run2 -= f64[t - o.cmn_window] ** 2
            cnt -= 1.0
        s0, s1, c = run.copy(), run2.copy(), cnt
        if c < o.cmn_window:
            cg = min(o.cmn_window - c, float(o.global_frames))
            if cg > 0.0:
                s0 = s0 + (cg / gcount) * global_stats[0, :D]
                if o.normalize_variance:
                    s1 = s1 + (cg / gcount) * global_stats[1, :D]
                c = c + (cg / gcount) * gcount
        if not o.normalize_mean:
            out[t] = feats[t]
            continue
        if not o.normalize_variance:
            alpha = f32(-1.0 / c)
            off = (np.float64(alpha) * s0).astype(f32)   # VectorBase<float>::AddVec(float alpha, Vector<double>)
            out[t] = feats[t] + off
        else:
            mean = s0 / c
            var = np.maximum(s1 / c - mean * mean, 1e-20)
            scale = 1.0 / np.sqrt(var)
            out[t] = (feats[t] * scale.astype(f32)) + (-(mean * scale)).astype(f32)
    return out


def splice(feats: np.ndarray, left: int, right: int) -> np.ndarray:
    """feat/online-feature.cc:504-519 (edge frames are repeated)."""
    T = feats.shape[0]
    idx = np.clip(np.arange(T)[:, None] + np.arange(-left, right + 1)[None, :], 0, T - 1)
    return feats[idx].reshape(T, -1)


@dataclass
class IvectorSetup:
    splice_left: int = 4
    splice_right: int = 4
    cmvn: CmvnOpts = field(default_factory=CmvnOpts)
    online_cmvn_iextractor: bool = False
    ivector_period: int = 10
    num_gselect: int = 5
    min_post: float = 0.025
    posterior_scale: float = 0.1
    max_count: float = 0.0
    num_cg_iters: int = 15
    lda: np.ndarray = None
    global_cmvn: np.ndarray = None
    gconsts: np.ndarray = None
    means_invvars: np.ndarray = None
    inv_vars: np.ndarray = None
    M: List[np.ndarray] = None
    sigma_inv: List[np.ndarray] = None
    prior_offset: float = 0.0
    sigma_inv_m: np.ndarray = None   # [G, D, R]
    U: np.ndarray = None             # [G, R, R]

    @staticmethod
    def from_conf(path: str) -> "IvectorSetup":
        kv = read_conf(path)
        s = IvectorSetup()
        if "splice-config" in kv:
            sp = read_conf(kv["splice-config"])
            s.splice_left = int(sp.get("left-context", 4))
            s.splice_right = int(sp.get("right-context", 4))
        if "cmvn-config" in kv:
            cm = read_conf(kv["cmvn-config"])
            s.cmvn = CmvnOpts(int(cm.get("cmn-window", 600)), int(cm.get("speaker-frames", 600)), int(cm.get("global-frames", 200)),
                              cm.get("norm-means", "true") == "true", cm.get("norm-vars", "false") == "true")
        s.online_cmvn_iextractor = kv.get("online-cmvn-iextractor", "false") == "true"
        s.ivector_period = int(kv.get("ivector-period", 10))
        s.num_gselect = int(kv.get("num-gselect", 5))
        s.min_post = float(kv.get("min-post", 0.025))
        s.posterior_scale = float(kv.get("posterior-scale", 0.1))
        s.max_count = float(kv.get("max-count", 0.0))
        s.num_cg_iters = int(kv.get("num-cg-iters", 15))
        s.lda = KReader(kv["lda-matrix"]).mat().astype(f32)
        s.global_cmvn = KReader(kv["global-cmvn-stats"]).mat().astype(np.float64)
        r = KReader(kv["diag-ubm"])
        r.expect("<DiagGMM>")
        t = r.token()
        if t == "<GCONSTS>":
            r.vec()
            r.expect("<WEIGHTS>")
        weights = r.vec().astype(f32)
        r.expect("<MEANS_INVVARS>")
        s.means_invvars = r.mat().astype(f32)
        r.expect("<INV_VARS>")
        s.inv_vars = r.mat().astype(f32)
        # gmm/diag-gmm.cc:94-124 ComputeGconsts (float accumulator, double increments)
        G, D = s.inv_vars.shape
        gc = np.empty(G, dtype=f32)
        offset = f32(-0.5 * 1.8378770664093454835606594728112 * D)
        for m in range(G):
            g = f32(logf(weights[m]) + offset)
            for d in range(D):
                iv, mi = s.inv_vars[m, d], s.means_invvars[m, d]
                g = f32(float(g) + (0.5 * float(logf(iv)) - 0.5 * float(mi) * float(mi) / float(iv)))
            gc[m] = g
        s.gconsts = gc
        r = KReader(kv["ivector-extractor"])
        r.expect("<IvectorExtractor>")
        r.expect("<w>")
        assert r.mat().shape[0] == 0
        r.expect("<w_vec>")
        r.vec()
        r.expect("<M>")
        n = r.i32()
        s.M = [r.mat().astype(np.float64) for _ in range(n)]
        r.expect("<SigmaInv>")
        s.sigma_inv = [r.spmat().astype(np.float64) for _ in range(n)]
        r.expect("<IvectorOffset>")
        s.prior_offset = r.f64()
        # ivector/ivector-extractor.cc:207-218
        s.sigma_inv_m = np.stack([s.sigma_inv[g] @ s.M[g] for g in range(n)])
        s.U = np.stack([s.M[g].T @ s.sigma_inv_m[g] for g in range(n)])
        return s


def lda_feats(s: IvectorSetup, mfcc_feats: np.ndarray, normalized: bool) -> np.ndarray:
    """OnlineSpliceFrames + OnlineTransform (feat/online-feature.cc:504-554) over raw or CMVN'd MFCCs."""
    x = online_cmvn(mfcc_feats, s.global_cmvn, s.cmvn) if normalized else mfcc_feats
    sp = splice(x, s.splice_left, s.splice_right)
    if s.lda.shape[1] == sp.shape[1] + 1:
        return (sp @ s.lda[:, :-1].T + s.lda[:, -1][None, :]).astype(f32)
    return (sp @ s.lda.T).astype(f32)


def gmm_posteriors(s: IvectorSetup, x: np.ndarray, weight: float = 1.0):
    """gmm/diag-gmm.cc:546-562 + hmm/posterior.cc:440-508 + online-ivector-feature.cc:188-245.
    Returns per frame a list of (gauss, posterior * posterior_scale * weight)."""
    ll = (s.gconsts[None, :] + x @ s.means_invvars.T + f32(-0.5) * ((x * x) @ s.inv_vars.T)).astype(f32)
    min_post = min(s.min_post / abs(weight), 0.99)
    out = []
    log_min_post = f32(_libm.logf(float(f32(min_post))))
    for t in range(ll.shape[0]):
        row = ll[t]
        mx = row.max()
        cut = f32(mx + log_min_post)
        sel = np.nonzero(row > cut)[0]
        if len(sel) == 0:
            sel = np.arange(len(row))
        post = np.exp((row[sel] - mx).astype(np.float64)).astype(f32)
        order = np.argsort(-post, kind="stable")[:s.num_gselect]
        sel, post = sel[order], post[order]
        tot = f32(0.0)
        for p in post:
            tot = f32(tot + p)
        cutoff = f32(f32(min_post) * tot)
        n = len(post)
        while n > 1 and post[n - 1] < cutoff:
            tot = f32(tot - post[n - 1])
            n -= 1
        inv = f32(1.0 / float(tot))
        sc = f32(f32(s.posterior_scale) * f32(weight))
        out.append([(int(sel[i]), f32(f32(post[i] * inv) * sc)) for i in range(n)])
    return out


class IvectorStats:
    """ivector/ivector-extractor.cc:611-668 (AccStats), :732-756 (GetIvector), :786-798 (ctor)."""

    def __init__(self, s: IvectorSetup):
        self.s = s
        R = s.M[0].shape[1]
        self.linear = np.zeros(R)
        self.quad = np.zeros((R, R))
        self.num_frames = 0.0
        self.linear[0] += s.prior_offset
        self.quad += np.eye(R)

    def acc(self, feats: np.ndarray, post):
        s = self.s
        tot = 0.0
        G = len(s.M)
        wf = np.zeros((G, feats.shape[1]))
        gw = np.zeros(G)
        for t, ent in enumerate(post):
            for g, w in ent:
                wf[g] += float(w) * feats[t].astype(np.float64)
                gw[g] += float(w)
        # the per-Gaussian totals are accumulated in float in the reference (GaussInfo::tot_weight)
        gw32 = np.zeros(G, dtype=f32)
        for t, ent in enumerate(post):
            for g, w in ent:
                gw32[g] = f32(gw32[g] + w)
        for g in np.nonzero(gw)[0]:
            self.linear += s.sigma_inv_m[g].T @ wf[g]
            self.quad += float(gw32[g]) * s.U[g]
            tot += float(gw32[g])
        if s.max_count > 0.0:
            old = max(self.num_frames, s.max_count) / s.max_count
            new = max(self.num_frames + tot, s.max_count) / s.max_count
            ch = new - old
            if ch != 0.0:
                self.linear[0] += s.prior_offset * ch
                self.quad += ch * np.eye(len(self.linear))
        self.num_frames += tot

    def get_ivector(self, x: np.ndarray) -> np.ndarray:
        s = self.s
        x = x.copy()
        if self.num_frames > 0.0:
            if x[0] == 0.0:
                x[0] = s.prior_offset
            return linear_cgd(self.quad, self.linear, x, s.num_cg_iters)
        x[:] = 0.0
        x[0] = s.prior_offset
        return x


def linear_cgd(A: np.ndarray, b: np.ndarray, x: np.ndarray, max_iters: int) -> np.ndarray:
    """matrix/optimization.cc:453-560 (double; max_error 0, recompute_residual_factor 0.01)."""
    M = A.shape[0]
    x = x.copy()
    x_orig = x.copy()
    p = b - A @ x
    r = -p
    r_cur = r @ r
    r_init = r_cur
    r_recompute = r_cur
    max_err_sq = np.finfo(np.float64).tiny
    rf = 0.01 * 0.01
    inv_rf = 1.0 / rf
    k = 0
    while k < M + 5 and k != max_iters:
        Ap = A @ p
        alpha = -(p @ r) / (p @ Ap)
        x = x + alpha * p
        r = r + alpha * Ap
        r_next = r @ r
        if r_next < rf * r_recompute or r_next > inv_rf * r_recompute:
            r = A @ x - b
            r_next = r @ r
            r_recompute = r_next
        if r_next <= max_err_sq:
            break
        beta = r_next / r_cur
        p = beta * p - r
        r_cur = r_next
        k += 1
    if r_cur > r_init and r_cur > r_init + 1.0e-10 * (b @ b):
        x = np.linalg.solve(A, b) if np.all(np.isfinite(A)) else x_orig
    return x


def ivector_offline(s: IvectorSetup, mfcc_feats: np.ndarray) -> np.ndarray:
    """The schedule of the WAV path (--online=false: use_most_recent_ivector = greedy = true, all audio
    accepted before decoding; online2-wav-nnet3-latgen-faster.cc:151-155, online-ivector-feature.cc:248-279,
    327-355): one stats update over all frames, one CG from the prior.  Returns the nnet input
    (float32, first element minus the prior offset)."""
    R = s.M[0].shape[1]
    st = IvectorStats(s)
    if mfcc_feats.shape[0] > 0:
        xn = lda_feats(s, mfcc_feats, True)
        post = gmm_posteriors(s, xn)
        st.acc(xn if s.online_cmvn_iextractor else lda_feats(s, mfcc_feats, False), post)
    iv = st.get_ivector(np.zeros(R))
    out = iv.astype(f32)
    out[0] = f32(out[0] - f32(s.prior_offset))
    return out


def ivectors_periodic(s: IvectorSetup, mfcc_feats: np.ndarray) -> np.ndarray:
    """ivector-extract-online2's schedule (use_most_recent_ivector=false): stats are updated and the CG
    (warm-started) is run at every t % period == 0; row i is the iVector of frame i*period."""
    R = s.M[0].shape[1]
    T = mfcc_feats.shape[0]
    xn = lda_feats(s, mfcc_feats, True)
    xr = xn if s.online_cmvn_iextractor else lda_feats(s, mfcc_feats, False)
    post = gmm_posteriors(s, xn)
    st = IvectorStats(s)
    cur = np.zeros(R)
    rows = []
    start = 0
    for t in range(T):
        if t % s.ivector_period == 0:
            st.acc(xr[start:t + 1], post[start:t + 1])
            start = t + 1
            cur = st.get_ivector(cur)
            rows.append(cur.copy())
    return np.asarray(rows)


def online_schedule(nsamp: int, num_frames: int, chunk: int, right_context: int, splice_right: int, sf: int,
                    window: int = 400, shift: int = 160) -> Tuple[List[int], List[int]]:
    """Which frames each nnet chunk's iVector has seen when audio is streamed to the reference's stream binary
    (online2bin/online2-cli-nnet3-decode-faster.cc:139-160: reads of 1024 samples, AdvanceDecoding after each).
    Chunks become ready per nnet3/decodable-online-looped.cc:56-84; a chunk takes the iVector of frame
    min(features_ready, ivector_frames_ready) - 1 (:186-194), the iVector stream lags by the splice right context
    until InputFinished (feat/online-feature.cc:496-502), and a new CG is run only when new frames were added
    (online2/online-ivector-feature.cc:248-279).  Returns (solve_frames, chunk_solve)."""
    solve_frames: List[int] = []
    chunk_solve: List[int] = []
    stats = 0

    def advance(F, iv_ready, ready_chunks):
        nonlocal stats
        while len(chunk_solve) < ready_chunks:
            frames = min(F - 1, iv_ready - 1) + 1 if iv_ready > 0 else 0
            if not solve_frames or frames > stats:
                solve_frames.append(frames)
                stats = frames
            chunk_solve.append(len(solve_frames) - 1)

    got = 0
    while got < nsamp:
        got = min(got + 1024, nsamp)
        F = 0 if got < window else 1 + (got - window) // shift
        if F == 0:
            continue
        advance(F, max(0, F - splice_right), max(0, F - right_context) // chunk)
    if num_frames > 0:
        total_out = (num_frames + sf - 1) // sf
        per = chunk // sf
        advance(num_frames, num_frames, (total_out + per - 1) // per)
    return solve_frames, chunk_solve


def ivectors_online(s: IvectorSetup, mfcc_feats: np.ndarray, solve_frames: Sequence[int]) -> np.ndarray:
    """The iVectors of the successive solves of online_schedule(): statistics of the first `frames` frames,
    CG warm-started from the previous solve (use_most_recent_ivector = true, greedy = false).  Rows are the
    nnet inputs (float32, prior offset removed)."""
    R = s.M[0].shape[1]
    xn = lda_feats(s, mfcc_feats, True) if mfcc_feats.shape[0] else np.zeros((0, 1))
    xr = xn if s.online_cmvn_iextractor or not mfcc_feats.shape[0] else lda_feats(s, mfcc_feats, False)
    post = gmm_posteriors(s, xn) if mfcc_feats.shape[0] else []
    st = IvectorStats(s)
    cur = np.zeros(R)
    start = 0
    rows = []
    for f in solve_frames:
        if f > start:
            st.acc(xr[start:f], post[start:f])
            start = f
            cur = st.get_ivector(cur)
        out = cur.astype(f32)
        out[0] = f32(out[0] - f32(s.prior_offset)) if f > 0 or start > 0 else f32(0.0)
        if start == 0:
            out[:] = 0
        rows.append(out)
    return np.asarray(rows)
