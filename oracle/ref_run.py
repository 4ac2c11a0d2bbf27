"""Drive the REAL reference binaries built by oracle/build_ref.py (oracle/_ref/bin).

TEST INFRASTRUCTURE ONLY -- nothing here is on the product path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

The command lines reproduce the pipeline rhasspy-speech shells out to
(reference rhasspy_speech/transcribe_wav.py:45-75):

    online2-wav-nnet3-latgen-faster --online=false --do-endpointing=false
        --word-symbol-table=words.txt --config=online.conf --max-active=7000 --lattice-beam=8.0
        --acoustic-scale=1.0 --beam=24.0 final.mdl HCLG.fst ark:spk2utt scp:wav.scp ark:-
      | lattice-to-nbest --n=N --acoustic-scale=1.0 ark:- ark:-
      | nbest-to-linear ark:- ark:/dev/null ark,t:-

and the per-stage probes used to pin the stages of the restatement (compute-mfcc-feats,
ivector-extract-online2, nnet3-compute, latgen-faster-mapped).
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
BIN = os.path.join(REF_DIR, "bin")


def available() -> bool:
    return os.path.isfile(os.path.join(BIN, "online2-wav-nnet3-latgen-faster"))


def _env(threads: int = 1):
    env = dict(os.environ)
    env["OPENBLAS_NUM_THREADS"] = str(threads)
    env["OMP_NUM_THREADS"] = str(threads)
    env["LD_LIBRARY_PATH"] = REF_DIR + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


def run(cmd: str, input: Optional[bytes] = None, check: bool = True, threads: int = 1) -> Tuple[bytes, bytes]:
    """Run a shell pipeline with oracle/_ref/bin on PATH."""
    env = _env(threads)
    env["PATH"] = BIN + os.pathsep + env["PATH"]
    p = subprocess.run(["bash", "-o", "pipefail", "-c", cmd], input=input, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env)
    if check and p.returncode != 0:
        raise RuntimeError("reference command failed (%d): %s\n%s" % (p.returncode, cmd, p.stderr.decode(errors="replace")[-4000:]))
    return p.stdout, p.stderr


# ----------------------------------------------------------------------------------------------
# Kaldi archive IO (binary float/double matrices and vectors), kaldi/src/util/kaldi-holder-inl.h,
# kaldi/src/matrix/kaldi-matrix.cc Read/Write


def write_mat_ark(path: str, mats: Dict[str, np.ndarray]):
    with open(path, "wb") as f:
        for k, m in mats.items():
            m = np.ascontiguousarray(m, dtype="<f4")
            f.write(k.encode() + b" \0BFM \x04" + struct.pack("<i", m.shape[0]) + b"\x04" + struct.pack("<i", m.shape[1]))
            f.write(m.tobytes())


def read_ark(data: bytes) -> Dict[str, np.ndarray]:
    """Binary ark of FM/DM/FV/DV objects -> dict (insertion ordered)."""
    out: Dict[str, np.ndarray] = {}
    p = 0
    n = len(data)
    while p < n:
        sp = data.index(b" ", p)
        key = data[p:sp].decode()
        p = sp + 1
        assert data[p:p + 2] == b"\0B", "only binary archives are supported"
        p += 2
        tok = data[p:p + 3]
        p += 3
        dt = "<f4" if tok[0:1] == b"F" else "<f8"
        es = 4 if dt == "<f4" else 8
        if tok[1:2] == b"M":
            rows = struct.unpack("<i", data[p + 1:p + 5])[0]
            cols = struct.unpack("<i", data[p + 6:p + 10])[0]
            p += 10
            out[key] = np.frombuffer(data, dtype=dt, count=rows * cols, offset=p).reshape(rows, cols).copy()
            p += rows * cols * es
        elif tok[1:2] == b"V":
            dim = struct.unpack("<i", data[p + 1:p + 5])[0]
            p += 5
            out[key] = np.frombuffer(data, dtype=dt, count=dim, offset=p).copy()
            p += dim * es
        else:
            raise ValueError("unsupported object token %r" % tok)
    return out


def parse_int_ark_text(text: bytes) -> Dict[str, List[int]]:
    """Text Int32Vector archive 'key i i i \\n' (kaldi/src/util/kaldi-holder-inl.h:244-251)."""
    out: Dict[str, List[int]] = {}
    for line in text.decode().splitlines():
        parts = line.split()
        if parts:
            out[parts[0]] = [int(x) for x in parts[1:]]
    return out


# ----------------------------------------------------------------------------------------------
# whole pipeline


def _write_lists(tmp: str, wavs: Sequence[str], keys: Optional[Sequence[str]] = None):
    keys = list(keys) if keys is not None else ["utt%05d" % i for i in range(len(wavs))]
    with open(os.path.join(tmp, "wav.scp"), "w") as f:
        for k, w in zip(keys, wavs):
            f.write("%s %s\n" % (k, w))
    with open(os.path.join(tmp, "spk2utt"), "w") as f:
        for k in keys:
            f.write("%s %s\n" % (k, k))
    return keys


def transcribe_wavs(final_mdl: str, online_conf: str, hclg: str, words_txt: str, wavs: Sequence[str],
                    nbest: int = 1, beam: float = 24.0, max_active: int = 7000, lattice_beam: float = 8.0,
                    acoustic_scale: float = 1.0, online: bool = False, extra: str = "",
                    threads: int = 1) -> Tuple[Dict[str, List[int]], bytes, bytes]:
    """WAV path of the reference.  Returns ({'utt00000-1': [word ids]}, nbest_stdout, stderr)."""
    with tempfile.TemporaryDirectory() as tmp:
        _write_lists(tmp, wavs)
        cmd = ("online2-wav-nnet3-latgen-faster --online=%s --do-endpointing=false --word-symbol-table=%s "
               "--config=%s --max-active=%d --lattice-beam=%g --acoustic-scale=1.0 --beam=%g %s %s %s "
               "ark:%s/spk2utt scp:%s/wav.scp ark:- 2>%s/err.log | lattice-to-nbest --n=%d --acoustic-scale=%g ark:- ark:- 2>/dev/null | "
               "nbest-to-linear ark:- ark:/dev/null ark,t:- 2>/dev/null"
               % ("true" if online else "false", words_txt, online_conf, max_active, lattice_beam, beam, extra,
                  final_mdl, hclg, tmp, tmp, tmp, nbest, acoustic_scale))
        out, _ = run(cmd, threads=threads)
        with open(os.path.join(tmp, "err.log"), "rb") as f:
            err = f.read()
    return parse_int_ark_text(out), out, err


def transcribe_wavs_costs(final_mdl: str, online_conf: str, hclg: str, words_txt: str, wavs: Sequence[str],
                          beam: float = 24.0, max_active: int = 7000, lattice_beam: float = 8.0,
                          ) -> Tuple[Dict[str, List[int]], Dict[str, Tuple[float, float]]]:
    """As transcribe_wavs (n = 1), plus the (graph, acoustic) cost of each best path
    (the two extra outputs of nbest-to-linear, kaldi/src/latbin/nbest-to-linear.cc:70-92)."""
    with tempfile.TemporaryDirectory() as tmp:
        _write_lists(tmp, wavs)
        cmd = ("online2-wav-nnet3-latgen-faster --online=false --do-endpointing=false --word-symbol-table=%s "
               "--config=%s --max-active=%d --lattice-beam=%g --acoustic-scale=1.0 --beam=%g %s %s "
               "ark:%s/spk2utt scp:%s/wav.scp ark:- 2>/dev/null | lattice-to-nbest --n=1 --acoustic-scale=1.0 ark:- ark:- 2>/dev/null | "
               "nbest-to-linear ark:- ark:/dev/null ark,t:%s/tr.txt ark,t:%s/lm.txt ark,t:%s/ac.txt 2>/dev/null"
               % (words_txt, online_conf, max_active, lattice_beam, beam, final_mdl, hclg, tmp, tmp, tmp, tmp, tmp))
        run(cmd)
        with open(os.path.join(tmp, "tr.txt"), "rb") as f:
            words = parse_int_ark_text(f.read())
        costs: Dict[str, Tuple[float, float]] = {}
        lm = {l.split()[0]: float(l.split()[1]) for l in open(os.path.join(tmp, "lm.txt")) if l.strip()}
        ac = {l.split()[0]: float(l.split()[1]) for l in open(os.path.join(tmp, "ac.txt")) if l.strip()}
        for k in lm:
            costs[k] = (lm[k], ac.get(k, float("nan")))
    return words, costs


def transcribe_stream(final_mdl: str, online_conf: str, hclg: str, words_txt: str, pcm: np.ndarray,
                      nbest: int = 1, beam: float = 24.0, max_active: int = 7000, lattice_beam: float = 8.0,
                      acoustic_scale: float = 1.0) -> Tuple[Dict[str, List[int]], bytes]:
    """Stream path (reference rhasspy_speech/transcribe_stream.py:51-99): raw s16le on stdin."""
    with tempfile.TemporaryDirectory() as tmp:
        lat = os.path.join(tmp, "lat.ark")
        cmd = ("online2-cli-nnet3-decode-faster --config=%s --max-active=%d --lattice-beam=%g --acoustic-scale=1.0 "
               "--beam=%g %s %s %s ark:%s 2>/dev/null" % (online_conf, max_active, lattice_beam, beam, final_mdl, hclg, words_txt, lat))
        run(cmd, input=np.asarray(pcm, dtype="<i2").tobytes())
        out, _ = run("lattice-to-nbest --n=%d --acoustic-scale=%g ark:%s ark:- 2>/dev/null | "
                     "nbest-to-linear ark:- ark:/dev/null ark,t:- 2>/dev/null" % (nbest, acoustic_scale, lat))
    return parse_int_ark_text(out), out


# ----------------------------------------------------------------------------------------------
# per-stage probes


def mfcc(mfcc_conf: str, wavs: Sequence[str]) -> List[np.ndarray]:
    with tempfile.TemporaryDirectory() as tmp:
        keys = _write_lists(tmp, wavs)
        out, _ = run("compute-mfcc-feats --config=%s scp:%s/wav.scp ark:- 2>/dev/null" % (mfcc_conf, tmp))
    d = read_ark(out)
    return [d[k] for k in keys]


def ivectors_periodic(ivector_conf: str, feats: Sequence[np.ndarray], repeat: bool = False) -> List[np.ndarray]:
    """ivector-extract-online2: one iVector per --ivector-period frames (periodic schedule, warm-started CG)."""
    with tempfile.TemporaryDirectory() as tmp:
        keys = ["utt%05d" % i for i in range(len(feats))]
        write_mat_ark(os.path.join(tmp, "feats.ark"), dict(zip(keys, feats)))
        with open(os.path.join(tmp, "spk2utt"), "w") as f:
            for k in keys:
                f.write("%s %s\n" % (k, k))
        out, _ = run("ivector-extract-online2 --config=%s --repeat=%s ark:%s/spk2utt ark:%s/feats.ark ark:- 2>/dev/null"
                     % (ivector_conf, "true" if repeat else "false", tmp, tmp))
    d = read_ark(out)
    return [d[k] for k in keys]


def nnet_loglikes(final_mdl: str, feats: Sequence[np.ndarray], ivectors: Optional[Sequence[np.ndarray]],
                  frame_subsampling_factor: int = 1, use_priors: bool = True, ivector_period: int = 10) -> List[np.ndarray]:
    """nnet3-compute on the AmNnetSimple (DecodableNnetSimple, non-looped); `ivectors[i]` is one
    vector per utterance and is supplied as a constant online-ivector matrix."""
    with tempfile.TemporaryDirectory() as tmp:
        keys = ["utt%05d" % i for i in range(len(feats))]
        write_mat_ark(os.path.join(tmp, "feats.ark"), dict(zip(keys, feats)))
        iv_opt = ""
        if ivectors is not None:
            mats = {}
            for k, f, iv in zip(keys, feats, ivectors):
                n = (f.shape[0] + ivector_period - 1) // ivector_period
                mats[k] = np.tile(np.asarray(iv, dtype=np.float32)[None, :], (n, 1))
            write_mat_ark(os.path.join(tmp, "iv.ark"), mats)
            iv_opt = "--online-ivectors=ark:%s/iv.ark --online-ivector-period=%d" % (tmp, ivector_period)
        out, _ = run("nnet3-compute --use-priors=%s --frame-subsampling-factor=%d %s %s ark:%s/feats.ark ark:- 2>/dev/null"
                     % ("true" if use_priors else "false", frame_subsampling_factor, iv_opt, final_mdl, tmp))
    d = read_ark(out)
    return [d[k] for k in keys]


def decode_loglikes(final_mdl: str, hclg: str, loglikes: Sequence[np.ndarray], beam: float = 24.0,
                    max_active: int = 7000, min_active: int = 200, lattice_beam: float = 8.0,
                    nbest: int = 1) -> Dict[str, List[int]]:
    """latgen-faster-mapped (LatticeFasterDecoder over a log-likelihood matrix) + the n-best tail."""
    with tempfile.TemporaryDirectory() as tmp:
        keys = ["utt%05d" % i for i in range(len(loglikes))]
        write_mat_ark(os.path.join(tmp, "ll.ark"), dict(zip(keys, loglikes)))
        out, _ = run("latgen-faster-mapped --acoustic-scale=1.0 --beam=%g --max-active=%d --min-active=%d --lattice-beam=%g "
                     "--allow-partial=true %s %s ark:%s/ll.ark ark:- 2>/dev/null | lattice-to-nbest --n=%d --acoustic-scale=1.0 ark:- ark:- 2>/dev/null | "
                     "nbest-to-linear ark:- ark:/dev/null ark,t:- 2>/dev/null"
                     % (beam, max_active, min_active, lattice_beam, final_mdl, hclg, tmp, nbest))
    return parse_int_ark_text(out)


def parse_lattice_text(text: str) -> Dict[str, dict]:
    """Text Lattice archive (kaldi/src/lat/kaldi-lattice.cc WriteLattice: key line, OpenFst text with
    'graph,acoustic' weights, blank line) -> {key: {src, dst, ilabel, olabel, graph, acoustic, n_states}};
    final weights are rows with dst == -1."""
    out: Dict[str, dict] = {}
    key = None
    rows: List[tuple] = []

    def weight(tok):
        g, a = tok.split(",")
        return float(g), float(a)

    def flush():
        if key is None:
            return
        a = np.array(rows, dtype=np.float64).reshape(-1, 6)
        n_states = int(max(a[:, 0].max(), a[:, 1].max())) + 1 if len(a) else 0
        out[key] = dict(src=a[:, 0].astype(np.int32), dst=a[:, 1].astype(np.int32), ilabel=a[:, 2].astype(np.int32),
                        olabel=a[:, 3].astype(np.int32), graph=a[:, 4].astype(np.float32),
                        acoustic=a[:, 5].astype(np.float32), n_states=n_states)
    for line in text.splitlines():
        parts = line.split()
        if not parts:
            flush()
            key, rows = None, []
            continue
        if key is None:
            key = parts[0]
            continue
        if len(parts) >= 4:      # arc: src dst ilabel olabel [weight]
            g, a = weight(parts[4]) if len(parts) > 4 else (0.0, 0.0)
            rows.append((int(parts[0]), int(parts[1]), int(parts[2]), int(parts[3]), g, a))
        else:                    # final: state [weight]
            g, a = weight(parts[1]) if len(parts) > 1 else (0.0, 0.0)
            rows.append((int(parts[0]), -1, 0, 0, g, a))
    flush()
    return out


def decode_loglikes_lattice(final_mdl: str, hclg: str, loglikes: Sequence[np.ndarray], nbest: int, beam: float = 24.0,
                            max_active: int = 7000, min_active: int = 200, lattice_beam: float = 8.0,
                            acoustic_scale: float = 1.0):
    """latgen-faster-mapped twice on the same log-likelihoods: (a) --determinize-lattice=false, the raw state-level
    lattice (GetRawLattice after FinalizeDecoding) as text; (b) the determinised lattice through
    lattice-to-nbest --n | nbest-to-linear with both cost archives.
    Returns ({key: raw lattice}, {key-k: (words, graph cost, acoustic cost)})."""
    with tempfile.TemporaryDirectory() as tmp:
        keys = ["utt%05d" % i for i in range(len(loglikes))]
        write_mat_ark(os.path.join(tmp, "ll.ark"), dict(zip(keys, loglikes)))
        common = ("latgen-faster-mapped --acoustic-scale=1.0 --beam=%g --max-active=%d --min-active=%d --lattice-beam=%g "
                  "--allow-partial=true" % (beam, max_active, min_active, lattice_beam))
        raw, _ = run("%s --determinize-lattice=false %s %s ark:%s/ll.ark ark,t:- 2>/dev/null" % (common, final_mdl, hclg, tmp))
        run("%s %s %s ark:%s/ll.ark ark:- 2>/dev/null | lattice-to-nbest --n=%d --acoustic-scale=%g ark:- ark:- 2>/dev/null | "
            "nbest-to-linear ark:- ark:/dev/null ark,t:%s/tr.txt ark,t:%s/lm.txt ark,t:%s/ac.txt 2>/dev/null"
            % (common, final_mdl, hclg, tmp, nbest, acoustic_scale, tmp, tmp, tmp))
        with open(os.path.join(tmp, "tr.txt"), "rb") as f:
            words = parse_int_ark_text(f.read())
        lm = {l.split()[0]: float(l.split()[1]) for l in open(os.path.join(tmp, "lm.txt")) if l.strip()}
        ac = {l.split()[0]: float(l.split()[1]) for l in open(os.path.join(tmp, "ac.txt")) if l.strip()}
    return parse_lattice_text(raw.decode()), {k: (words[k], lm[k], ac[k]) for k in words}


# ----------------------------------------------------------------------------------------------
# fuzzy matcher (rhasspy_speech/transcribe_util.py:11-88) through the reference's OpenFst (oracle/fuzzy_probe.cc)


def fuzzy_available() -> bool:
    return os.path.isfile(os.path.join(BIN, "fuzzy-probe"))


def fuzzy_compile(text_fst: str, words_txt: str, out_fst: str):
    """fstcompile --isymbols=words.txt --osymbols=words.txt --keep_isymbols --keep_osymbols (kaldi.py:390-407)."""
    run("fuzzy-probe compile %s %s %s" % (text_fst, words_txt, out_fst))


def fuzzy_reference(nbest: Sequence[Sequence[int]], g_fuzzy_fst: str, words_txt: str) -> Optional[Tuple[str, float]]:
    """get_fuzzy_text restated around the probe: the text FST exactly as hassil_fst.Fst.write prints it
    (one chain per hypothesis, every arc weighted with the running penalty), then the parse of fstprint's output."""
    lines, finals = [], []
    penalty = 0
    nstate = 0
    for hyp in nbest:
        state = 0
        for w in hyp:
            nstate += 1
            lines.append("%d %d %s %s %s" % (state, nstate, w, w, penalty))
            state = nstate
        finals.append(state)
        penalty += 0.1
    for s in dict.fromkeys(finals):
        lines.append(str(s))
    out, _ = run("fuzzy-probe fuzzy %s %s" % (g_fuzzy_fst, words_txt), input=("\n".join(lines) + "\n").encode())
    words: List[str] = []
    cost = 0.0
    for line in out.decode().splitlines():
        parts = line.strip().split()
        if len(parts) < 4:
            continue
        if len(parts) > 4:
            cost += float(parts[4])
        if parts[3] == "<eps>":
            continue
        words.append(parts[3])
    return (" ".join(words), cost) if words else None
