#!/usr/bin/env python3
"""Generate tests/golden/golden.npz by running the REFERENCE (oracle/_ref, built from /root/reference).

Run in the build container (needs /root/reference for the fixture WAVs and oracle/_ref):

    python tests/golden/make_golden.py

Contents (all small): four utterances (three clipped reference fixtures from tests/en_US-zamia and
one synthetic), and for each the reference's own outputs on the seeded TINY synthetic model:
  mfcc_*      compute-mfcc-feats                                  (feature-mfcc.cc)
  ivp_*       ivector-extract-online2 (periodic schedule)         (online-ivector-feature.cc)
  ll_*        nnet3-compute --frame-subsampling-factor=3, fed the oracle's offline iVector
  words_*     online2-wav-nnet3-latgen-faster | lattice-to-nbest | nbest-to-linear (the transcribe_wav.py argv)
  stream_*    online2-cli-nnet3-decode-faster fed raw s16le (the transcribe_stream.py argv)
  win, fft    windowed frames and packed split-radix spectra of 16 frames (oracle/ref_probe.cc)
plus the sha256 of the generated final.mdl / HCLG.fst so a drifting generator is detected.
"""
import glob
import hashlib
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import kaldi_np as K  # noqa: E402
from oracle import ref_run  # noqa: E402
from tools import synth  # noqa: E402


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def main():
    assert ref_run.available(), "build oracle/_ref first (python oracle/build_ref.py)"
    out = {}
    wavs = sorted(glob.glob("/root/reference/tests/en_US-zamia/*.wav"))
    picks = [wavs[0], wavs[17], wavs[40]]
    utts = [synth.read_wav(w)[0][:20000] for w in picks] + [synth.synth_speech(1.3, 4242)]
    with tempfile.TemporaryDirectory() as tmp:
        p = synth.write_model(tmp, synth.TINY)
        conf = os.path.join(p.model_dir, "model", "online", "conf")
        paths = []
        for i, pcm in enumerate(utts):
            w = os.path.join(tmp, "g%d.wav" % i)
            synth.write_wav(w, pcm)
            paths.append(w)
            out["pcm_%d" % i] = pcm.astype(np.int16)
        feats = ref_run.mfcc(os.path.join(conf, "mfcc.conf"), paths)
        ivp = ref_run.ivectors_periodic(os.path.join(conf, "ivector_extractor.conf"), feats)
        s = K.IvectorSetup.from_conf(os.path.join(conf, "ivector_extractor.conf"))
        ivo = [K.ivector_offline(s, f) for f in feats]
        ll = ref_run.nnet_loglikes(p.final_mdl, feats, ivo, frame_subsampling_factor=3)
        words, _, _ = ref_run.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, paths)
        for i in range(len(utts)):
            out["mfcc_%d" % i] = feats[i]
            out["ivp_%d" % i] = ivp[i]
            out["ivo_%d" % i] = ivo[i]
            out["ll_%d" % i] = ll[i]
            out["words_%d" % i] = np.asarray(words.get("utt%05d-1" % i, []), dtype=np.int32)
            sw, _ = ref_run.transcribe_stream(p.final_mdl, p.online_conf, p.hclg, p.words_txt, utts[i])
            out["stream_%d" % i] = np.asarray(sw.get("utt-1", []), dtype=np.int32)
        out["sha_final_mdl"] = np.frombuffer(sha(p.final_mdl).encode(), dtype=np.uint8)
        out["sha_hclg"] = np.frombuffer(sha(p.hclg).encode(), dtype=np.uint8)
        # FFT probe
        n = 16
        frames = np.stack([utts[0][2000 + i * 160:2000 + i * 160 + 400] for i in range(n)]).astype(np.float32)
        env = dict(os.environ, LD_LIBRARY_PATH=ref_run.REF_DIR)
        raw = subprocess.run([os.path.join(ref_run.BIN, "ref-probe")], input=struct.pack("<i", n) + frames.tobytes(),
                             stdout=subprocess.PIPE, env=env, check=True).stdout
        a = np.frombuffer(raw, np.float32).reshape(n, 512 + 512 + 40)
        out["probe_frames"] = frames.astype(np.int16)
        out["probe_win"] = a[:, :512].copy()
        out["probe_fft"] = a[:, 512:1024].copy()
        out["probe_mfcc"] = a[:, 1024:].copy()
    dst = os.path.join(ROOT, "tests", "golden", "golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    for i in range(len(utts)):
        print(i, out["words_%d" % i].tolist(), out["stream_%d" % i].tolist())


if __name__ == "__main__":
    main()
