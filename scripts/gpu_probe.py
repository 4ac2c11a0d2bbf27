"""Scratch probe for the GPU box: stage timings and parity on the zamia-like model."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rhasspy_speech_b200 import _lib
from tools import synth
from oracle import ref_run

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tmp = tempfile.mkdtemp()
p = synth.write_model(tmp, synth.ZAMIA_LIKE)
utts = synth.make_utterances(n, seed=1234)
t0 = time.time()
model = _lib.Model(p.final_mdl, p.online_conf, 0)
graph = _lib.Graph(p.hclg, p.words_txt, 0)
dec = _lib.Decoder(model, graph)
print("load s", time.time() - t0, flush=True)
print(model.plan())
for it in range(3):
    t0 = time.time()
    hyp = dec.decode_pcm(utts)
    wall = time.time() - t0
    t = dec.timings()
    print("iter", it, "wall", round(wall, 4), json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items()}), flush=True)
print("status", np.bincount(hyp.status), "rtfx(total_ms)", t["audio_seconds"] / (t["total_ms"] / 1e3))
if ref_run.available():
    k = min(n, 8)
    wavs = []
    for u in range(k):
        w = os.path.join(tmp, "u%d.wav" % u); synth.write_wav(w, utts[u]); wavs.append(w)
    t0 = time.time()
    want, _, err = ref_run.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, wavs)
    print("ref wall", time.time() - t0, err.decode()[-300:])
    bad = 0
    for u in range(k):
        ok = want.get("utt%05d-1" % u) == hyp.words[u]
        bad += not ok
        print(u, ok, hyp.words[u], want.get("utt%05d-1" % u))
    feats = [dec.fetch(0, u) for u in range(k)]
    ivs = [dec.fetch(1, u)[0] for u in range(k)]
    ll = ref_run.nnet_loglikes(p.final_mdl, feats, ivs, frame_subsampling_factor=3)
    for u in range(k):
        got = dec.fetch(2, u)
        f64 = synth.nnet_forward(p.nnet_params, feats[u].astype(np.float64), ivs[u].astype(np.float64))[::3]
        print("ll", u, got.shape, "vs kaldi", np.abs(got - ll[u]).max(), "gpu vs f64", np.abs(got - f64).max(), "kaldi vs f64", np.abs(ll[u] - f64).max())
    print("mismatches", bad)
