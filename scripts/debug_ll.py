"""GPU box probe: log-likelihood error budget on the bench model (GPU vs nnet3-compute vs an fp64 forward)."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rhasspy_speech_b200 import _lib
from tools import synth
from oracle import ref_run
tmp = tempfile.mkdtemp()
p = synth.write_model(tmp, synth.ZAMIA_LIKE)
n = 8
utts = synth.make_utterances(n, seed=1234, pool=synth.load_pool())
dec = _lib.Decoder(_lib.Model(p.final_mdl, p.online_conf, 0), _lib.Graph(p.hclg, p.words_txt, 0))
dec.decode_pcm(utts)
feats = [dec.fetch(0, u) for u in range(n)]
ivs = [dec.fetch(1, u)[0] for u in range(n)]
ll = ref_run.nnet_loglikes(p.final_mdl, feats, ivs, frame_subsampling_factor=3)
for u in range(n):
    got = dec.fetch(2, u)
    f64 = synth.nnet_forward(p.nnet_params, feats[u].astype(np.float64), ivs[u].astype(np.float64))[::3]
    mag = np.abs(f64)
    e_gk, e_gf, e_kf = np.abs(got - ll[u]), np.abs(got - f64), np.abs(ll[u] - f64)
    print("utt", u, "max|ll| %.1f mean|ll| %.1f" % (mag.max(), mag.mean()), "| gpu-kaldi max %.2e | gpu-f64 max %.2e rms %.2e | kaldi-f64 max %.2e rms %.2e"
          % (e_gk.max(), e_gf.max(), np.sqrt((e_gf ** 2).mean()), e_kf.max(), np.sqrt((e_kf ** 2).mean())),
          "| rel gpu-f64 %.2e kaldi-f64 %.2e" % ((e_gf / np.maximum(mag, 1)).max(), (e_kf / np.maximum(mag, 1)).max()))
    for lo, hi in ((0, 8), (8, 32), (32, 128), (128, 1e9)):
        m = (mag >= lo) & (mag < hi)
        if m.any():
            print("    |ll| in [%g,%g): %d values, gpu-kaldi max %.2e, gpu-f64 max %.2e, kaldi-f64 max %.2e" % (lo, hi, m.sum(), e_gk[m].max(), e_gf[m].max(), e_kf[m].max()))
