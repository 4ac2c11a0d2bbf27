"""A few passes of the bench workload for ncu launch lists / full captures and quick stage timings.
    python scripts/ncu_step.py [batch] [passes] [nbest]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rhasspy_speech_b200 import _lib
from tools import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tmp = tempfile.mkdtemp()
p = synth.write_model(tmp, synth.ZAMIA_LIKE)
utts = synth.make_utterances(n, seed=1234, pool=synth.load_pool())
model = _lib.Model(p.final_mdl, p.online_conf, 0)
graph = _lib.Graph(p.hclg, p.words_txt, 0)
dec = _lib.Decoder(model, graph)
if len(sys.argv) > 3:       # n-best tail: lattice recording + lattice_prune_kernel
    dec.set_nbest(int(sys.argv[3]))
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    hyp = dec.decode_pcm(utts)
    t = dec.timings()
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items() if k.endswith("_ms") or k == "kernel_launches"})
