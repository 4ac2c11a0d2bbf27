"""Stage times and wall time of the bench batch from page-locked input with the staging overlap on / off.
    python scripts/overlap_probe.py [batch]"""
import os, sys, tempfile, time
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
pinned = _lib.PinnedAudio.from_utterances(utts)
for src, name in ((pinned, "pinned"), (utts, "pageable")):
    for on in (False, True, False, True):
        dec.set_staging_overlap(on)
        dec.decode_pcm(src)
        walls, tots, h2d, feat = [], [], [], []
        for it in range(8):
            t0 = time.perf_counter()
            dec.decode_pcm(src)
            walls.append((time.perf_counter() - t0) * 1e3)
            t = dec.timings()
            tots.append(t["total_ms"]); h2d.append(t["h2d_ms"]); feat.append(t["feature_ms"])
        f = lambda v: round(sorted(v)[len(v) // 2], 3)
        print(name, "overlap", on, "wall", f(walls), "device total", f(tots), "h2d", f(h2d), "feature", f(feat), "launches", t["kernel_launches"])
