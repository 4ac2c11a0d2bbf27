"""Where the wall time of the multi-device pool goes: start / end of every rs_decode_wavs call (per device) and of the
whole job, config 5 workload.    python scripts/debug_pool.py [n_utts]"""
import asyncio, dataclasses, json, os, sys, tempfile, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rhasspy_speech_b200 as pkg
from rhasspy_speech_b200 import _lib
from tools import synth
n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
tmp = tempfile.mkdtemp()
Z = synth.ZAMIA_LIKE
variants = [dataclasses.replace(Z, name="v%d" % i, seed=10 + i) for i in range(8)]
models = [synth.write_model(os.path.join(tmp, "m%d" % i), s) for i, s in enumerate(variants)]
utts = synth.make_utterances(n_utts, seed=1234, min_s=0.8, max_s=3.3, pool=synth.load_pool())
wavs = []
for i, pcm in enumerate(utts):
    w = os.path.join(tmp, "u%05d.wav" % i)
    synth.write_wav(w, pcm)
    wavs.append(w)
parts = [wavs[i::len(models)] for i in range(len(models))]
log = []
orig = _lib.Decoder.decode_wavs
def timed(self, paths):
    t0 = time.perf_counter()
    r = orig(self, paths)
    t1 = time.perf_counter()
    log.append((self.model.device if hasattr(self.model, "device") else -1, len(paths), t0, t1, dict(self.timings())))
    return r
_lib.Decoder.decode_wavs = timed
trs = [pkg.KaldiNnet3WavTranscriber(p.model_dir, os.path.dirname(p.hclg), None, device="all") for p in models]
async def job():
    return await asyncio.gather(*[t.async_transcribe_many(pt, tmp) for t, pt in zip(trs, parts)])
asyncio.run(job())
for rep in range(2):
    log.clear()
    T0 = time.perf_counter()
    asyncio.run(job())
    T1 = time.perf_counter()
    print("job wall %.1f ms" % ((T1 - T0) * 1e3))
    for dev, n, t0, t1, tm in sorted(log, key=lambda x: x[2]):
        print("  dev %d n=%3d  start %6.1f  end %6.1f  (call %5.1f ms; device feature+nnet+decode %.1f, h2d %.1f, total %.1f)" %
              (dev, n, (t0 - T0) * 1e3, (t1 - T0) * 1e3, (t1 - t0) * 1e3, tm["feature_ms"] + tm["nnet_ms"] + tm["decode_ms"], tm["h2d_ms"], tm["total_ms"]))
