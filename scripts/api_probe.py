"""Where the time of async_transcribe_many goes (256 WAV paths -> strings): the C call against everything around it.
    RS_B200_PACK_THREADS=k python scripts/api_probe.py"""
import asyncio, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rhasspy_speech_b200 as pkg
from rhasspy_speech_b200 import _lib
from tools import synth
tmp = tempfile.mkdtemp()
p = synth.write_model(tmp, synth.ZAMIA_LIKE)
utts = synth.make_utterances(256, seed=1234, pool=synth.load_pool())
wavs = []
for i, pcm in enumerate(utts):
    w = os.path.join(tmp, "u%05d.wav" % i)
    synth.write_wav(w, pcm)
    wavs.append(w)
tr = pkg.KaldiNnet3WavTranscriber(p.model_dir, os.path.dirname(p.hclg), None)
c_ms = []
orig = _lib.Decoder.decode_wavs
def timed(self, paths):
    t0 = time.perf_counter()
    r = orig(self, paths)
    c_ms.append((time.perf_counter() - t0) * 1e3)
    return r
_lib.Decoder.decode_wavs = timed
async def job(reps):
    walls = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = await tr.async_transcribe_many(wavs, tmp)
        walls.append((time.perf_counter() - t0) * 1e3)
    return walls, out
asyncio.run(job(2))
c_ms.clear()
walls, out = asyncio.run(job(8))
med = lambda v: round(sorted(v)[len(v) // 2], 3)
print("pack threads", os.environ.get("RS_B200_PACK_THREADS"), "wall ms", med(walls), "decode_wavs (python wrapper + C call) ms", med(c_ms),
      "transcripts", sum(1 for o in out if o))
