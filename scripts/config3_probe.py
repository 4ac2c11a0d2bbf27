"""Config 3 (ARPA-shaped HCLG, batch 256, every utterance re-decoded by the strict-order host decoder): wall time and
the host decoder's share, two repetitions.    python scripts/config3_probe.py"""
import dataclasses, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rhasspy_speech_b200 import _lib
from tools import synth
tmp = tempfile.mkdtemp()
spec = dataclasses.replace(synth.ZAMIA_LIKE, name="zamia_arpa", graph="arpa", vocab_size=2000, bigrams_per_word=20, eps_hops=2)
p = synth.write_model(os.path.join(tmp, "arpa"), spec)
utts = synth.make_utterances(256, seed=1234, pool=synth.load_pool())
for i in range(0, 256, 10):
    utts[i] = utts[i][::-1].copy()
dec = _lib.Decoder(_lib.Model(p.final_mdl, p.online_conf, 0), _lib.Graph(p.hclg, p.words_txt, 0), max_tokens_per_utt=1 << 21)
audio_s = sum(len(u) for u in utts) / 16000.0
dec.decode_pcm(utts)
out = []
for _ in range(2):
    t0 = time.perf_counter()
    h = dec.decode_pcm(utts)
    w = time.perf_counter() - t0
    t = dec.timings()
    out.append({"wall_ms": w * 1e3, "strict_ms": t["strict_ms"], "strict_utts": t["strict_utts"], "decode_ms": t["decode_ms"], "rtfx_e2e": audio_s / w})
print(json.dumps({"config": "3: ARPA-shaped HCLG, batch 256, strict-order host decoder on every flagged utterance", "audio_s": audio_s, "runs": out}))
