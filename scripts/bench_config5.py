"""BASELINE config 5 through the product's multi-device pool: 2048 utterances round-robin over 8 (model, HCLG) pairs that
differ the way the reference's language models do, every pair served by one transcriber over ALL visible GPUs
(KaldiNnet3WavTranscriber(device="all"): one engine per device, request list dealt longest-first by shard.py, shares run
concurrently, no collective).  Prints one JSON line: whole-job RTFx end to end from WAV paths to strings, and the same
job on one device for the scaling efficiency.    python scripts/bench_config5.py [n_utts | weak]
"weak" = 2048 utterances per visible GPU (fixed work per device; the one-device run then decodes its 2048 only)."""
import asyncio, dataclasses, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rhasspy_speech_b200 as pkg
from rhasspy_speech_b200 import _lib
from tools import synth


def main():
    weak = len(sys.argv) > 1 and sys.argv[1] == "weak"
    n_utts = 2048 * _lib.device_count() if weak else int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    tmp = tempfile.mkdtemp()
    Z = synth.ZAMIA_LIKE
    variants = [Z, dataclasses.replace(Z, name="v1", seed=11), dataclasses.replace(Z, name="v2", seed=12, priors=True),
                dataclasses.replace(Z, name="v3", seed=13, lda_bias=True), dataclasses.replace(Z, name="v4", seed=14, nnet_cmvn=True),
                dataclasses.replace(Z, name="v5", seed=15, num_gauss=256, ivector_dim=60),
                dataclasses.replace(Z, name="v6", seed=16, sentences=synth.EN_US_SENTENCES[:20]),
                dataclasses.replace(Z, name="v7", seed=17, binary=False)]
    models = [synth.write_model(os.path.join(tmp, "m%d" % i), s) for i, s in enumerate(variants)]
    utts = synth.make_utterances(n_utts, seed=1234, min_s=0.8, max_s=3.3, pool=synth.load_pool())   # fixture durations
    audio_s = sum(len(u) for u in utts) / 16000.0
    wavs = []
    for i, pcm in enumerate(utts):
        w = os.path.join(tmp, "u%05d.wav" % i)
        synth.write_wav(w, pcm)
        wavs.append(w)
    parts = [wavs[i::len(models)] for i in range(len(models))]
    n_dev = _lib.device_count()

    def run(device, parts=parts):
        trs = [pkg.KaldiNnet3WavTranscriber(p.model_dir, os.path.dirname(p.hclg), None, device=device) for p in models]

        async def job():
            return await asyncio.gather(*[t.async_transcribe_many(pt, tmp) for t, pt in zip(trs, parts)])
        asyncio.run(job())                      # loads every replica, warms the kernels
        walls = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = asyncio.run(job())
            walls.append(time.perf_counter() - t0)
        return float(np.min(walls)), out
    w_all, out_all = run("all")
    line = {"config": "5: %d utterances over 8 (model, HCLG) pairs, product pool over %d GPU(s)" % (n_utts, n_dev), "audio_s": audio_s,
            "n_gpus": n_dev, "wall_ms": w_all * 1e3, "rtfx_e2e_paths_to_strings": audio_s / w_all,
            "decoded": int(sum(1 for part in out_all for o in part if o))}
    if n_dev > 1 and weak:
        # fixed work per device: one device decodes the first 2048 utterances, the pool all of them
        sub = wavs[:2048]
        sub_audio = sum(len(u) for u in utts[:2048]) / 16000.0
        w_one, out_one = run(0, [sub[i::len(models)] for i in range(len(models))])
        line.update({"scaling": "weak", "wall_ms_one_gpu_2048": w_one * 1e3, "rtfx_one_gpu": sub_audio / w_one,
                     "efficiency": (audio_s / w_all) / (n_dev * sub_audio / w_one)})
    elif n_dev > 1:
        w_one, out_one = run(0)
        line.update({"scaling": "strong", "wall_ms_one_gpu": w_one * 1e3, "rtfx_one_gpu": audio_s / w_one, "speedup": w_one / w_all,
                     "efficiency": w_one / w_all / n_dev, "identical_to_one_gpu": out_one == out_all})
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
