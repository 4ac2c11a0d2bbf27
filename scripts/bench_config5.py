"""BASELINE config 5 through the product's multi-device pool: 2048 utterances round-robin over 8 (model, HCLG) pairs that
differ the way the reference's language models do, every pair served by one transcriber over ALL visible GPUs
(KaldiNnet3WavTranscriber(device="all"): one engine per device, request list dealt longest-first by shard.py, shares run
concurrently, no collective).  Prints one JSON line: whole-job RTFx end to end from WAV paths to strings, and the same
job on one device for the scaling efficiency.    python scripts/bench_config5.py [n_utts | weak]
"weak" = 2048 utterances per visible GPU (fixed work per device; the one-device run then decodes its 2048 only).
"procs" = one PROCESS per visible GPU (the layout bench.py uses under torchrun), each with its own 8 transcribers and
2048 utterances, started together through a file barrier: whole-job RTFx = all audio / slowest rank's mean wall time."""
import asyncio, dataclasses, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rhasspy_speech_b200 as pkg
from rhasspy_speech_b200 import _lib
from tools import synth


def procs_mode():
    """Parent: one child per GPU (CUDA_VISIBLE_DEVICES), file barrier, aggregate of the children's JSON lines."""
    import subprocess
    n_dev = _lib.device_count()
    bar = tempfile.mkdtemp(prefix="c5bar_")
    kids = []
    for r in range(n_dev):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(r), RS_C5_BARRIER=bar, RS_C5_RANK=str(r),
                   RS_B200_PACK_THREADS=str(max(1, min(3, (os.cpu_count() or 8) // n_dev - 1))))
        kids.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "2048"], env=env, stdout=subprocess.PIPE, text=True))
    t0 = time.time()
    while sum(os.path.exists(os.path.join(bar, "ready%d" % r)) for r in range(n_dev)) < n_dev:
        if any(k.poll() is not None for k in kids) or time.time() - t0 > 600:
            raise SystemExit("a rank ended before the barrier")
        time.sleep(0.05)
    open(os.path.join(bar, "go"), "w").close()
    lines = [json.loads([l for l in k.communicate()[0].splitlines() if l.startswith("{")][-1]) for k in kids]
    audio = sum(l["audio_s"] for l in lines)
    wall = max(l["wall_ms_mean"] for l in lines)
    print(json.dumps({"config": "5: %d processes (one per GPU) x 2048 utterances over 8 (model, HCLG) pairs each, WAV paths -> strings" % n_dev,
                      "n_gpus": n_dev, "audio_s": audio, "wall_ms_slowest_rank_mean_of_3": wall, "rtfx_e2e_paths_to_strings": audio / wall * 1e3,
                      "per_rank_rtfx": [round(l["audio_s"] / l["wall_ms_mean"] * 1e3) for l in lines],
                      "decoded": sum(l["decoded"] for l in lines)}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "procs":
        return procs_mode()
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
        bar = os.environ.get("RS_C5_BARRIER")
        if bar:                                 # "procs" mode: every rank starts its timed repetitions together
            asyncio.run(job())
            open(os.path.join(bar, "ready" + os.environ["RS_C5_RANK"]), "w").close()
            while not os.path.exists(os.path.join(bar, "go")):
                time.sleep(0.0005)
        async def timed():                      # one event loop for the repetitions (a service keeps its loop)
            walls, out = [], None
            for _ in range(3):
                t0 = time.perf_counter()
                out = await job()
                walls.append(time.perf_counter() - t0)
            return walls, out
        walls, out = asyncio.run(timed())
        run.mean = float(np.mean(walls))
        return float(np.min(walls)), out
    w_all, out_all = run("all")
    line = {"wall_ms_mean": run.mean * 1e3, "config": "5: %d utterances over 8 (model, HCLG) pairs, product pool over %d GPU(s)" % (n_utts, n_dev), "audio_s": audio_s,
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
