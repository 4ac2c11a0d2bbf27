#!/usr/bin/env python3
"""Benchmark of the hot path: batched utterance decoding (RTFx = audio seconds / wall seconds).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): batch = 256 utterances of 3-5 s, 16 kHz, zamia-shaped TDNN-F
chain model (1024/128, 13 layers, 3026 pdfs, 100-dim iVectors, frame-subsampling 3) and the
en_US grammar HCLG, all synthetic and seeded (rhasspy_speech_b200/synth.py; the real en_US-zamia
artefacts are a download and there is no network).  One step = one pass of the whole hot path
(MFCC + iVector -> TDNN-F -> token passing -> best-path word ids) over the batch.  With N > 1
(torchrun, one rank per GPU) every rank decodes its own 256 utterances: weak scaling, no collective
on the data path.

  value  : RTFx with the audio already resident in HBM -- sum of the stage times measured by CUDA
           events on the decoder's stream inside the library (feature + nnet + decode), max over ranks.
  e2e    : RTFx through the C ABI call a user makes (rs_decode_pcm) with host buffers: host staging,
           H2D of the PCM, all kernels, D2H of the word ids, host-side result construction.
  --impl reference : the reference's own Kaldi CPU path (oracle/_ref binaries, the exact argv of
           rhasspy_speech/transcribe_wav.py:47-74) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BATCH = 256
# DRAM bytes the 30 launches of the nnet stage moved in one step under ncu (profiles/r2g_launches.csv, round 2;
# round 1: 10.45 GB)
NNET_DRAM_BYTES_PER_STEP = 9391758592
METRIC = "RTFx (audio-sec/wall-sec) en_US-zamia 16kHz at 1/2/4/8 B200; WER vs ref"
UNIT = "audio-sec/wall-sec"
WORKLOAD = "configs[1]: batch=256 per GPU, grammar-HCLG, 3-5 s 16 kHz utterances cut from tests/en_US-zamia WAVs (+ sigma=2 noise)"
DATA = "synthetic (seeded random-weight model and lexicon; audio = the reference's en_US-zamia fixture WAVs, concatenated)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                parts = [x.strip() for x in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for nme, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def make_workload(tmp, n_utts, seed):
    from tools import synth
    p = synth.write_model(tmp, synth.ZAMIA_LIKE)
    # BASELINE config 2 audio: 3-5 s utterances cut from the reference's tests/en_US-zamia WAVs (committed under
    # tests/golden/en_US-zamia), seed 1234, + sigma = 2 LSB noise (seed 5678) so that the lanes differ
    utts = synth.make_utterances(n_utts, seed=seed, pool=synth.load_pool())
    return p, utts


def _write_wavs(tmp, utts):
    from tools import synth
    wavs = []
    for i, pcm in enumerate(utts):
        w = os.path.join(tmp, "u%04d.wav" % i)
        synth.write_wav(w, pcm)
        wavs.append(w)
    return wavs


def _kaldi_rtf(err_log: bytes):
    """(decode seconds, audio seconds) from Kaldi's own timing line (online2/online-timing.cc:53-60):
    'real-time factor for offline decoding was X = a / b seconds' -- model load and nnet3 compilation excluded."""
    import re
    m = re.search(rb"real-time factor[^=]*=\s*([0-9.eE+-]+)\s*seconds\s*/\s*([0-9.eE+-]+)\s*seconds", err_log)
    return (float(m.group(1)), float(m.group(2))) if m else None


def reference_warm_pass(p, wavs, cores):
    """B-warm of BASELINE.md section 3: `cores` processes, each one online2-wav-nnet3-latgen-faster (+ lattice-to-nbest |
    nbest-to-linear) over a 1/cores shard of the batch through real spk2utt / wav.scp files, OPENBLAS_NUM_THREADS=1.
    Returns (wall seconds, transcripts by utterance index, max over processes of Kaldi's own decode seconds)."""
    from oracle import ref_run
    shards = [list(range(c, len(wavs), cores)) for c in range(cores)]
    outs = [None] * cores

    def work(c):
        if shards[c]:
            outs[c] = ref_run.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, [wavs[i] for i in shards[c]])
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(c,)) for c in range(cores)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    words = [None] * len(wavs)
    kaldi_s = []
    for c in range(cores):
        if outs[c] is None:
            continue
        hyp, _, err = outs[c]
        for k, i in enumerate(shards[c]):
            words[i] = hyp.get("utt%05d-1" % k)
        r = _kaldi_rtf(err)
        if r:
            kaldi_s.append(r[0])
    return wall, words, (max(kaldi_s) if kaldi_s else None)


def reference_cold_pass(p, wavs, cores):
    """B-cold ("as shipped"): one 3-process pipeline per utterance with the exact argv of transcribe_wav.py:45-75 -- model
    load and nnet3 compilation per call -- `cores` concurrent workers."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref_run
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(lambda w: ref_run.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, [w]), wavs))
    return time.perf_counter() - t0


def run_reference(args):
    """The reference's own Kaldi CPU path on this box's host cores, on the bench workload: every step is one B-warm pass over
    the whole 256-utterance batch (256 / cores utterances per process)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_run
    if not ref_run.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (run oracle/build_ref.py)"}))
        return
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        p, utts = make_workload(tmp, BATCH, 1234)
        wavs = _write_wavs(tmp, utts)
        audio_s = sum(len(u) for u in utts) / 16000.0
        warmup = min(args.warmup, 1)
        for _ in range(warmup):
            reference_warm_pass(p, wavs[:cores], cores)       # page the binaries and the model files in
        steps = max(1, min(args.steps, 3))
        runs = [reference_warm_pass(p, wavs, cores) for _ in range(steps)]
        dt = float(np.mean([r[0] for r in runs]))
        kaldi = [r[2] for r in runs if r[2]]
        n_cold = min(len(wavs), 2 * cores)
        cold_s = reference_cold_pass(p, wavs[:n_cold], cores)
        cold_audio = sum(len(u) for u in utts[:n_cold]) / 16000.0
    value = audio_s / dt
    sample = ("B-warm: the whole bench batch, %d utterances (%.0f s of audio), %d processes x %d utterances, wall clock from the "
              "first process start to the last exit (model load + nnet3 compile once per process)" % (BATCH, audio_s, cores, BATCH // cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": DATA,
        "config": {"workload": WORKLOAD, "model": "zamia-like TDNN-F (synthetic)", "audio_seconds_per_step": audio_s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "kaldi_rtf_rtfx": (audio_s / float(np.mean(kaldi))) if kaldi else None,
        "kaldi_rtf_note": "audio seconds / the slowest process's own 'real-time factor for offline decoding' seconds (model load excluded)",
        "b_cold": {"value": cold_audio / cold_s, "unit": UNIT, "utterances": n_cold, "workers": cores,
                   "how": "one 3-process pipeline per utterance as transcribe_wav.py:45-75 spawns it (model load per call)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def edit_distance(a, b):
    """Word-level Levenshtein distance."""
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def parity_and_cpu_baseline(p, utts, hyp, tmp):
    """Outside the timed region (rank 0, N = 1): one B-warm pass of the reference over the SAME batch gives the CPU
    baseline and the transcripts the timed step's hypotheses are compared with (the metric's 'WER vs ref')."""
    from oracle import ref_run
    if not ref_run.available():
        return None, None
    cores = os.cpu_count() or 1
    wavs = _write_wavs(tmp, utts)
    audio_s = sum(len(u) for u in utts) / 16000.0
    wall, ref_words, kaldi_s = reference_warm_pass(p, wavs, cores)
    mism, errs, nref = 0, 0, 0
    for u in range(len(utts)):
        got, want = hyp.words[u] or [], ref_words[u] or []
        mism += got != want
        errs += edit_distance(got, want)
        nref += len(want)
    flags = {}
    for st in hyp.status:
        flags[int(st)] = flags.get(int(st), 0) + 1
    parity = {"utterances": len(utts), "mismatches": mism, "reference_words": nref, "word_errors_vs_reference": errs,
              "wer_delta": errs / max(nref, 1), "status_counts": flags,
              "how": "hypotheses of the last timed step vs online2-wav-nnet3-latgen-faster | lattice-to-nbest | nbest-to-linear "
                     "(oracle/_ref) on the same 256 WAVs, outside the timed region; log-likelihood and path-cost parity on "
                     "this model: tests/test_gpu_zamia.py"}
    cb = {"value": audio_s / wall, "unit": UNIT, "cores": cores, "kind": "reference",
          "sample": "B-warm over the whole batch: %d utterances (%.0f s audio), %d processes x %d utterances, model load included"
                    % (len(utts), audio_s, cores, len(utts) // cores),
          "kaldi_rtf_rtfx": (audio_s / kaldi_s) if kaldi_s else None}
    return parity, cb


def e2e_api(p, utts, tmp, reps=6):
    """The reference's API shape: 256 WAV *paths* -> strings through KaldiNnet3WavTranscriber.async_transcribe_many
    (file reads, H2D, kernels, D2H, word-id -> text, decode_meta)."""
    import asyncio
    import rhasspy_speech_b200 as pkg
    wavs = _write_wavs(tmp, utts)
    lang_dir = os.path.join(tmp, "lang")
    os.makedirs(lang_dir, exist_ok=True)
    tr = pkg.KaldiNnet3WavTranscriber(p.model_dir, os.path.dirname(p.hclg), None)

    async def job(n):       # one event loop for all repetitions, as a long-lived service has (loop set-up is not the call)
        walls, out = [], None
        for _ in range(n):
            t0 = time.perf_counter()
            out = await tr.async_transcribe_many(wavs, lang_dir)
            walls.append(time.perf_counter() - t0)
        return walls, out
    asyncio.run(job(2))
    walls, out = asyncio.run(job(reps))
    return float(np.mean(walls)), sum(1 for o in out if o)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    # host staging threads per rank: share the box's cores between the ranks
    os.environ.setdefault("RS_B200_PACK_THREADS", str(max(1, min(3, (os.cpu_count() or 8) // max(world, 1) - 1))))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from rhasspy_speech_b200 import _lib
    tmp = tempfile.mkdtemp(prefix="rsbench%d_" % rank)
    p, utts = make_workload(tmp, BATCH, 1234 + rank)
    model = _lib.Model(p.final_mdl, p.online_conf, local)
    graph = _lib.Graph(p.hclg, p.words_txt, local)
    dec = _lib.Decoder(model, graph)
    audio_s = sum(len(u) for u in utts) / 16000.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # the step's inputs start in page-locked host memory (one rs_host_alloc block, utterances back to back); the same
    # call from ordinary numpy arrays (one more host memcpy into the decoder's staging area) is timed as e2e_pageable
    utts_pageable = utts
    pinned = _lib.PinnedAudio.from_utterances(utts)
    utts = pinned

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        hyp = dec.decode_pcm(utts)
    assert all(int(s) & 15 == 0 for s in hyp.status), "decoder reported capacity problems"
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    stage_ms, wall_s, launches = [], [], 0
    # `value`: the K steps with the batch's audio copied first and every kernel behind it on one stream, so that the
    # stage times (feature + nnet + decode, CUDA events on the decoder's stream) are those of inputs resident in HBM
    dec.set_staging_overlap(False)
    dec.decode_pcm(utts)
    barrier()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (256 MiB > 126 MB L2)
        torch.cuda.synchronize()
        hyp = dec.decode_pcm(utts)
        t = dec.timings()
        stage_ms.append((t["feature_ms"], t["nnet_ms"], t["decode_ms"], t["h2d_ms"], t["d2h_ms"], t["total_ms"]))
        launches += t["kernel_launches"]
    # `e2e`: the same K steps as a user calls them (default staging: the audio travels in a few items on a copy stream
    # and the MFCC kernel of an item runs under the copy of the next), wall clock around the call
    dec.set_staging_overlap(True)
    dec.decode_pcm(utts)
    barrier()
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hyp = dec.decode_pcm(utts)
        wall_s.append(time.perf_counter() - t0)
    barrier()
    t_all = time.perf_counter() - t_all0
    wall_pageable = []
    for _ in range(max(2, args.steps // 2)):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec.decode_pcm(utts_pageable)
        wall_pageable.append(time.perf_counter() - t0)
    barrier()
    # the same call with two batches in flight per GPU (two decoders, two host threads): the host staging
    # and result assembly of one batch overlap the kernels of the other.  Every batch still pays its own
    # pinned staging, H2D, kernels and D2H inside the timed region.
    dec2 = _lib.Decoder(model, graph)
    for _ in range(2):
        dec2.decode_pcm(utts)
    n_pipe = max(args.steps, 4)

    def worker(dx, k):
        for _ in range(k):
            dx.decode_pcm(utts)
    barrier()
    tp0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(dx, (n_pipe + 1) // 2)) for dx in (dec, dec2)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join()
    torch.cuda.synchronize()
    pipe_s = (time.perf_counter() - tp0) / (2 * ((n_pipe + 1) // 2))
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    sm = np.asarray(stage_ms)
    dev_s = float((sm[:, 0] + sm[:, 1] + sm[:, 2]).mean() / 1e3)
    e2e_s = float(np.mean(wall_s))
    # max over ranks (device time and wall time), sum of audio
    vec = torch.tensor([dev_s, e2e_s, pipe_s, float(np.mean(wall_pageable))], dtype=torch.float64, device="cuda")
    aud = torch.tensor([audio_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(aud, op=dist.ReduceOp.SUM)
    dev_s_max, e2e_s_max, pipe_s_max, pageable_s_max = [float(x) for x in vec.tolist()]
    audio_total = float(aud.item())
    if rank == 0:
        pk, pk_kind = peaks()
        t = dec.timings()
        nnet_ms = float(sm[:, 1].mean())
        f16_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])   # fp16 and bf16 share the pipe and the rate
        achieved = t["nnet_flops"] / (nnet_ms / 1e3) / 1e12
        hbm_achieved = t["nnet_bytes"] / (nnet_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": audio_total / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_s_max * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tensor-core products as 3 x fp16 split terms, fp32 accumulate)", "data": DATA,
            "config": {"workload": WORKLOAD,
                       "model": "zamia-like TDNN-F chain (synthetic, seeded): 40-dim hires MFCC + 100-dim iVector, 1024/128 x 12 TDNN-F, 3026 pdfs, sf=3",
                       "graph": "en_US grammar HCLG (synthetic lexicon), %d states" % graph.num_states,
                       "decoder": "beam 24, max-active 7000, lattice-beam 8 (best path; the reference's token order reproduced on the device)", "l2": "flushed between iterations (256 MiB write)",
                       "audio_seconds_per_step": audio_total},
            "e2e": {"value": audio_total / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(t["h2d_bytes"]),
                    "d2h_bytes_per_step": int(t["d2h_bytes"]), "ms_per_step": e2e_s_max * 1e3,
                    "input": "int16 PCM in one page-locked host block (rs_host_alloc), copied H2D inside the timed call "
                             "(4 items on a copy stream, the MFCC kernel of an item under the copy of the next)"},
            "e2e_pageable": {"value": audio_total / pageable_s_max, "unit": UNIT, "ms_per_step": pageable_s_max * 1e3,
                             "input": "the same call from 256 ordinary numpy arrays: packed into pinned staging first"},
            "e2e_two_in_flight": {"value": audio_total / pipe_s_max, "unit": UNIT, "ms_per_batch": pipe_s_max * 1e3,
                                  "how": "two decoders per GPU driven by two host threads, same call, same per-batch copies"},
            "gpu_launches": int(launches),
            "value_how": "feature + nnet + decode stage times of K separate steps run with rs_decoder_set_staging_overlap(0): "
                         "all copies first, every kernel behind them on one stream",
            "stages_ms": {"feature": float(sm[:, 0].mean()), "nnet": nnet_ms, "decode": float(sm[:, 2].mean()),
                          "h2d": float(sm[:, 3].mean()), "d2h": float(sm[:, 4].mean())},
            # dominant kernel = gemm_tc3_kernel (30 launches per step, the whole nnet stage between two CUDA
            # events on the decoder's stream).  achieved = algorithmic FLOPs (2*rows*K*N per layer, counted
            # once although every product is 3 MMAs => attainable frac <= 1/3) / stage time.
            "roofline": {"bound": "tensor", "kernel": "gemm_tc3_kernel (TDNN-F affine layers, 30 launches = the nnet stage)",
                         "achieved": achieved, "peak": f16_peak, "unit": "TFLOP/s", "frac": achieved / f16_peak,
                         "traffic": NNET_DRAM_BYTES_PER_STEP,
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the stage's launches, "
                                           "profiles/r2g_launches.csv (same workload, batch 256)",
                         "peak_source": pk_kind + " bf16_tflops_sustained (fp16 tensor pipe)",
                         "note": "3 MMAs per algorithmic product: attainable frac <= 0.333"},
            # the same launches against HBM: the stage streams every layer's activations through HBM once
            "roofline_hbm": {"bound": "hbm", "kernel": "gemm_tc3_kernel (same 30 launches)", "achieved": hbm_achieved,
                             "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_achieved / pk["hbm_gbs"],
                             "traffic": NNET_DRAM_BYTES_PER_STEP, "algorithmic_bytes": int(t["nnet_bytes"])},
            "decoder_counters": {k: int(t[k]) for k in ("frames_decoded", "tokens_expanded", "arcs_visited", "tokens_created")},
            "clocks": sampler.summary(),
            "wall_s_total": t_all,
        }
        line["strict_host_decoder"] = {"utterances_per_step": int(t["strict_utts"]), "ms_per_step": float(t["strict_ms"])}
        if world == 1 and not args.no_cpu_baseline:
            parity, cb = parity_and_cpu_baseline(p, utts_pageable, hyp, tmp)
            if cb:
                line["cpu_baseline"] = cb
                line["parity"] = parity
            api_s, api_n = e2e_api(p, utts_pageable, tmp)
            line["e2e_api"] = {"value": audio_total / api_s, "unit": UNIT, "ms_per_step": api_s * 1e3, "transcripts": api_n,
                               "how": "256 WAV paths -> strings through KaldiNnet3WavTranscriber.async_transcribe_many (file reads, "
                                      "H2D, kernels, D2H, word ids -> text)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
