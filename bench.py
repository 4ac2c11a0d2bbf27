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
# DRAM bytes the nnet stage moved in one step under ncu (profiles/r1_launches_i_dram.csv): 7.39 GB read + 3.06 GB written
NNET_DRAM_BYTES_PER_STEP = 10454000000
METRIC = "RTFx (audio-sec/wall-sec) en_US-zamia 16kHz at 1/2/4/8 B200; WER vs ref"
UNIT = "audio-sec/wall-sec"


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
    utts = synth.make_utterances(n_utts, seed=seed)
    return p, utts


def run_reference(args):
    """The reference's CPU implementation of the path on this box's host cores (bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_run
    from tools import synth
    if not ref_run.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (run oracle/build_ref.py)"}))
        return
    cores = os.cpu_count() or 1
    per_core = 3
    n = cores * per_core
    with tempfile.TemporaryDirectory() as tmp:
        p, utts = make_workload(tmp, n, 1234)
        wavs = []
        for i, pcm in enumerate(utts):
            w = os.path.join(tmp, "u%04d.wav" % i)
            synth.write_wav(w, pcm)
            wavs.append(w)
        audio_s = sum(len(u) for u in utts) / 16000.0

        def one_pass():
            # B-warm of BASELINE.md: one online2-wav-nnet3-latgen-faster per core over a 1/cores shard
            threads, outs = [], [None] * cores
            t0 = time.perf_counter()

            def work(c):
                shard = wavs[c::cores]
                if shard:
                    outs[c] = ref_run.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, shard)[0]
            for c in range(cores):
                t = threading.Thread(target=work, args=(c,))
                t.start()
                threads.append(t)
            for t in threads:
                t.join()
            return time.perf_counter() - t0
        for _ in range(min(args.warmup, 1)):
            one_pass()
        steps = max(1, min(args.steps, 3))
        times = [one_pass() for _ in range(steps)]
        dt = float(np.mean(times))
    value = audio_s / dt
    sample = "%d utterances (%.0f s of audio) of the bench workload, %d processes x %d utterances, model load included" % (
        n, audio_s, cores, per_core)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: batch=256 grammar-HCLG, 3-5 s 16 kHz utterances (bounded sample)", "model": "zamia-like TDNN-F (synthetic)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline_sample():
    """Reference timed on the host cores next to the GPU numbers (rank 0, N=1 only), ~10-30 s of CPU work."""
    from oracle import ref_run
    from tools import synth
    if not ref_run.available():
        return None
    cores = os.cpu_count() or 1
    n = cores * 2
    with tempfile.TemporaryDirectory() as tmp:
        p, utts = make_workload(tmp, n, 1234)
        wavs = []
        for i, pcm in enumerate(utts):
            w = os.path.join(tmp, "u%04d.wav" % i)
            synth.write_wav(w, pcm)
            wavs.append(w)
        audio_s = sum(len(u) for u in utts) / 16000.0
        threads = []
        t0 = time.perf_counter()
        for c in range(cores):
            shard = wavs[c::cores]
            t = threading.Thread(target=lambda s=shard: ref_run.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, s))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
    return {"value": audio_s / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d utterances (%.0f s audio) of the bench workload, one online2-wav-nnet3-latgen-faster per core, model load included" % (n, audio_s)}


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
    assert all(s == 0 for s in hyp.status), "decoder reported capacity problems"
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    stage_ms, wall_s, launches = [], [], 0
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (256 MiB > 126 MB L2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hyp = dec.decode_pcm(utts)
        wall_s.append(time.perf_counter() - t0)
        t = dec.timings()
        stage_ms.append((t["feature_ms"], t["nnet_ms"], t["decode_ms"], t["h2d_ms"], t["d2h_ms"], t["total_ms"]))
        launches += t["kernel_launches"]
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
            "vs_baseline": None, "dtype": "f32 (tensor-core products as 3 x fp16 split terms, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "configs[1]: batch=256 per GPU, grammar-HCLG, 3-5 s 16 kHz utterances",
                       "model": "zamia-like TDNN-F chain (synthetic, seeded): 40-dim hires MFCC + 100-dim iVector, 1024/128 x 12 TDNN-F, 3026 pdfs, sf=3",
                       "graph": "en_US grammar HCLG (synthetic lexicon), %d states" % graph.num_states,
                       "decoder": "beam 24, max-active 7000, lattice-beam 8 (best path)", "l2": "flushed between iterations (256 MiB write)",
                       "audio_seconds_per_step": audio_total},
            "e2e": {"value": audio_total / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": int(t["h2d_bytes"]),
                    "d2h_bytes_per_step": int(t["d2h_bytes"]), "ms_per_step": e2e_s_max * 1e3,
                    "input": "int16 PCM in one page-locked host block (rs_host_alloc), copied H2D inside the timed call"},
            "e2e_pageable": {"value": audio_total / pageable_s_max, "unit": UNIT, "ms_per_step": pageable_s_max * 1e3,
                             "input": "the same call from 256 ordinary numpy arrays: packed into pinned staging first"},
            "e2e_two_in_flight": {"value": audio_total / pipe_s_max, "unit": UNIT, "ms_per_batch": pipe_s_max * 1e3,
                                  "how": "two decoders per GPU driven by two host threads, same call, same per-batch copies"},
            "gpu_launches": int(launches),
            "stages_ms": {"feature": float(sm[:, 0].mean()), "nnet": nnet_ms, "decode": float(sm[:, 2].mean()),
                          "h2d": float(sm[:, 3].mean()), "d2h": float(sm[:, 4].mean())},
            # dominant kernel = gemm_tc_kernel (30 launches per step, the whole nnet stage between two CUDA
            # events on the decoder's stream).  achieved = algorithmic FLOPs (2*rows*K*N per layer, counted
            # once although every product is 3 MMAs => attainable frac <= 1/3) / stage time.
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (TDNN-F affine layers, 30 launches = the nnet stage)",
                         "achieved": achieved, "peak": f16_peak, "unit": "TFLOP/s", "frac": achieved / f16_peak,
                         "traffic": NNET_DRAM_BYTES_PER_STEP,
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the stage's launches, "
                                           "profiles/r1_launches_i_dram.csv (same command, batch 256)",
                         "peak_source": pk_kind + " bf16_tflops_sustained (fp16 tensor pipe)",
                         "note": "3 MMAs per algorithmic product: attainable frac <= 0.333"},
            # the same launches against HBM: the stage streams every layer's activations through HBM once
            "roofline_hbm": {"bound": "hbm", "kernel": "gemm_tc_kernel (same 30 launches)", "achieved": hbm_achieved,
                             "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_achieved / pk["hbm_gbs"],
                             "traffic": NNET_DRAM_BYTES_PER_STEP, "algorithmic_bytes": int(t["nnet_bytes"])},
            "decoder_counters": {k: int(t[k]) for k in ("frames_decoded", "tokens_expanded", "arcs_visited", "tokens_created")},
            "clocks": sampler.summary(),
            "wall_s_total": t_all,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline_sample()
            if cb:
                line["cpu_baseline"] = cb
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
