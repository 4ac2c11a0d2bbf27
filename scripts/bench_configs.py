"""Extra measurements for the BASELINE configs that are parity cases rather than the bench line:
config 3 (ARPA-shaped HCLG, batch 256, 10 % out-of-grammar audio) and config 4 (64 concurrent streams fed
80 ms chunks, online schedule).  Prints one JSON line per config with the stage times of rs_timings."""
import dataclasses, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rhasspy_speech_b200 import _lib
from tools import synth


def run(name, dec, fn, audio_s, reps=4):
    for _ in range(2):
        hyp = fn()
    ts, wall = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        hyp = fn()
        wall.append(time.perf_counter() - t0)
        ts.append(dec.timings())
    t = {k: float(np.mean([x[k] for x in ts])) for k in ("h2d_ms", "feature_ms", "nnet_ms", "decode_ms", "total_ms")}
    dev = t["feature_ms"] + t["nnet_ms"] + t["decode_ms"]
    flags = {int(s): int((np.asarray(hyp.status) == s).sum()) for s in set(int(x) for x in hyp.status)}
    out = {"config": name, "audio_s": audio_s, "rtfx_device": audio_s / (dev / 1e3), "rtfx_e2e": audio_s / float(np.mean(wall)),
           "wall_ms": float(np.mean(wall)) * 1e3, "stages_ms": t, "status_counts": flags,
           "strict_host_decoder": {"utterances": int(ts[-1]["strict_utts"]), "ms": float(np.mean([x["strict_ms"] for x in ts]))},
           "tokens_per_frame": ts[-1]["tokens_expanded"] / max(1, ts[-1]["frames_decoded"])}
    if ts[-1]["lattice_arcs"]:
        out["lattice"] = {k: ts[-1][k] for k in ("lattice_states", "lattice_arcs", "lattice_links_recorded", "d2h_bytes")}
        out["hyps_per_utt"] = float(np.mean(hyp.n_hyp))
    print(json.dumps(out), flush=True)


def main():
    tmp = tempfile.mkdtemp()
    if "--mixed" in sys.argv:
        # config 5, one GPU's share: 256 utterances of a mixed batch routed to 8 different (model, HCLG) pairs that are all
        # resident (the reference's 8 language models differ in exactly the ways the variants below do), 32 utterances each
        import threading
        variants = [
            synth.ZAMIA_LIKE,
            dataclasses.replace(synth.ZAMIA_LIKE, name="v1", seed=11),
            dataclasses.replace(synth.ZAMIA_LIKE, name="v2", seed=12, priors=True),
            dataclasses.replace(synth.ZAMIA_LIKE, name="v3", seed=13, lda_bias=True),
            dataclasses.replace(synth.ZAMIA_LIKE, name="v4", seed=14, nnet_cmvn=True),
            dataclasses.replace(synth.ZAMIA_LIKE, name="v5", seed=15, num_gauss=256, ivector_dim=60),
            dataclasses.replace(synth.ZAMIA_LIKE, name="v6", seed=16, graph="arpa", vocab_size=400, bigrams_per_word=10, eps_hops=2),
            dataclasses.replace(synth.ZAMIA_LIKE, name="v7", seed=17, binary=False),
        ]
        decs = []
        for i, spec in enumerate(variants):
            p = synth.write_model(os.path.join(tmp, "m%d" % i), spec)
            decs.append(_lib.Decoder(_lib.Model(p.final_mdl, p.online_conf, 0), _lib.Graph(p.hclg, p.words_txt, 0),
                                     max_tokens_per_utt=1 << 20))
        utts = synth.make_utterances(256, seed=1234, pool=synth.load_pool())
        parts = [utts[i::len(decs)] for i in range(len(decs))]
        audio_s = sum(len(u) for u in utts) / 16000.0

        def sequential():
            return [d.decode_pcm(pt) for d, pt in zip(decs, parts)]

        def threaded():
            out = [None] * len(decs)
            th = [threading.Thread(target=lambda i=i: out.__setitem__(i, decs[i].decode_pcm(parts[i]))) for i in range(len(decs))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            return out
        res = {}
        for name, fn in (("sequential", sequential), ("8 host threads", threaded)):
            for _ in range(2):
                hyps = fn()
            walls = []
            for _ in range(4):
                t0 = time.perf_counter()
                hyps = fn()
                walls.append(time.perf_counter() - t0)
            res[name] = float(np.mean(walls)) * 1e3
            assert all(h.n_utts == len(pt) for h, pt in zip(hyps, parts))
        dev = sum(sum(d.timings()[k] for k in ("feature_ms", "nnet_ms", "decode_ms")) for d in decs)
        print(json.dumps({"config": "5 (one GPU's share): 256 utterances over 8 resident (model, HCLG) pairs, 32 each", "audio_s": audio_s,
                          "wall_ms": res, "rtfx_e2e": {k: audio_s / (v / 1e3) for k, v in res.items()},
                          "device_ms_sum_of_8_batches": dev, "decoded": int(sum(int((np.asarray(h.n_hyp) > 0).sum()) for h in hyps))}), flush=True)
        return
    if "--surface" in sys.argv:
        # config 4 through the Python mirror: 64 concurrent KaldiNnet3StreamTranscriber.async_transcribe coroutines fed
        # 80 ms chunks; the host dynamic batcher turns them into a few device batches
        import asyncio
        import rhasspy_speech_b200 as pkg
        p = synth.write_model(os.path.join(tmp, "gram"), synth.ZAMIA_LIKE)
        st = pkg.KaldiNnet3StreamTranscriber(p.model_dir, os.path.dirname(p.hclg), None)
        utts64 = synth.make_utterances(64, seed=4321, pool=synth.load_pool())
        raws = [np.asarray(u, dtype="<i2").tobytes() for u in utts64]

        async def chunks(raw):
            for o in range(0, len(raw), 2560):
                yield raw[o:o + 2560]
                await asyncio.sleep(0)

        async def burst():
            return await asyncio.gather(*[st.async_transcribe(chunks(r), tmp) for r in raws])
        asyncio.run(burst())
        eng = st._get_engine()
        walls, sizes = [], []
        for _ in range(4):
            n0 = len(eng.batcher.batches)
            t0 = time.perf_counter()
            out = asyncio.run(burst())
            walls.append(time.perf_counter() - t0)
            sizes.append(eng.batcher.batches[n0:])
        audio_s = sum(len(u) for u in utts64) / 16000.0
        print(json.dumps({"config": "4 (Python surface): 64 concurrent async streams x 80 ms chunks", "audio_s": audio_s,
                          "wall_ms": float(np.mean(walls)) * 1e3, "rtfx_e2e": audio_s / float(np.mean(walls)),
                          "device_batches": sizes[-1], "decoded": sum(1 for o in out if o)}), flush=True)
        return
    if "--nbest" in sys.argv:
        # the n-best tail on the bench workload (configs[1]): batch 256, grammar graph, n = 1 (device back-trace) vs
        # n = 5 (lattice recorded + pruned on the device, best-first search on the host); decode_ms covers
        # decode_kernel (+ lattice_prune_kernel), wall - total covers the host search
        p = synth.write_model(os.path.join(tmp, "gram"), synth.ZAMIA_LIKE)
        dec = _lib.Decoder(_lib.Model(p.final_mdl, p.online_conf, 0), _lib.Graph(p.hclg, p.words_txt, 0))
        utts = synth.make_utterances(256, seed=1234, pool=synth.load_pool())
        audio_s = sum(len(u) for u in utts) / 16000.0
        for n in (1, 5):
            dec.set_nbest(n)
            run("2: grammar HCLG, batch 256, nbest=%d" % n, dec, lambda: dec.decode_pcm(utts), audio_s)
        return
    # config 3: zamia-like model, ARPA-shaped graph
    spec = dataclasses.replace(synth.ZAMIA_LIKE, name="zamia_arpa", graph="arpa", vocab_size=2000, bigrams_per_word=20, eps_hops=2)
    p = synth.write_model(os.path.join(tmp, "arpa"), spec)
    utts = synth.make_utterances(256, seed=1234, pool=synth.load_pool())
    for i in range(0, 256, 10):
        utts[i] = utts[i][::-1].copy()          # out-of-grammar audio: time-reversed
    m = _lib.Model(p.final_mdl, p.online_conf, 0)
    g = _lib.Graph(p.hclg, p.words_txt, 0)
    audio_s = sum(len(u) for u in utts) / 16000.0
    # the device search alone (order-sensitive utterances only flagged), then with the strict-order host decoder on
    dec0 = _lib.Decoder(m, g, strict_fallback=0, max_tokens_per_utt=1 << 21)
    run("3 (device search only, flags not resolved): ARPA-shaped HCLG (%d states, %d arcs), batch 256, 10%% reversed audio"
        % (g.num_states, g.num_arcs), dec0, lambda: dec0.decode_pcm(utts), audio_s, reps=3)
    del dec0
    dec = _lib.Decoder(m, g, max_tokens_per_utt=1 << 21)
    run("3: ARPA-shaped HCLG (%d states, %d arcs), batch 256, 10%% reversed audio; order-sensitive utterances re-decoded by the "
        "strict-order host decoder" % (g.num_states, g.num_arcs), dec, lambda: dec.decode_pcm(utts), audio_s, reps=3)
    # config 4: 64 streams, 80 ms chunks, grammar graph, online schedule
    p2 = synth.write_model(os.path.join(tmp, "gram"), synth.ZAMIA_LIKE)
    m2 = _lib.Model(p2.final_mdl, p2.online_conf, 0)
    g2 = _lib.Graph(p2.hclg, p2.words_txt, 0)
    dec2 = _lib.Decoder(m2, g2)
    utts64 = synth.make_utterances(64, seed=4321, pool=synth.load_pool())
    raws = [np.asarray(u, dtype="<i2").tobytes() for u in utts64]
    streams = [dec2.open_stream() for _ in utts64]

    def stream_pass():
        for s, raw in zip(streams, raws):
            for o in range(0, len(raw), 2560):
                s.accept(raw[o:o + 2560])
        return dec2.finish_streams(streams)
    run("4: 64 streams x 80 ms chunks, online iVector schedule, grammar HCLG", dec2, stream_pass,
        sum(len(u) for u in utts64) / 16000.0)
    # latency from end of input to the result: all 64 streams ending in the same tick (one device batch),
    # and streams ending one at a time (each its own batch of one); accept() only buffers, so this is the whole
    # feature + nnet + search time of the utterance(s)
    lat_all, lat_one = [], []
    for _ in range(3):
        for s, raw in zip(streams, raws):
            for o in range(0, len(raw), 2560):
                s.accept(raw[o:o + 2560])
        t0 = time.perf_counter()
        dec2.finish_streams(streams)
        lat_all.append((time.perf_counter() - t0) * 1e3)
    for rep in range(2):
        for s, raw in zip(streams, raws):
            for o in range(0, len(raw), 2560):
                s.accept(raw[o:o + 2560])
        for s in streams:
            t0 = time.perf_counter()
            dec2.finish_streams([s])
            if rep:
                lat_one.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"config": "4: latency from end of input to result", "all_64_end_together_ms": float(np.mean(lat_all)),
                      "one_stream_ends_alone_ms": {"p50": float(np.percentile(lat_one, 50)), "p99": float(np.percentile(lat_one, 99)),
                                                   "max": float(np.max(lat_one))},
                      "utterance_seconds": {"min": min(len(u) for u in utts64) / 16000.0, "max": max(len(u) for u in utts64) / 16000.0}}), flush=True)


if __name__ == "__main__":
    main()
