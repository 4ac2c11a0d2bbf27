"""The n-best tail on the GPU against the reference: lattice recorded by
decode_kernel<true>, lattice-beam pruning on the device, n cheapest distinct word sequences.

Oracle: latgen-faster-mapped / online2-wav-nnet3-latgen-faster (determinised lattice) ->
lattice-to-nbest --n --acoustic-scale -> nbest-to-linear, from oracle/_ref.  Word sequences and their order
must be identical; the costs are compared to the 7 digits the archives print; the size of the pruned
state-level lattice must equal the reference's raw lattice (--determinize-lattice=false) state for state."""
import os

import numpy as np
import pytest

from test_decoder_oracle import arc_multiset

pytestmark = pytest.mark.gpu


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
def tiny(lib, tiny_model):
    m = lib.Model(tiny_model.final_mdl, tiny_model.online_conf, 0)
    g = lib.Graph(tiny_model.hclg, tiny_model.words_txt, 0)
    return m, g, lib.Decoder(m, g)


def _close(a, b):
    return abs(a - b) <= 3e-4 * max(1.0, abs(b))


def _compare(got, want_nb, n_utts, tag):
    n_multi = 0
    for u in range(n_utts):
        key = "utt%05d-" % u
        want = [want_nb[k] for k in sorted(want_nb, key=lambda k: int(k.rsplit("-", 1)[1])) if k.startswith(key)]
        have = got.nbest[u]
        assert len(have) == len(want), (tag, u, have, want)
        n_multi += len(want) > 1
        for h, ((w, g, a), (ww, wg, wa)) in enumerate(zip(have, want)):
            assert w == ww, (tag, u, h, w, ww)
            assert _close(g, wg) and _close(a, wa), (tag, u, h, g, wg, a, wa)
        if have:        # the head of the list is the device back-trace
            assert got.words[u] == have[0][0]
    return n_multi


def test_nbest_and_lattice_match_reference_on_loglikes(tiny, tiny_model, utterances, ref):
    _, _, dec = tiny
    dec.set_nbest(1)
    dec.decode_pcm(utterances)
    lls = [dec.fetch(2, u) for u in range(len(utterances))]
    n_multi = 0
    try:
        for ll_scale, nb_scale in ((1.0, 1.0), (0.3, 1.0), (0.3, 0.5), (3.0, 1.0)):
            mats = [np.ascontiguousarray(l * np.float32(ll_scale)) for l in lls]
            raw, want = ref.decode_loglikes_lattice(tiny_model.final_mdl, tiny_model.hclg, mats, nbest=5, acoustic_scale=nb_scale)
            dec.set_nbest(5, nb_scale)
            got = dec.decode_loglikes(mats)
            assert all(int(s) & 15 == 0 for s in got.status), list(got.status)
            n_multi += _compare(got, want, len(mats), (ll_scale, nb_scale))
            # a20: the pruned lattice is the reference's raw lattice, state for state and arc for arc
            t = dec.timings()
            assert t["lattice_states"] == sum(raw[k]["n_states"] for k in raw), (t["lattice_states"], [raw[k]["n_states"] for k in raw])
            assert t["lattice_arcs"] == sum(len(raw[k]["src"]) for k in raw)
            assert t["lattice_links_recorded"] >= t["lattice_arcs"] - sum(int((raw[k]["dst"] < 0).sum()) for k in raw)
            for u in range(len(mats)):
                A = dec.fetch(5, u)
                mine = dict(dst=A[:, 1], olabel=A[:, 2], graph=A[:, 3], acoustic=A[:, 4])
                diff = arc_multiset(mine) - arc_multiset(raw["utt%05d" % u])
                assert sum(diff.values()) <= 2, (ll_scale, u, diff)          # rounding of the 2-digit key only
    finally:
        dec.set_nbest(1)
    assert n_multi >= 8
    # back on the single-best path: same words as the head of the lists
    one = dec.decode_loglikes(mats)
    assert [w for w in one.words] == [got.nbest[u][0][0] for u in range(len(mats))]
    assert all(len(x) == 1 for x in one.nbest)


def test_nbest_transcripts_match_reference_pipeline(tiny, tiny_model, utterances, ref, synth, tmp_path):
    """Whole pipeline: WAVs -> `utt-1 .. utt-n` of online2-wav-nnet3-latgen-faster | lattice-to-nbest --n=5 | nbest-to-linear."""
    from rhasspy_speech_b200 import transcribe as T
    _, _, dec = tiny
    wavs = []
    for i, pcm in enumerate(utterances):
        p = os.path.join(str(tmp_path), "u%03d.wav" % i)
        synth.write_wav(p, pcm)
        wavs.append(p)
    want, _, _ = ref.transcribe_wavs(tiny_model.final_mdl, tiny_model.online_conf, tiny_model.hclg, tiny_model.words_txt, wavs, nbest=5)
    try:
        dec.set_nbest(5)
        got = dec.decode_wavs(wavs + [wavs[0]])       # ragged batch, one utterance twice
    finally:
        dec.set_nbest(1)
    for u in range(len(wavs)):
        keys = sorted((k for k in want if k.startswith("utt%05d-" % u)), key=lambda k: int(k.rsplit("-", 1)[1]))
        assert [h[0] for h in got.nbest[u]] == [want[k] for k in keys], (u, got.nbest[u], [want[k] for k in keys])
        # the bytes the Python tail parses (transcribe_wav.py:99-103)
        text = T.nbest_text(got, u).decode().splitlines()
        assert text == ["utt-%d %s" % (i + 1, "".join("%d " % w for w in want[k])) for i, k in enumerate(keys)]
    assert got.nbest[len(wavs)] == got.nbest[0]
    # an utterance too short to decode has no hypotheses in either mode
    try:
        dec.set_nbest(3)
        e = dec.decode_pcm([np.zeros(100, np.int16), utterances[1]])
    finally:
        dec.set_nbest(1)
    assert e.nbest[0] == [] and e.n_hyp[0] == 0 and len(e.nbest[1]) >= 1


def test_nbest_stream_and_python_surface(tiny_model, utterances, ref, synth, tmp_path):
    """nbest through the mirrors of KaldiNnet3WavTranscriber / KaldiNnet3StreamTranscriber (transcribe_wav.py:35-105,
    transcribe_stream.py:38-129): the strings of every hypothesis, in the reference's order."""
    import asyncio
    import rhasspy_speech_b200 as pkg
    model_dir = tiny_model.model_dir
    graph_dir = os.path.dirname(tiny_model.hclg)
    wav = os.path.join(str(tmp_path), "a.wav")
    pcm = utterances[3]
    synth.write_wav(wav, pcm)
    words = {}
    with open(tiny_model.words_txt) as f:
        for line in f:
            w, i = line.split()
            words[int(i)] = w
    want, _, _ = ref.transcribe_wavs(tiny_model.final_mdl, tiny_model.online_conf, tiny_model.hclg, tiny_model.words_txt, [wav], nbest=4)
    keys = sorted(want, key=lambda k: int(k.rsplit("-", 1)[1]))
    want_text = [" ".join(words[i] for i in want[k]) for k in keys if want[k]]
    tr = pkg.KaldiNnet3WavTranscriber(model_dir, graph_dir, None)
    got = asyncio.run(tr.async_transcribe(wav, tmp_path, nbest=4))
    assert got == want_text and len(got) > 1, (got, want_text)
    assert asyncio.run(tr.async_transcribe(wav, tmp_path)) == want_text[:1]
    # stream surface against the stream binary's lattice
    swant, _ = ref.transcribe_stream(tiny_model.final_mdl, tiny_model.online_conf, tiny_model.hclg, tiny_model.words_txt, pcm, nbest=4)
    skeys = sorted(swant, key=lambda k: int(k.rsplit("-", 1)[1]))
    swant_text = [" ".join(words[i] for i in swant[k]) for k in skeys if swant[k]]

    async def chunks():
        raw = np.asarray(pcm, dtype="<i2").tobytes()
        for o in range(0, len(raw), 2560):
            yield raw[o:o + 2560]
    st = pkg.KaldiNnet3StreamTranscriber(model_dir, graph_dir, None)
    assert asyncio.run(st.async_transcribe(chunks(), tmp_path, nbest=4)) == swant_text


def test_nbest_on_arpa_graph_with_epsilon_chains(lib, ref, synth, utterances, tmp_path):
    """BASELINE config 3 in miniature: ARPA-shaped HCLG (back-off epsilon chains, global-memory state tables), plus
    out-of-grammar audio; the lattices are an order of magnitude larger than on the grammar graph."""
    import dataclasses
    spec = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
    p = synth.write_model(str(tmp_path / "m"), spec)
    utts = list(utterances[:4]) + [utterances[0][::-1].copy()]
    wavs = []
    for i, pcm in enumerate(utts):
        w = os.path.join(str(tmp_path), "a%03d.wav" % i)
        synth.write_wav(w, pcm)
        wavs.append(w)
    m, g = lib.Model(p.final_mdl, p.online_conf, 0), lib.Graph(p.hclg, p.words_txt, 0)
    assert g.num_states > 1024
    n_lists = n_flagged = 0
    # Every list must be the reference's: beam 16 keeps fewer than --max-active tokens per frame; at beam 24 the
    # order-sensitive utterances (status bit 4: tokens or links the reference's transient next_cutoff lets through,
    # :780-787, could matter) come from the strict-order host decoder, lattice included.
    for beam in (24.0, 16.0):
        dec = lib.Decoder(m, g, beam=beam)
        want, _, _ = ref.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, wavs, nbest=5, beam=beam)
        dec.set_nbest(5)
        got = dec.decode_wavs(wavs)
        t = dec.timings()
        assert t["lattice_arcs"] > 0 and t["lattice_links_recorded"] > 0
        for u in range(len(wavs)):
            assert int(got.status[u]) & 15 == 0, (beam, u, got.status[u])
            keys = sorted((k for k in want if k.startswith("utt%05d-" % u)), key=lambda k: int(k.rsplit("-", 1)[1]))
            assert [h[0] for h in got.nbest[u]] == [want[k] for k in keys], (beam, u, got.status[u], got.nbest[u], [want[k] for k in keys])
            n_lists += len(keys) > 1
            n_flagged += bool(got.status[u] & 16)
    print("arpa n-best: %d lists identical, %d of them from the strict-order host decoder" % (n_lists, n_flagged))
    assert n_lists >= 2


def test_lattice_that_does_not_fit_falls_back_to_the_best_path(tiny, utterances, monkeypatch):
    """An utterance whose lattice outgrows its slice of the device budget still returns its best path (status bit 5);
    the other utterances of the call are unaffected."""
    _, _, dec = tiny
    batch = [utterances[i % len(utterances)] for i in range(64)]
    dec.set_nbest(1)
    one = dec.decode_pcm(batch)
    from rhasspy_speech_b200 import _lib
    model, graph, _ = tiny
    small = _lib.Decoder(model, graph)                    # its own decoder: the lattice budget only ever grows
    monkeypatch.setenv("RS_B200_LATTICE_MB", "16")        # 16 MB / 64 utterances: 4096 tokens each
    monkeypatch.setenv("RS_B200_LATTICE_MAX_MB", "16")    # ... and no growth
    small.set_nbest(4)
    got = small.decode_pcm(batch)
    assert all(s & 32 for s in got.status) and all((s & ~(48 | 64)) == 0 for s in got.status), list(got.status[:8])
    assert list(got.n_hyp) == [1] * 64
    assert got.words == one.words
    # with room to grow, the same call re-runs the stage with a larger budget and returns the lists
    monkeypatch.delenv("RS_B200_LATTICE_MAX_MB")
    full = small.decode_pcm(batch)
    assert all(int(s) & 15 == 0 for s in full.status) and max(full.n_hyp) > 1
    assert [h[0][0] for h in full.nbest] == one.words
    dec.set_nbest(4)
    try:
        ref_lists = dec.decode_pcm(batch[:6])
    finally:
        dec.set_nbest(1)
    assert [x for x in full.nbest[:6]] == [x for x in ref_lists.nbest]
