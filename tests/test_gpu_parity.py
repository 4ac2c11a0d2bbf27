"""Parity of the CUDA path (through the C ABI) against the oracle, stage by stage and end to end.

The oracle is (a) the reference binaries in oracle/_ref (built from /root/reference by
oracle/build_ref.py; they travel to the GPU box) and (b) the numpy restatement oracle/kaldi_np.py.
Tolerances are stated per test; word sequences and frame counts must be identical.
"""
import os

import numpy as np
import pytest

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


def _write_wavs(synth, tmp, utts):
    paths = []
    for i, pcm in enumerate(utts):
        p = os.path.join(str(tmp), "u%03d.wav" % i)
        synth.write_wav(p, pcm)
        paths.append(p)
    return paths


def test_mfcc_matches_reference(tiny, tiny_model, utterances, ref, synth, tmp_path):
    _, _, dec = tiny
    hyp = dec.decode_pcm(utterances)
    conf = os.path.join(tiny_model.model_dir, "model", "online", "conf", "mfcc.conf")
    feats = ref.mfcc(conf, _write_wavs(synth, tmp_path, utterances))
    for u, f in enumerate(feats):
        got = dec.fetch(0, u)
        assert got.shape == f.shape
        # window + FFT are bit-exact by construction; the mel/DCT dot products differ in summation
        # order from OpenBLAS: a few ulp of C0 (~1e2)
        assert np.abs(got - f).max() <= 2e-4, np.abs(got - f).max()
    assert hyp.n_utts == len(utterances)


def test_ivector_matches_oracle(tiny, tiny_model, utterances):
    from oracle import kaldi_np as K
    _, _, dec = tiny
    dec.decode_pcm(utterances)
    conf = os.path.join(tiny_model.model_dir, "model", "online", "conf")
    s = K.IvectorSetup.from_conf(os.path.join(conf, "ivector_extractor.conf"))
    for u in range(len(utterances)):
        mf = dec.fetch(0, u)
        want = K.ivector_offline(s, mf)
        got = dec.fetch(1, u)[0]
        assert np.abs(got - want).max() <= 1e-4, np.abs(got - want).max()
        # CMVN'd features and LDA features, against the restatement on the same MFCCs
        norm = K.online_cmvn(mf, s.global_cmvn, s.cmvn)
        assert np.abs(dec.fetch(3, u) - norm).max() <= 1e-5
        lda = K.lda_feats(s, mf, True)
        assert np.abs(dec.fetch(4, u) - lda).max() <= 2e-4


def test_ubm_posteriors_match_oracle(tiny, tiny_model, utterances):
    """Rows a9 / a10 directly: DiagGmm::LogLikelihoods + VectorToPosteriorEntry (top num_gselect above min_post,
    renormalised, x posterior_scale) on the device against the restatement (itself pinned to ivector-extract-online2):
    the same Gaussians in the same order, weights to 1e-5."""
    from oracle import kaldi_np as K
    _, _, dec = tiny
    dec.decode_pcm(utterances)
    conf = os.path.join(tiny_model.model_dir, "model", "online", "conf")
    s = K.IvectorSetup.from_conf(os.path.join(conf, "ivector_extractor.conf"))
    n_frames = n_pruned = 0
    for u in range(len(utterances)):
        xn = dec.fetch(4, u)                       # the device's own LDA features (normalised stream)
        want = K.gmm_posteriors(s, xn)
        got = dec.fetch(6, u)
        assert got.shape == (len(want), 2 * s.num_gselect)
        for t, row in enumerate(want):
            idx = [int(x) for x in got[t, 0::2] if x >= 0]
            # two Gaussians whose posteriors differ in the last bit may swap places; compare as sets + weights by index
            assert sorted(idx) == sorted(g for g, _ in row), (u, t, idx, row)
            w = {int(g): float(x) for g, x in zip(got[t, 0::2], got[t, 1::2]) if g >= 0}
            for g, x in row:
                assert abs(w[g] - float(x)) <= 1e-5, (u, t, g, w[g], x)
            n_pruned += len(row) < s.num_gselect
        n_frames += len(want)
    assert n_frames > 500 and n_pruned > 0          # the min_post pruning was exercised


def test_loglikes_match_reference(tiny, tiny_model, utterances, ref, synth):
    _, _, dec = tiny
    dec.decode_pcm(utterances)
    feats = [dec.fetch(0, u) for u in range(len(utterances))]
    ivs = [dec.fetch(1, u)[0] for u in range(len(utterances))]
    want = ref.nnet_loglikes(tiny_model.final_mdl, feats, ivs, frame_subsampling_factor=3)
    for u, w in enumerate(want):
        got = dec.fetch(2, u)
        assert got.shape == w.shape, (got.shape, w.shape)
        assert np.abs(got - w).max() <= 1e-4, np.abs(got - w).max()   # north-star tolerance
        f64 = synth.nnet_forward(tiny_model.nnet_params, feats[u].astype(np.float64), ivs[u].astype(np.float64))[::3]
        assert np.abs(got - f64).max() <= 1e-4


def test_decoder_matches_reference_on_loglikes(tiny, tiny_model, utterances, ref):
    _, _, dec = tiny
    dec.decode_pcm(utterances)
    lls = [dec.fetch(2, u) for u in range(len(utterances))]
    # sharper and flatter score distributions move the beam pruning
    for scale in (1.0, 0.3, 3.0):
        mats = [np.ascontiguousarray(l * np.float32(scale)) for l in lls]
        want = ref.decode_loglikes(tiny_model.final_mdl, tiny_model.hclg, mats)
        got = dec.decode_loglikes(mats)
        for u in range(len(mats)):
            assert got.words[u] == want.get("utt%05d-1" % u), (scale, u, got.words[u], want.get("utt%05d-1" % u))


def test_transcripts_match_reference(tiny, tiny_model, utterances, ref, synth, tmp_path):
    _, graph, dec = tiny
    wavs = _write_wavs(synth, tmp_path, utterances)
    want, _, _ = ref.transcribe_wavs(tiny_model.final_mdl, tiny_model.online_conf, tiny_model.hclg, tiny_model.words_txt, wavs)
    got = dec.decode_wavs(wavs)
    assert list(got.status) == [0] * len(wavs)
    for u in range(len(wavs)):
        assert got.words[u] == want.get("utt%05d-1" % u), (u, got.words[u], want.get("utt%05d-1" % u))


def test_edge_cases(tiny, utterances):
    """Empty, too-short (< one window), exactly one window, and ragged batches."""
    _, _, dec = tiny
    one = utterances[0]
    batch = [np.zeros(0, np.int16), one[:399], one[:400], one[:401], one, one[:8000]]
    hyp = dec.decode_pcm(batch)
    assert list(hyp.n_hyp[:2]) == [0, 0] and hyp.words[0] is None and hyp.words[1] is None
    assert list(hyp.num_frames) == [0, 0, 1, 1, (1 + (len(one) - 400) // 160 + 2) // 3, (1 + (8000 - 400) // 160 + 2) // 3]
    # batching must not change a result: decode the same utterances one by one
    for u in (2, 4, 5):
        single = dec.decode_pcm([batch[u]])
        assert single.words[0] == hyp.words[u]
    assert dec.decode_pcm([]).n_utts == 0


@pytest.mark.parametrize("k,n,offsets,stride", [(40, 220, (-1, 0, 1), 1), (1024, 128, (-3, 0), 1), (128, 1024, (0, 1), 3),
                                                (192, 3026, (0,), 1), (36, 40, (-2, -1, 0, 1), 2)])
def test_tensor_core_layer_matches_fp64(lib, k, n, offsets, stride):
    """The tcgen05 3xTF32 layer kernel against an fp64 product: error at the fp32 level, far inside 1e-4."""
    rng = np.random.default_rng(k * 131 + n)
    rows = 700
    src = rng.standard_normal((rows, k)).astype(np.float32)
    w = (rng.standard_normal((n, k * len(offsets))) / np.sqrt(k * len(offsets))).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    m = rows // stride
    r = np.arange(m) * stride
    want = np.zeros((m, n))
    valid = np.ones(m, dtype=bool)
    for i, o in enumerate(offsets):
        idx = r + o
        valid &= (idx >= 0) & (idx < rows)
        want += src[np.clip(idx, 0, rows - 1)].astype(np.float64) @ w[:, i * k:(i + 1) * k].astype(np.float64).T
    want = np.maximum(want + bias, 0)
    for path in (0, 1, 2):
        got, _ = lib.debug_gemm(src, w, offsets, stride, bias, True, path=path)
        err = np.abs(got - want)[valid].max()
        assert err <= 2e-5, (path, err)


def _check_transcripts(lib, ref, synth, p, utts, tmp, **dec_opts):
    """Whole pipeline through the C ABI vs the reference's 3-process pipeline on the same WAVs."""
    m = lib.Model(p.final_mdl, p.online_conf, 0)
    g = lib.Graph(p.hclg, p.words_txt, 0)
    dec = lib.Decoder(m, g, **dec_opts)
    wavs = _write_wavs(synth, tmp, utts)
    ref_kw = {k: v for k, v in dec_opts.items() if k in ("beam", "max_active")}
    want, _, _ = ref.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, wavs, **ref_kw)
    got = dec.decode_wavs(wavs)
    assert all(int(st) & 15 == 0 for st in got.status), list(got.status)   # bits 4 / 6: order-sensitive, decoded by the host
    n_words = 0
    for u in range(len(wavs)):
        w = want.get("utt%05d-1" % u)
        assert got.words[u] == w, (u, got.words[u], w)
        n_words += len(w or [])
    return dec, got, n_words


def test_arpa_graph_with_out_of_grammar_audio(lib, ref, synth, utterances, tmp_path):
    """BASELINE config 3: ARPA-LM-shaped HCLG (back-off epsilon arcs, tens of thousands of arcs) and
    out-of-grammar audio (time-reversed utterances); also with a tight --max-active so that
    the max-active cutoff (GetCutoff's nth_element value) decides the beam on every frame."""
    import dataclasses
    spec = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
    p = synth.write_model(str(tmp_path / "m"), spec)
    utts = list(utterances) + [u[::-1].copy() for u in utterances[:3]]
    dec, got, n_words = _check_transcripts(lib, ref, synth, p, utts, tmp_path)
    assert n_words > 0 and dec.graph.num_states > 1024          # the global-memory table path
    t = dec.timings()
    assert t["tokens_expanded"] > 20 * t["frames_decoded"]
    _check_transcripts(lib, ref, synth, p, utts, tmp_path, beam=12.0)
    # fewer table slots than graph states: the open-addressing (hashed) path of the state tables
    dec_h, got_h, _ = _check_transcripts(lib, ref, synth, p, utts, tmp_path, beam=9.0, max_tokens_per_frame=2048)
    assert dec_h.graph.num_states > 2 * 2048 and all(int(st) & 15 == 0 for st in got_h.status)
    # A binding --max-active: the reference's token list then also holds the order-dependent tokens its transient
    # next_cutoff let through (lattice-faster-decoder.cc:780-787).  The device search detects the frames on which they
    # could matter (status bit 4) and such utterances are decoded again by the strict-order host decoder
    # (strict_decode.cc): EVERY utterance must carry the reference's words, at every --max-active.
    m = lib.Model(p.final_mdl, p.online_conf, 0)
    g = lib.Graph(p.hclg, p.words_txt, 0)
    wavs = _write_wavs(synth, tmp_path, utts)
    flagged = 0
    for max_active in (300, 1000, 7000):
        want, _, _ = ref.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, wavs, max_active=max_active)
        dec_m = lib.Decoder(m, g, max_active=max_active)
        got = dec_m.decode_wavs(wavs)
        for u in range(len(wavs)):
            assert int(got.status[u]) & 15 == 0 and got.n_hyp[u] == 1, (max_active, u, got.status[u])
            assert got.words[u] == want.get("utt%05d-1" % u), (max_active, u, got.status[u], got.words[u], want.get("utt%05d-1" % u))
            flagged += bool(got.status[u] & 16)
        assert dec_m.timings()["strict_utts"] == sum(1 for s_ in got.status if s_ & 64)
        # with the host decoder switched off the flag is still raised, and unflagged utterances are still identical
        raw = lib.Decoder(m, g, max_active=max_active, strict_fallback=0).decode_wavs(wavs)
        for u in range(len(wavs)):
            assert not (raw.status[u] & 64)
            if not (raw.status[u] & 16):
                assert raw.words[u] == want.get("utt%05d-1" % u), (max_active, u)
        # 5-best lists under the same limits (lattice recorded on the device, or by the strict decoder when flagged)
        dec_m.set_nbest(5)
        got5 = dec_m.decode_wavs(wavs)
        want5, _, _ = ref.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, wavs, nbest=5, max_active=max_active)
        for u in range(len(wavs)):
            w5 = [want5[k] for k in sorted(want5) if k.startswith("utt%05d-" % u)]
            assert [h[0] for h in got5.nbest[u]] == w5, (max_active, u, got5.status[u])
    print("order-sensitive utterances over max-active 300/1000/7000:", flagged)


@pytest.mark.parametrize("variant", ["softmax_sf1", "text_priors_ldabias", "nnet_cmvn"])
def test_model_variants_match_reference(lib, ref, synth, utterances, tmp_path, variant):
    """BASELINE config 5 in miniature: differently shaped models side by side in one process (the
    reference's 8 language models differ in exactly these respects): frame-subsampling 1 with a
    log-softmax output and priors, text-format files with an LDA offset column, CMVN on the nnet input."""
    import dataclasses
    if variant == "softmax_sf1":
        spec = dataclasses.replace(synth.TINY, name=variant, seed=21, chain=False, frame_subsampling_factor=1, log_softmax=True,
                                   priors=True, tdnnf_strides=(1, 0, 1, 1))
    elif variant == "text_priors_ldabias":
        spec = dataclasses.replace(synth.TINY, name=variant, seed=22, binary=False, priors=True, lda_bias=True)
    else:
        spec = dataclasses.replace(synth.TINY, name=variant, seed=23, nnet_cmvn=True, num_gauss=64, ivector_dim=40)
    p = synth.write_model(str(tmp_path / "m"), spec)
    dec, got, n_words = _check_transcripts(lib, ref, synth, p, utterances, tmp_path)
    assert n_words > 0
    # log-likelihoods of this variant against nnet3-compute (1e-4, the north-star tolerance)
    feats = [dec.fetch(3 if spec.nnet_cmvn else 0, u) for u in range(len(utterances))]
    ivs = [dec.fetch(1, u)[0] for u in range(len(utterances))]
    want = ref.nnet_loglikes(p.final_mdl, feats, ivs, frame_subsampling_factor=spec.frame_subsampling_factor)
    for u, w in enumerate(want):
        ll = dec.fetch(2, u)
        assert ll.shape == w.shape and np.abs(ll - w).max() <= 1e-4, (u, ll.shape, w.shape, np.abs(ll - w).max())


def test_stream_surface_matches_reference_stream_binary(tiny, tiny_model, utterances, ref):
    """rs_stream_*: 80 ms chunks (BASELINE config 4 framing), many streams finished in one device batch.
    The stream surface reproduces the ONLINE schedule of online2-cli-nnet3-decode-faster (1024-sample reads,
    one warm-started iVector per nnet chunk): iVectors of every solve against the restatement, words against
    the reference binary fed the same bytes."""
    from oracle import kaldi_np as K
    model, _, dec = tiny
    assert dec.finish_streams([]).n_utts == 0
    streams = [dec.open_stream() for _ in utterances]
    for s, pcm in zip(streams, utterances):
        raw = np.asarray(pcm, dtype="<i2").tobytes()
        for o in range(0, len(raw), 2560):
            s.accept(raw[o:o + 2560])
    got = dec.finish_streams(streams)
    conf = os.path.join(tiny_model.model_dir, "model", "online", "conf")
    setup = K.IvectorSetup.from_conf(os.path.join(conf, "ivector_extractor.conf"))
    n_multi = 0
    for u, pcm in enumerate(utterances):
        mf = dec.fetch(0, u)
        solves, _ = K.online_schedule(len(pcm), mf.shape[0], 24, model.right_context, setup.splice_right, 3)
        want_iv = K.ivectors_online(setup, mf, solves)
        iv = dec.fetch(1, u)
        assert iv.shape == want_iv.shape, (iv.shape, want_iv.shape, solves)
        assert np.abs(iv - want_iv).max() <= 1e-4, np.abs(iv - want_iv).max()
        n_multi += len(solves) > 1
        want, _ = ref.transcribe_stream(tiny_model.final_mdl, tiny_model.online_conf, tiny_model.hclg, tiny_model.words_txt, pcm)
        assert got.words[u] == want.get("utt-1"), (u, got.words[u], want.get("utt-1"))
    assert n_multi == len(utterances)
    # the WAV path keeps the offline schedule: one solve per utterance
    dec.decode_pcm(utterances[:1])
    assert dec.fetch(1, 0).shape[0] == 1
    # a stream object is reusable after finish
    streams[0].accept(np.asarray(utterances[0], dtype="<i2").tobytes())
    assert streams[0].finish().words[0] == got.words[0]


@pytest.mark.parametrize("k,n,offsets,relu", [(64, 128, (0,), False), (100, 200, (-1, 0), True), (256, 96, (0, 3), True)])
def test_tensor_core_layer_many_tiles_per_cta(lib, k, n, offsets, relu):
    """More tiles than SMs: every persistent CTA walks several tiles, so both epilogue groups, the TMEM
    accumulator ring and the shared-memory stage ring wrap around many times (odd and even K-block counts)."""
    rng = np.random.default_rng(k + n)
    rows = 128 * 148 * 3 + 77
    src = rng.standard_normal((rows, k)).astype(np.float32)
    w = (rng.standard_normal((n, k * len(offsets))) / np.sqrt(k * len(offsets))).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    r = np.arange(rows)
    want = np.zeros((rows, n))
    valid = np.ones(rows, dtype=bool)
    for i, o in enumerate(offsets):
        idx = r + o
        valid &= (idx >= 0) & (idx < rows)
        want += src[np.clip(idx, 0, rows - 1)].astype(np.float64) @ w[:, i * k:(i + 1) * k].astype(np.float64).T
    want += bias
    if relu:
        want = np.maximum(want, 0)
    for path in (1, 2):
        got, _ = lib.debug_gemm(src, w, offsets, 1, bias, relu, path=path)
        err = np.abs(got - want)[valid].max()
        assert err <= 2e-5, (path, err)
