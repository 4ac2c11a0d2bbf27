"""The strict-order host decoder (csrc/strict_decode.cc, row a19) against the reference where the reference's pruning is
ORDER-dependent: a binding --max-active (lattice-faster-decoder.cc:780-787 lets tokens past the frame's final cutoff
depending on when they are reached) and narrow beams, on the grammar graph and on the ARPA-shaped graph.  Golden:
tests/golden/strict_golden.npz (latgen-faster-mapped, one process per utterance; made by make_strict_golden.py).
The pruned state-level lattice must have the reference's state and arc counts -- they move with every token the search
keeps or drops -- and the 5-best lists must be the reference's, words and costs.  Host only: no GPU."""
import os

import numpy as np
import pytest

from conftest import golden_dir

import importlib.util

_spec = importlib.util.spec_from_file_location("make_strict_golden", os.path.join(golden_dir(), "make_strict_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from rhasspy_speech_b200 import _lib
    return _lib


@pytest.mark.parametrize("graph", ["grammar", "arpa"])
def test_strict_decoder_reproduces_order_dependent_pruning(lib, synth, tiny_model, tmp_path, graph):
    gold = np.load(os.path.join(golden_dir(), "strict_golden.npz"))
    case = [c for c in gen.CASES if c[0] == graph][0]
    if graph == "grammar":
        p, src = tiny_model, np.load(os.path.join(golden_dir(), "nbest_golden.npz"))
    else:
        p, src = synth.write_model(str(tmp_path / "a"), gen.ARPA_SPEC), np.load(os.path.join(golden_dir(), "arpa_golden.npz"))
    sizes = set()
    for u in case[1]:
        ll = src["ll_%d" % u]
        for ci, (ma, mi, beam) in enumerate(case[2]):
            tag = "%s_u%d_c%d" % (graph, u, ci)
            hyps, lat = lib.strict_decode(p.hclg, p.tid2pdf, ll, nbest=5, max_active=ma, min_active=mi, beam=beam)
            assert lat == tuple(int(x) for x in gold[tag + "_lat"]), (tag, lat, gold[tag + "_lat"])
            assert len(hyps) == int(gold[tag + "_n_hyp"]), tag
            for h, (words, gc, ac) in enumerate(hyps):
                assert words == [int(x) for x in gold["%s_h%d_words" % (tag, h)]], (tag, h)
                wg, wa = gold["%s_h%d_cost" % (tag, h)]
                assert abs(gc - wg) <= 2e-3 * max(1.0, abs(wg)) and abs(ac - wa) <= 2e-3 * max(1.0, abs(wa)), (tag, h)
            # the best path alone (no lattice) is the first hypothesis
            best, _ = lib.strict_decode(p.hclg, p.tid2pdf, ll, max_active=ma, min_active=mi, beam=beam)
            assert best[0][0] == hyps[0][0], tag
            sizes.add((u, lat))
    # the settings do bind: the same utterance yields different lattices under different limits
    assert len(sizes) > len(list(case[1]))


def test_strict_decoder_edge_cases(lib, tiny_model):
    ll = np.zeros((0, tiny_model.num_pdfs), np.float32)
    assert lib.strict_decode(tiny_model.hclg, tiny_model.tid2pdf, ll)[0] == []
    # hopeless scores: no token survives a tiny beam on some frame -> nothing decoded, as the reference prints nothing
    rng = np.random.default_rng(0)
    ll = (rng.standard_normal((30, tiny_model.num_pdfs)) * 50).astype(np.float32)
    hyps, _ = lib.strict_decode(tiny_model.hclg, tiny_model.tid2pdf, ll, beam=0.001, min_active=0)
    assert len(hyps) <= 1


def test_strict_decoder_on_the_lm_sized_graph_of_config_3(lib, synth, tmp_path):
    """The configuration whose every utterance takes the host path (BASELINE config 3): the zamia-shaped model on the
    127 k-state ARPA-shaped HCLG, two utterances cut from the reference's en_US-zamia WAVs (one of them time-reversed,
    i.e. out of grammar), log-likelihoods from the reference's own nnet3-compute.  Against latgen-faster-mapped run here,
    one process per utterance as rhasspy does: raw-lattice state and arc counts and the 3-best lists, with --max-active
    at its default (binds on some frames of ~4 k tokens) and at 2000 (binds on most).  ~20 s, needs oracle/_ref."""
    import dataclasses
    from oracle import ref_run
    if not ref_run.available():
        pytest.skip("oracle/_ref not built")
    spec = dataclasses.replace(synth.ZAMIA_LIKE, name="zamia_arpa", graph="arpa", vocab_size=2000, bigrams_per_word=20, eps_hops=2)
    p = synth.write_model(str(tmp_path / "arpa"), spec)
    utts = synth.make_utterances(2, seed=1234, pool=synth.load_pool(os.path.join(golden_dir(), "en_US-zamia")))
    utts[0] = utts[0][::-1].copy()
    wavs = []
    for i, pcm in enumerate(utts):
        wavs.append(str(tmp_path / ("u%d.wav" % i)))
        synth.write_wav(wavs[-1], pcm)
    conf = os.path.join(p.model_dir, "model", "online", "conf")
    feats = ref_run.mfcc(os.path.join(conf, "mfcc.conf"), wavs)
    iv = ref_run.ivectors_periodic(os.path.join(conf, "ivector_extractor.conf"), feats, repeat=True)
    lls = ref_run.nnet_loglikes(p.final_mdl, feats, [v[-1] for v in iv], frame_subsampling_factor=3)
    sizes = {}
    for u, ll in enumerate(lls):
        for max_active in (7000, 2000):
            raw, nb = ref_run.decode_loglikes_lattice(p.final_mdl, p.hclg, [ll], nbest=3, max_active=max_active)
            r = raw["utt00000"]
            want = [nb[k] for k in sorted(nb)]
            hyps, lat = lib.strict_decode(p.hclg, p.tid2pdf, ll, nbest=3, max_active=max_active)
            tag = (u, max_active)
            assert lat == (r["n_states"], len(r["src"])), (tag, lat, r["n_states"], len(r["src"]))
            assert [h[0] for h in hyps] == [w[0] for w in want], tag
            for (_, gc, ac), (_, wg, wa) in zip(hyps, want):
                assert abs(gc - wg) <= 2e-3 * max(1.0, abs(wg)) and abs(ac - wa) <= 2e-3 * max(1.0, abs(wa)), tag
            sizes[tag] = lat
    # --max-active does bind at this scale: the tighter limit leaves a different lattice for at least one utterance
    assert any(sizes[(u, 7000)] != sizes[(u, 2000)] for u in range(len(lls))), sizes
