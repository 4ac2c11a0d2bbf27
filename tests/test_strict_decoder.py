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
