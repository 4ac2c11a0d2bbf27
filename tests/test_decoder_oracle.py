"""oracle/decoder_np.py (the CPU restatement of LatticeFasterDecoder + FinalizeDecoding + GetRawLattice) pinned
against the reference: the pruned state-level lattice must equal the one `latgen-faster-mapped
--determinize-lattice=false` wrote for the same log-likelihoods (tests/golden/nbest_golden.npz, made by
tests/golden/make_nbest_golden.py from oracle/_ref), and the n-best read from it must equal the reference's."""
import collections
import os

import numpy as np
import pytest

from conftest import golden_dir


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(golden_dir(), "nbest_golden.npz"))


def arc_multiset(lat):
    """Arcs of a state-level lattice up to state renumbering: (is final, olabel, graph cost, acoustic cost)."""
    return collections.Counter((bool(d < 0), int(o), round(float(g), 3), round(float(a), 2))
                               for d, o, g, a in zip(lat["dst"], lat["olabel"], lat["graph"], lat["acoustic"]))


def arcs_close(lat, ref):
    """Robust form of the same comparison for big lattices, where a 2-digit key flips on rounding: same labels, and the
    sorted graph / acoustic costs agree to the precision of the text archive (6 significant digits)."""
    if collections.Counter(int(o) for o in lat["olabel"]) != collections.Counter(int(o) for o in ref["olabel"]):
        return False
    if int((np.asarray(lat["dst"]) < 0).sum()) != int((np.asarray(ref["dst"]) < 0).sum()):
        return False
    for f in ("graph", "acoustic"):
        a, b = np.sort(np.asarray(lat[f], np.float64)), np.sort(np.asarray(ref[f], np.float64))
        if not np.allclose(a, b, rtol=1e-5, atol=1e-5):
            return False
    return True


def test_restated_decoder_reproduces_reference_lattices(golden, tiny_model):
    from oracle import decoder_np as D
    import __graft_entry__ as ge
    ge.build()
    from rhasspy_speech_b200 import _lib
    fst = D.ConstFst(tiny_model.hclg)
    for ci in (0, 1):                       # log-likelihood scales 1.0 and 0.3
        scale = np.float32(golden["ll_scales"][ci])
        for u in range(6):
            tag = "c%d_u%d" % (ci, u)
            r = D.decode(fst, golden["ll_%d" % u] * scale, tiny_model.tid2pdf)
            lat = r["lattice"]
            ref = {f: golden["%s_%s" % (tag, f)] for f in ("src", "dst", "olabel", "graph", "acoustic")}
            assert lat["n_states"] == int(golden[tag + "_n_states"]), tag
            assert len(lat["src"]) == len(ref["src"]), tag
            diff = arc_multiset(lat) - arc_multiset(ref)
            assert sum(diff.values()) <= 2, (tag, diff)        # rounding of the 2-digit key only
            # the n-best tail on the restated lattice
            got = _lib.lattice_nbest(lat["src"], lat["dst"], lat["olabel"], lat["graph"], lat["acoustic"], lat["n_states"], 5)
            want = [[int(x) for x in golden["%s_h%d_words" % (tag, h)]] for h in range(int(golden[tag + "_n_hyp"]))]
            assert [w for w, _, _ in got] == want, tag


def test_restated_decoder_without_periodic_pruning_is_identical(golden, tiny_model):
    """PruneActiveTokens every 25 frames only removes what FinalizeDecoding removes anyway: the device decoder,
    which prunes once at the end, relies on it."""
    from oracle import decoder_np as D
    fst = D.ConstFst(tiny_model.hclg)
    ll = golden["ll_3"] * np.float32(0.3)
    a = D.decode(fst, ll, tiny_model.tid2pdf)["lattice"]
    b = D.decode(fst, ll, tiny_model.tid2pdf, prune_interval=0)["lattice"]
    assert a["n_states"] == b["n_states"] and arc_multiset(a) == arc_multiset(b)


def test_restated_decoder_on_epsilon_heavy_graph(tmp_path_factory, synth):
    """ARPA-shaped HCLG (back-off epsilon chains, several hops deep): ProcessNonemitting's closure and the epsilon links of
    the lattice.  Golden: tests/golden/arpa_golden.npz (latgen-faster-mapped at beam 16, made by make_arpa_golden.py)."""
    import dataclasses
    import __graft_entry__ as ge
    ge.build()
    from oracle import decoder_np as D
    from rhasspy_speech_b200 import _lib
    g = np.load(os.path.join(golden_dir(), "arpa_golden.npz"))
    spec = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
    p = synth.write_model(str(tmp_path_factory.mktemp("arpa")), spec)
    fst = D.ConstFst(p.hclg)
    assert sum(len(a) for a in fst.eps) > 500        # back-off arcs
    for i in range(2):
        lat = D.decode(fst, g["ll_%d" % i], p.tid2pdf, beam=float(g["beam"]))["lattice"]
        ref = {f: g["u%d_%s" % (i, f)] for f in ("src", "dst", "olabel", "graph", "acoustic")}
        assert lat["n_states"] == int(g["u%d_n_states" % i]) and len(lat["src"]) == len(ref["src"]), i
        assert arcs_close(lat, ref), i
        got = _lib.lattice_nbest(lat["src"], lat["dst"], lat["olabel"], lat["graph"], lat["acoustic"], lat["n_states"], 5)
        want = [[int(x) for x in g["u%d_h%d_words" % (i, h)]] for h in range(int(g["u%d_n_hyp" % i]))]
        assert [w for w, _, _ in got] == want, i
