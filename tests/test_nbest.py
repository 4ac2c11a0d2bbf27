"""Host half of the n-best tail (csrc/nbest.cc through rs_debug_lattice_nbest) against the reference:
raw state-level lattices written by latgen-faster-mapped --determinize-lattice=false, expected output from the
reference's determinise -> lattice-to-nbest --n=5 --acoustic-scale=S -> nbest-to-linear on the same decode
(tests/golden/make_nbest_golden.py).  No GPU involved."""
import os

import numpy as np
import pytest

from conftest import golden_dir


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from rhasspy_speech_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(golden_dir(), "nbest_golden.npz"))


def _case(golden, tag):
    lat = [golden["%s_%s" % (tag, f)] for f in ("src", "dst", "olabel", "graph", "acoustic")]
    want = [([int(x) for x in golden["%s_h%d_words" % (tag, h)]], golden["%s_h%d_cost" % (tag, h)])
            for h in range(int(golden[tag + "_n_hyp"]))]
    return lat, int(golden[tag + "_n_states"]), want


def test_nbest_matches_reference_determinize_and_shortest_paths(lib, golden):
    n_multi = 0
    for tag, scale in zip(golden["cases"], golden["scales"]):
        lat, n_states, want = _case(golden, str(tag))
        got = lib.lattice_nbest(*lat, n_states, 5, float(scale))
        assert len(got) == len(want), (tag, len(got), len(want))
        n_multi += len(want) > 1
        for h, ((words, g, a), (wwords, wcost)) in enumerate(zip(got, want)):
            assert words == wwords, (tag, h, words, wwords)        # same word sequences, same order
            # costs of the sequence's best path: the archives print 7 significant digits
            assert abs(g - wcost[0]) <= 2e-4 * max(1.0, abs(wcost[0])) and abs(a - wcost[1]) <= 2e-4 * max(1.0, abs(wcost[1]))
        # ranked by graph + scale * acoustic
        tot = [g + float(scale) * a for _, g, a in got]
        assert all(tot[i] <= tot[i + 1] + 1e-3 for i in range(len(tot) - 1)), (tag, tot)
    assert n_multi >= 12        # the fixture does exercise lists with several hypotheses


def test_nbest_prefix_property_and_edge_cases(lib, golden):
    """n = 1 is the head of n = 5; an empty lattice and a lattice without a final state give no hypothesis."""
    lat, n_states, want = _case(golden, "c1_u0")
    one = lib.lattice_nbest(*lat, n_states, 1)
    five = lib.lattice_nbest(*lat, n_states, 5)
    assert len(one) == 1 and one[0] == five[0]
    z = np.zeros(0, np.int32)
    assert lib.lattice_nbest(z, z, z, np.zeros(0, np.float32), np.zeros(0, np.float32), 0, 3) == []
    src, dst, ol, g, a = lat
    keep = dst >= 0                                 # drop every final weight
    assert lib.lattice_nbest(src[keep], dst[keep], ol[keep], g[keep], a[keep], n_states, 3) == []
    # two parallel paths with the same words: one hypothesis, carrying the cheaper path's costs
    src = np.array([0, 0, 1, 2, 3], np.int32)
    dst = np.array([1, 2, 3, 3, -1], np.int32)
    ol = np.array([7, 7, 0, 0, 0], np.int32)
    g = np.array([1.0, 0.5, 0.0, 0.0, 0.25], np.float32)
    a = np.array([2.0, 3.0, 0.0, 0.0, 0.0], np.float32)
    assert lib.lattice_nbest(src, dst, ol, g, a, 4, 5) == [([7], 1.25, 2.0)]
