"""Golden vectors for the strict-order host decoder (csrc/strict_decode.cc), made by the reference itself with pruning
settings under which its result depends on the ORDER it visits tokens in: a binding --max-active and narrow beams on the
grammar graph and on the ARPA-shaped graph.  One latgen-faster-mapped process per utterance, as rhasspy runs its decoder
(the reference's token hash keeps its grown size between the utterances of one process, and with it the order).
Log-likelihoods: those of tests/golden/nbest_golden.npz / arpa_golden.npz.

    python tests/golden/make_strict_golden.py      # writes tests/golden/strict_golden.npz
"""
import dataclasses
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_run  # noqa: E402
from tools import synth  # noqa: E402

ARPA_SPEC = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
# (graph, utterances, [(max_active, min_active, beam)])
CASES = [("grammar", range(6), [(300, 20, 24.0), (100, 20, 24.0), (50, 20, 24.0), (7000, 200, 10.0)]),
         ("arpa", range(2), [(7000, 200, 24.0), (1000, 200, 24.0), (300, 200, 24.0), (100, 50, 24.0), (7000, 200, 12.0),
                             (1000, 200, 12.0)])]


def main():
    tmp = tempfile.mkdtemp()
    gold = os.path.join(ROOT, "tests", "golden")
    src = {"grammar": (synth.write_model(tmp + "/g", synth.TINY), np.load(os.path.join(gold, "nbest_golden.npz"))),
           "arpa": (synth.write_model(tmp + "/a", ARPA_SPEC), np.load(os.path.join(gold, "arpa_golden.npz")))}
    out = {}
    n = 0
    for graph, utts, settings in CASES:
        p, g = src[graph]
        for u in utts:
            ll = g["ll_%d" % u]
            for ci, (ma, mi, beam) in enumerate(settings):
                raw, nb = ref_run.decode_loglikes_lattice(p.final_mdl, p.hclg, [ll], nbest=5, max_active=ma, min_active=mi, beam=beam)
                tag = "%s_u%d_c%d" % (graph, u, ci)
                r = raw["utt00000"]
                out[tag + "_lat"] = np.array([r["n_states"], len(r["src"])], np.int32)
                hyps = [nb[k] for k in sorted(nb)]
                out[tag + "_n_hyp"] = np.int32(len(hyps))
                for h, (words, gc, ac) in enumerate(hyps):
                    out["%s_h%d_words" % (tag, h)] = np.array(words, np.int32)
                    out["%s_h%d_cost" % (tag, h)] = np.array([gc, ac], np.float32)
                n += 1
    np.savez_compressed(os.path.join(gold, "strict_golden.npz"), **out)
    print("wrote", n, "cases")


if __name__ == "__main__":
    main()
