"""Golden vectors for the decoder restatement on an epsilon-heavy graph (ARPA-shaped HCLG with back-off chains), made by
the reference itself: latgen-faster-mapped at beam 16 on log-likelihoods of the tiny model, raw state-level lattice
(--determinize-lattice=false) and the n-best lists of the determinised one.

    python tests/golden/make_arpa_golden.py      # writes tests/golden/arpa_golden.npz
"""
import dataclasses
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import kaldi_np as K  # noqa: E402
from oracle import ref_run  # noqa: E402
from tools import synth  # noqa: E402

SPEC = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
BEAM = 16.0
UTTS = (0, 3)


def main():
    tmp = tempfile.mkdtemp()
    p = synth.write_model(tmp, SPEC)
    utts = synth.make_utterances(6, seed=42, min_s=1.0, max_s=3.0)
    conf = os.path.join(p.model_dir, "model", "online", "conf")
    mc = K.MfccComputer(K.MfccOpts.from_conf(os.path.join(conf, "mfcc.conf")))
    s = K.IvectorSetup.from_conf(os.path.join(conf, "ivector_extractor.conf"))
    feats = [mc.compute(utts[u]) for u in UTTS]
    ivs = [K.ivector_offline(s, f) for f in feats]
    lls = ref_run.nnet_loglikes(p.final_mdl, feats, ivs, frame_subsampling_factor=3)
    raw, nb = ref_run.decode_loglikes_lattice(p.final_mdl, p.hclg, lls, nbest=5, beam=BEAM)
    out = {"beam": np.float32(BEAM)}
    for i in range(len(UTTS)):
        key = "utt%05d" % i
        out["ll_%d" % i] = lls[i].astype(np.float32)
        for f in ("src", "dst", "olabel", "graph", "acoustic"):
            out["u%d_%s" % (i, f)] = raw[key][f]
        out["u%d_n_states" % i] = np.int32(raw[key]["n_states"])
        hyps = [nb[k] for k in sorted(nb) if k.startswith(key + "-")]
        out["u%d_n_hyp" % i] = np.int32(len(hyps))
        for h, (words, g, a) in enumerate(hyps):
            out["u%d_h%d_words" % (i, h)] = np.array(words, np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "arpa_golden.npz"), **out)
    print("wrote", [int(out["u%d_n_states" % i]) for i in range(len(UTTS))], "states")


if __name__ == "__main__":
    main()
