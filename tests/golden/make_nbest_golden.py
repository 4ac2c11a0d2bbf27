"""Golden vectors for the host half of the n-best tail (csrc/nbest.cc), made by the reference itself.

Runs latgen-faster-mapped (oracle/_ref, i.e. /root/reference/kaldi compiled as is) on log-likelihoods of the
tiny synthetic model: once with --determinize-lattice=false for the raw state-level lattice, once through
`lattice-to-nbest --n=5 --acoustic-scale=S | nbest-to-linear` for the expected word sequences and costs.

    python tests/golden/make_nbest_golden.py      # writes tests/golden/nbest_golden.npz
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import kaldi_np as K  # noqa: E402
from oracle import ref_run  # noqa: E402
from tools import synth  # noqa: E402


def main():
    tmp = tempfile.mkdtemp()
    p = synth.write_model(tmp, synth.TINY)
    utts = synth.make_utterances(6, seed=42, min_s=1.0, max_s=3.0)
    conf = os.path.join(p.model_dir, "model", "online", "conf")
    mc = K.MfccComputer(K.MfccOpts.from_conf(os.path.join(conf, "mfcc.conf")))
    s = K.IvectorSetup.from_conf(os.path.join(conf, "ivector_extractor.conf"))
    feats = [mc.compute(u) for u in utts]
    ivs = [K.ivector_offline(s, f) for f in feats]
    lls = ref_run.nnet_loglikes(p.final_mdl, feats, ivs, frame_subsampling_factor=3)
    out = {"ll_%d" % u: l.astype(np.float32) for u, l in enumerate(lls)}     # input of the decoder restatement
    out["ll_scales"] = np.array([1.0, 0.3, 0.3, 0.3], np.float32)            # log-likelihood scale of case c<i>
    cases = []
    # (log-likelihood scale, lattice-to-nbest --acoustic-scale): flatter scores give bigger lattices
    for ci, (ll_scale, nb_scale) in enumerate(((1.0, 1.0), (0.3, 1.0), (0.3, 0.5), (0.3, 2.0))):
        mats = [np.ascontiguousarray(l * np.float32(ll_scale)) for l in lls]
        raw, nb = ref_run.decode_loglikes_lattice(p.final_mdl, p.hclg, mats, nbest=5, acoustic_scale=nb_scale)
        for u in range(len(mats)):
            key = "utt%05d" % u
            L = raw[key]
            tag = "c%d_u%d" % (ci, u)
            for f in ("src", "dst", "olabel", "graph", "acoustic"):
                out[tag + "_" + f] = L[f]
            out[tag + "_n_states"] = np.int32(L["n_states"])
            hyps = [nb[k] for k in sorted(nb) if k.startswith(key + "-")]
            out[tag + "_n_hyp"] = np.int32(len(hyps))
            for h, (words, g, a) in enumerate(hyps):
                out["%s_h%d_words" % (tag, h)] = np.array(words, np.int32)
                out["%s_h%d_cost" % (tag, h)] = np.array([g, a], np.float64)
            cases.append((tag, nb_scale))
    out["cases"] = np.array([c[0] for c in cases])
    out["scales"] = np.array([c[1] for c in cases], np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nbest_golden.npz"), **out)
    print("wrote", len(cases), "lattices")


if __name__ == "__main__":
    main()
