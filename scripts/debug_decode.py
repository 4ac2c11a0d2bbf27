"""Decoder-only comparison on the failing fixtures: our log-likelihoods through the reference's
latgen-faster-mapped (with path costs) and through rs_decode_loglikes at several beams."""
import dataclasses, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rhasspy_speech_b200 import _lib
from tools import synth
from oracle import ref_run

which = sys.argv[1] if len(sys.argv) > 1 else "arpa"
if which == "arpa":
    spec = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
else:
    spec = dataclasses.replace(synth.TINY, name="v", seed=21, chain=False, frame_subsampling_factor=1, log_softmax=True,
                               priors=True, tdnnf_strides=(1, 0, 1, 1))
tmp = tempfile.mkdtemp()
p = synth.write_model(tmp, spec)
utts = synth.make_utterances(6, seed=42, min_s=1.0, max_s=3.0)
m = _lib.Model(p.final_mdl, p.online_conf, 0)
g = _lib.Graph(p.hclg, p.words_txt, 0)
dec = _lib.Decoder(m, g)
hyp = dec.decode_pcm(utts)
lls = [dec.fetch(2, u) for u in range(len(utts))]


def ref_decode(mats, beam=24.0):
    keys = ["utt%05d" % i for i in range(len(mats))]
    ref_run.write_mat_ark(os.path.join(tmp, "ll.ark"), dict(zip(keys, mats)))
    out, _ = ref_run.run("latgen-faster-mapped --acoustic-scale=1.0 --beam=%g --max-active=7000 --lattice-beam=8 --allow-partial=true %s %s ark:%s/ll.ark ark:- 2>/dev/null | "
                         "lattice-to-nbest --n=1 --acoustic-scale=1.0 ark:- ark:- 2>/dev/null | "
                         "nbest-to-linear ark:- ark:/dev/null ark,t:%s/tr.txt ark,t:%s/lm.txt ark,t:%s/ac.txt 2>/dev/null" % (beam, p.final_mdl, p.hclg, tmp, tmp, tmp, tmp))
    res = {}
    for name in ("tr", "lm", "ac"):
        for line in open(os.path.join(tmp, name + ".txt")):
            parts = line.split()
            res.setdefault(parts[0], {})[name] = parts[1:]
    return res


for scale in (1.0,):
    ref = ref_decode(lls)
    ref_big = ref_decode(lls, beam=100.0)
    for beam in (24.0, 100.0):
        d2 = _lib.Decoder(m, g, beam=beam)
        got = d2.decode_loglikes(lls)
        for u in range(len(utts)):
            r = (ref if beam == 24.0 else ref_big).get("utt%05d-1" % u, {})
            rw = [int(x) for x in r.get("tr", [])]
            rc = float(r.get("lm", ["nan"])[0]) + float(r.get("ac", ["nan"])[0])
            oc = float(got.graph_cost[u] + got.acoustic_cost[u])
            print("beam %5.0f utt %d %s ours cost %.4f (g %.4f a %.4f) ref cost %.4f (g %s a %s) | ours %s ref %s" % (
                beam, u, "SAME" if got.words[u] == rw else "DIFF", oc, got.graph_cost[u], got.acoustic_cost[u], rc,
                r.get("lm", ["?"])[0], r.get("ac", ["?"])[0], got.words[u][:8], rw[:8]))
print("full pipeline words:", [w[:6] if w else w for w in hyp.words])
