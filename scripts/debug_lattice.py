"""Diff the device lattice (rs_debug_fetch item 5) against the reference's raw lattice on the same log-likelihoods."""
import collections
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_run  # noqa: E402
from rhasspy_speech_b200 import _lib  # noqa: E402
from tools import synth  # noqa: E402

tmp = tempfile.mkdtemp()
p = synth.write_model(tmp, synth.TINY)
utts = synth.make_utterances(6, seed=42, min_s=1.0, max_s=3.0)
dec = _lib.Decoder(_lib.Model(p.final_mdl, p.online_conf, 0), _lib.Graph(p.hclg, p.words_txt, 0))
dec.decode_pcm(utts)
lls = [dec.fetch(2, u) for u in range(len(utts))]
for scale in (1.0, 0.3):
    mats = [np.ascontiguousarray(l * np.float32(scale)) for l in lls]
    raw, want = ref_run.decode_loglikes_lattice(p.final_mdl, p.hclg, mats, nbest=5)
    dec.set_nbest(5)
    got = dec.decode_loglikes(mats)
    for u in range(len(mats)):
        L = raw["utt%05d" % u]
        A = dec.fetch(5, u)
        n_states = int(max(A[:, 0].max(), A[:, 1].max())) + 1
        key = lambda ol, g, a: (int(ol), round(float(g), 3), round(float(a), 2))
        mine = collections.Counter(key(r[2], r[3], r[4]) if r[1] >= 0 else ("final", round(float(r[3]), 3)) for r in A)
        theirs = collections.Counter(key(o, g, a) if d >= 0 else ("final", round(float(g), 3))
                                     for d, o, g, a in zip(L["dst"], L["olabel"], L["graph"], L["acoustic"]))
        print(scale, u, "states", n_states, L["n_states"], "arcs", len(A), len(L["src"]), "only mine", sum((mine - theirs).values()),
              "only ref", sum((theirs - mine).values()))
        for k, v in list((mine - theirs).items())[:6]:
            rows = [r for r in A if (key(r[2], r[3], r[4]) if r[1] >= 0 else ("final", round(float(r[3]), 3))) == k]
            print("    mine only", k, v, rows[:2])
        for k, v in list((theirs - mine).items())[:6]:
            print("    ref only", k, v)
        # in-degree-0 states other than the start
        indeg = collections.Counter(int(r[1]) for r in A if r[1] >= 0)
        orphans = [s for s in range(1, n_states) if indeg[s] == 0]
        print("    orphans (mine)", orphans[:10])
    if os.environ.get("RS_DUMP"):
        np.savez_compressed(os.path.join(os.environ["RS_DUMP"], "lat_scale%g.npz" % scale),
                            **{"mine_%d" % u: dec.fetch(5, u) for u in range(len(mats))},
                            **{"ref_%d_%s" % (u, f): raw["utt%05d" % u][f] for u in range(len(mats)) for f in ("src", "dst", "olabel", "graph", "acoustic")})
