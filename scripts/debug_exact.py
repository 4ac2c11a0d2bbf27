"""GPU box probe: the device search (decode_small.cu on small graphs, decode.cu + safe-frame flags otherwise) against
the strict-order host decoder on the same log-likelihoods -- words, path costs and pruned-lattice sizes per utterance."""
import dataclasses, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rhasspy_speech_b200 import _lib
from tools import synth


def compare(name, spec, n, **opts):
    tmp = tempfile.mkdtemp()
    p = synth.write_model(tmp, spec)
    pool = synth.load_pool()
    utts = synth.make_utterances(n, seed=99, pool=pool)
    for i in range(0, n, 7):
        utts[i] = utts[i][::-1].copy()
    m, g = _lib.Model(p.final_mdl, p.online_conf, 0), _lib.Graph(p.hclg, p.words_txt, 0)
    dev = _lib.Decoder(m, g, strict_fallback=0, **opts)
    host = _lib.Decoder(m, g, strict_fallback=2, **opts)
    for nb in (1, 4):
        dev.set_nbest(nb)
        host.set_nbest(nb)
        t0 = time.time()
        a = dev.decode_pcm(utts)
        ta = time.time() - t0
        td = dev.timings()
        b = host.decode_pcm(utts)
        th = host.timings()
        bad = flagged = bad_flagged = 0
        for u in range(n):
            fl = bool(a.status[u] & 16)
            flagged += fl
            same = a.words[u] == b.words[u] and abs(a.graph_cost[u] - b.graph_cost[u]) < 1e-3 and abs(a.acoustic_cost[u] - b.acoustic_cost[u]) < 2e-2
            if nb > 1:
                same = same and [h[0] for h in a.nbest[u]] == [h[0] for h in b.nbest[u]]
                la, lb = dev.fetch(5, u), host.fetch(5, u)
                same = same and la.shape == lb.shape
            if not same:
                if fl:
                    bad_flagged += 1
                else:
                    bad += 1
                    if bad <= 3:
                        print("  MISMATCH utt", u, "status", a.status[u], a.words[u], b.words[u], a.graph_cost[u], b.graph_cost[u],
                              a.acoustic_cost[u], b.acoustic_cost[u], (dev.fetch(5, u).shape, host.fetch(5, u).shape) if nb > 1 else "")
        import collections
        rules = collections.Counter()
        for st in a.status:
            for bit, nm in ((256, "min-active"), (512, "max-active count"), (1024, "extra in cutoff"), (2048, "last frame"), (4096, "maybe link")):
                rules[nm] += bool(st & bit)
        print("  rules:", dict(rules))
        if nb > 1 and bad:
            for u in range(n):
                if a.status[u] & 16:
                    continue
                la, lb = dev.fetch(5, u), host.fetch(5, u)
                if la.shape == lb.shape:
                    continue
                key = lambda A: collections.Counter((int(r[1] < 0), int(r[2]), round(float(r[3]), 3), round(float(r[4]), 2)) for r in A)
                d1, d2 = key(lb) - key(la), key(la) - key(lb)
                print("  utt", u, "frames", a.num_frames[u], "host-only arcs:", list(d1.items())[:12], "| device-only:", list(d2.items())[:6])
                break
        print("%s nbest=%d: %d utts, %d flagged (%d of them differ), UNFLAGGED MISMATCHES %d | device decode %.2f ms wall %.1f ms | strict %d utts %.1f ms | tok/frame %.0f"
              % (name, nb, n, flagged, bad_flagged, bad, td["decode_ms"], ta * 1e3, th["strict_utts"], th["strict_ms"],
                 td["tokens_expanded"] / max(1, td["frames_decoded"])), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny", "zamia", "tiny_arpa", "tiny_arpa300", "zamia_arpa"]
    if "tiny" in which:
        compare("tiny grammar", synth.TINY, 48)
        compare("tiny grammar max-active 100", synth.TINY, 48, max_active=100, min_active=20)
    if "zamia" in which:
        compare("zamia grammar", synth.ZAMIA_LIKE, 64)
    arpa = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
    if "tiny_arpa" in which:
        compare("tiny arpa", arpa, 32)
    if "tiny_arpa300" in which:
        compare("tiny arpa max-active 300", arpa, 32, max_active=300)
        compare("tiny arpa max-active 1000", arpa, 32, max_active=1000)
    if "zamia_arpa" in which:
        spec = dataclasses.replace(synth.ZAMIA_LIKE, name="zamia_arpa", graph="arpa", vocab_size=2000, bigrams_per_word=20, eps_hops=2)
        compare("zamia arpa 127k", spec, 32, max_tokens_per_utt=1 << 21)
