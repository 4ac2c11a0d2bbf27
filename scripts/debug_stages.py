"""Per-stage diff against the reference / oracle for a model variant and batch (debugging aid)."""
import dataclasses, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rhasspy_speech_b200 import _lib
from tools import synth
from oracle import ref_run, kaldi_np as K

which = sys.argv[1] if len(sys.argv) > 1 else "arpa"
nrev = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which == "arpa":
    spec = dataclasses.replace(synth.TINY, name="tiny_arpa", seed=11, graph="arpa", vocab_size=300, bigrams_per_word=8, eps_hops=2)
else:
    spec = dataclasses.replace(synth.TINY, name="v", seed=21, chain=False, frame_subsampling_factor=1, log_softmax=True,
                               priors=True, tdnnf_strides=(1, 0, 1, 1))
tmp = tempfile.mkdtemp()
p = synth.write_model(tmp, spec)
utts = synth.make_utterances(6, seed=42, min_s=1.0, max_s=3.0)
utts = list(utts) + [u[::-1].copy() for u in utts[:nrev]]
wavs = []
for i, u in enumerate(utts):
    w = os.path.join(tmp, "u%03d.wav" % i)
    synth.write_wav(w, u)
    wavs.append(w)
m = _lib.Model(p.final_mdl, p.online_conf, 0)
g = _lib.Graph(p.hclg, p.words_txt, 0)
dec = _lib.Decoder(m, g)
hyp = dec.decode_pcm(utts)
conf = os.path.join(p.model_dir, "model", "online", "conf")
ref_feats = ref_run.mfcc(os.path.join(conf, "mfcc.conf"), wavs)
s = K.IvectorSetup.from_conf(os.path.join(conf, "ivector_extractor.conf"))
feats = [dec.fetch(0, u) for u in range(len(utts))]
ivs = [dec.fetch(1, u)[0] for u in range(len(utts))]
lls = [dec.fetch(2, u) for u in range(len(utts))]
want_ll = ref_run.nnet_loglikes(p.final_mdl, feats, ivs, frame_subsampling_factor=spec.frame_subsampling_factor)
want, _, _ = ref_run.transcribe_wavs(p.final_mdl, p.online_conf, p.hclg, p.words_txt, wavs)
dl = ref_run.decode_loglikes(p.final_mdl, p.hclg, lls)
for u in range(len(utts)):
    iv_want = K.ivector_offline(s, feats[u])
    print("utt %d frames %d mfcc %.2e ivec %.2e ll %.2e (shape %s/%s) | pipeline %s | ours==ref(ll->dec) %s" % (
        u, feats[u].shape[0], np.abs(feats[u] - ref_feats[u]).max(), np.abs(ivs[u] - iv_want).max(),
        np.abs(lls[u] - want_ll[u]).max() if lls[u].shape == want_ll[u].shape else -1, lls[u].shape, want_ll[u].shape,
        "SAME" if hyp.words[u] == want.get("utt%05d-1" % u) else "DIFF ours %s ref %s" % (hyp.words[u][:5], (want.get("utt%05d-1" % u) or [])[:5]),
        hyp.words[u] == dl.get("utt%05d-1" % u)))
