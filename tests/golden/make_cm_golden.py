"""Golden vectors for the CompressedMatrix reader: the same matrix written by the reference's copy-feats in its three
compressed formats (CM: one byte + per-column headers, CM2: two bytes, CM3: one byte), and as the reference itself
expands each of them (copy-matrix to text).

    python tests/golden/make_cm_golden.py      # writes tests/golden/cm_golden.npz
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_run  # noqa: E402

# copy-feats --compression-method: 2 = speech features (CM), 3 = two-byte auto (CM2), 5 = one-byte auto (CM3)
METHODS = {"CM": 2, "CM2": 3, "CM3": 5}


def main():
    rng = np.random.default_rng(5)
    mat = (rng.standard_normal((37, 13)) * np.linspace(0.5, 30.0, 13)[None, :] + np.linspace(-5, 90, 13)[None, :]).astype(np.float32)
    out = {"source": mat}
    with tempfile.TemporaryDirectory() as tmp:
        ref_run.write_mat_ark(os.path.join(tmp, "in.ark"), {"m": mat})
        for tok, method in METHODS.items():
            ark = os.path.join(tmp, tok + ".ark")
            ref_run.run("copy-feats --compress=true --compression-method=%d ark:%s/in.ark ark:%s 2>/dev/null" % (method, tmp, ark))
            raw = open(ark, "rb").read()
            obj = raw[raw.index(b" ") + 1:]          # the Kaldi object behind the archive key
            assert obj[:2] == b"\0B" and obj[2:2 + len(tok) + 1] == tok.encode() + b" ", (tok, obj[:8])
            out[tok + "_file"] = np.frombuffer(obj, np.uint8)
            txt, _ = ref_run.run("copy-feats ark:%s ark,t:- 2>/dev/null" % ark)
            rows = [[float(x) for x in line.replace("]", "").split()] for line in txt.decode().splitlines()[1:] if line.strip()]
            out[tok + "_expanded"] = np.array(rows, np.float32)
            assert out[tok + "_expanded"].shape == mat.shape
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cm_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
