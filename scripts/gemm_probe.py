"""GPU probe of the affine-layer kernels: error of the tcgen05 3xTF32 path and of the fp32 CUDA-core
path against an fp64 product, and time per launch, on the layer shapes of the zamia-like model.
Usage (on a B200): python scripts/gemm_probe.py [rows]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rhasspy_speech_b200 import _lib  # noqa: E402


def ref64(src, w, offsets, stride, bias, relu):
    rows, k = src.shape
    m = max(rows // stride, 1)
    out = np.zeros((m, w.shape[0]), dtype=np.float64)
    valid = np.ones(m, dtype=bool)
    r = np.arange(m) * stride
    for i, o in enumerate(offsets):
        idx = r + o
        valid &= (idx >= 0) & (idx < rows)
        out += src[np.clip(idx, 0, rows - 1)].astype(np.float64) @ w[:, i * k:(i + 1) * k].astype(np.float64).T
    if bias is not None:
        out += bias.astype(np.float64)
    if relu:
        out = np.maximum(out, 0)
    return out, valid


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rng = np.random.default_rng(1)
    shapes = [  # (k, n, offsets, stride, bias, relu)
        (40, 220, (-1, 0, 1), 1, True, False),
        (220, 1024, (0,), 1, True, True),
        (1024, 128, (-1, 0), 1, False, False),
        (128, 1024, (0, 1), 1, True, True),
        (128, 1024, (0, 1), 3, True, True),
        (1024, 128, (-3, 0), 1, False, False),
        (1024, 192, (0,), 1, False, False),
        (192, 3026, (0,), 1, True, False),
        (100, 64, (0,), 1, False, False),
        (36, 40, (-2, -1, 0, 1), 2, True, True),
    ]
    res = []
    for k, n, offsets, stride, use_bias, relu in shapes:
        src = rng.standard_normal((rows, k)).astype(np.float32)
        src *= np.exp(rng.uniform(-2, 2, size=(1, k))).astype(np.float32)
        w = (rng.standard_normal((n, k * len(offsets))) / np.sqrt(k * len(offsets))).astype(np.float32)
        bias = rng.standard_normal(n).astype(np.float32) if use_bias else None
        want, valid = ref64(src, w, offsets, stride, bias, relu)
        mag = np.zeros_like(want)
        r = np.arange(want.shape[0]) * stride
        for i, o in enumerate(offsets):
            mag += np.abs(src[np.clip(r + o, 0, rows - 1)]).astype(np.float64) @ np.abs(w[:, i * k:(i + 1) * k]).astype(np.float64).T
        row = {"k": k, "n": n, "offsets": offsets, "stride": stride, "rows": rows}
        for path, name in ((0, "simt"), (1, "tc"), (2, "tc_split")):
            got, ms = _lib.debug_gemm(src, w, offsets, stride, bias, relu, path=path, iters=5)
            err = (got.astype(np.float64) - want)[valid]
            row[name] = {"max_abs": float(np.abs(err).max()), "rms": float(np.sqrt((err ** 2).mean())),
                         "mean_signed_rel_to_sumabs": float((err / mag[valid]).mean()),
                         "max_rel_to_sumabs": float(np.abs(err / mag[valid]).max()), "ms": ms,
                         "tflops": 2.0 * want.shape[0] * n * k * len(offsets) / (ms * 1e-3) / 1e12 if ms > 0 else None}
        print(json.dumps(row), flush=True)
        res.append(row)
    return res


if __name__ == "__main__":
    main()
