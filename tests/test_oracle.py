"""The CPU restatement (oracle/kaldi_np.py) against golden vectors produced by the reference itself
(tests/golden/make_golden.py ran oracle/_ref, i.e. /root/reference/kaldi compiled as is)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import golden_dir


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(golden_dir(), "golden.npz"))


@pytest.fixture(scope="module")
def K():
    from oracle import kaldi_np
    return kaldi_np


def _conf(tiny_model):
    return os.path.join(tiny_model.model_dir, "model", "online", "conf")


def test_generator_is_stable(golden, tiny_model):
    """The seeded fixture generator must reproduce the files the golden vectors were made with."""
    for key, path in (("sha_final_mdl", tiny_model.final_mdl), ("sha_hclg", tiny_model.hclg)):
        with open(path, "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == bytes(golden[key]).decode(), key


def test_window_and_fft_are_bit_exact(golden, K):
    """srfft.cc split-radix network + feature-window.cc restated op for op: identical bits."""
    mc = K.MfccComputer(K.MfccOpts(num_bins=40, num_ceps=40, use_energy=False, dither=0.0, high_freq=-400))
    frames = golden["probe_frames"]
    for i in range(frames.shape[0]):
        win = mc.windows(frames[i])
        assert win.shape == (1, 512)
        assert np.array_equal(win[0], golden["probe_win"][i])
    spec = mc.fft.compute(golden["probe_win"])
    assert np.array_equal(spec, golden["probe_fft"])


def test_mfcc_matches_reference(golden, K, tiny_model):
    mc = K.MfccComputer(K.MfccOpts.from_conf(os.path.join(_conf(tiny_model), "mfcc.conf")))
    for i in range(4):
        got = mc.compute(golden["pcm_%d" % i])
        want = golden["mfcc_%d" % i]
        assert got.shape == want.shape
        # only the BLAS dot products (mel, DCT) differ in summation order: a few ulp of C0 ~ 1e2
        assert np.abs(got - want).max() <= 2e-4


def test_mfcc_edge_cases(K):
    mc = K.MfccComputer(K.MfccOpts(num_bins=40, num_ceps=40, use_energy=False, dither=0.0, high_freq=-400))
    assert mc.compute(np.zeros(0, np.int16)).shape == (0, 40)
    assert mc.compute(np.zeros(399, np.int16)).shape == (0, 40)
    one = mc.compute(np.zeros(400, np.int16))       # digital silence: floored mel energies
    assert one.shape == (1, 40) and np.isfinite(one).all()
    assert mc.compute(np.full(560, 1000, np.int16)).shape == (2, 40)


def test_periodic_ivectors_match_reference(golden, K, tiny_model):
    s = K.IvectorSetup.from_conf(os.path.join(_conf(tiny_model), "ivector_extractor.conf"))
    for i in range(4):
        iv = K.ivectors_periodic(s, golden["mfcc_%d" % i])
        got = iv.astype(np.float32)
        got[:, 0] = got[:, 0] - np.float32(s.prior_offset)
        want = golden["ivp_%d" % i]
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-5


def test_offline_ivector_is_stable(golden, K, tiny_model):
    s = K.IvectorSetup.from_conf(os.path.join(_conf(tiny_model), "ivector_extractor.conf"))
    for i in range(4):
        assert np.abs(K.ivector_offline(s, golden["mfcc_%d" % i]) - golden["ivo_%d" % i]).max() <= 1e-6


def test_float64_forward_matches_reference_loglikes(golden, synth, tiny_model):
    """The generator's float64 forward (same architecture) agrees with nnet3-compute to fp32 accuracy."""
    for i in range(4):
        f64 = synth.nnet_forward(tiny_model.nnet_params, golden["mfcc_%d" % i].astype(np.float64),
                                 golden["ivo_%d" % i].astype(np.float64))[::3]
        want = golden["ll_%d" % i]
        assert f64.shape == want.shape
        assert np.abs(f64 - want).max() <= 1e-4


def test_cmvn_window_and_smoothing(K):
    """Sliding window (600), global smoothing (200 frames) -- online-feature.cc:337-452."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((700, 3)).astype(np.float32) * 5 + 2
    g = np.zeros((2, 4))
    g[0, :3] = [10.0, -20.0, 30.0]
    g[0, 3] = 10.0
    out = K.online_cmvn(x, g, K.CmvnOpts())
    # frame 0: one frame of data + 200 frames of the global mean
    mean0 = (x[0].astype(np.float64) + 200.0 * g[0, :3] / 10.0) / 201.0
    assert np.allclose(out[0], x[0] - mean0, atol=1e-5)
    # frame 650: the window holds exactly frames 51..650, no smoothing
    mean = x[51:651].astype(np.float64).mean(axis=0)
    assert np.allclose(out[650], x[650] - mean, atol=1e-5)


def test_online_schedule_properties():
    """The stream binary's iVector schedule (oracle restatement): solves grow monotonically, the last one has
    seen every frame, every chunk maps to a solve, and a chunk never uses frames that had not arrived."""
    from oracle import kaldi_np as K
    for nsamp in (0, 399, 400, 5000, 16000, 47311, 80000):
        T = 0 if nsamp < 400 else 1 + (nsamp - 400) // 160
        for chunk, right, sf in ((24, 8, 3), (24, 28, 3), (24, 0, 1), (21, 5, 3)):
            solves, chunk_solve = K.online_schedule(nsamp, T, chunk, right, 3, sf)
            assert solves == sorted(solves) and len(set(solves)) == len(solves)
            n_chunks = -(-(-(-T // sf)) // (chunk // sf)) if T else 0
            assert len(chunk_solve) == n_chunks
            if T:
                assert solves[-1] == T and chunk_solve == sorted(chunk_solve) and chunk_solve[-1] == len(solves) - 1
                for c, j in enumerate(chunk_solve[:-1]):
                    # chunk c needs input frames < (c + 1) * chunk + right; while streaming the iVector lags by the splice
                    assert solves[j] <= T
            else:
                assert solves == [] and chunk_solve == []
