"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def synth():
    from tools import synth as s
    return s


@pytest.fixture(scope="session")
def tiny_model(tmp_path_factory, synth):
    """Tiny zamia-shaped model + grammar HCLG (seeded)."""
    return synth.write_model(str(tmp_path_factory.mktemp("tiny")), synth.TINY)


@pytest.fixture(scope="session")
def utterances(synth):
    """Seeded speech-like utterances, 1-3 s, plus the edge cases the reference handles."""
    return synth.make_utterances(6, seed=42, min_s=1.0, max_s=3.0)


def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
