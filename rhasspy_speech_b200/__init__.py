"""B200-native batched utterance decoder behind rhasspy-speech's transcribe surface.

Only the hot path lives here: ``csrc/`` holds the CUDA kernels and the C ABI
(include/rs_b200.h), ``_lib`` the ctypes binding, ``transcribe`` the mirror of the reference's
transcriber classes, ``shard`` the multi-GPU utterance sharder.  The seeded fixture generator used by the
tests and the bench is ``tools/synth.py`` at the repository root, outside the package.
"""
from .transcribe import (KaldiNnet3StreamTranscriber, KaldiNnet3WavTranscriber, KaldiTranscriber,  # noqa: F401
                         decode_meta)

__all__ = ["KaldiNnet3WavTranscriber", "KaldiNnet3StreamTranscriber", "KaldiTranscriber", "decode_meta"]
