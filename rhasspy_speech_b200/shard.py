"""Multi-GPU sharding of utterance batches.

Utterances are independent (the reference carries no state between them: one process and
spk == utt per call, transcribe_wav.py:58), so the path shards with NO data-path collective:
every rank holds a replica of the model and graph and decodes its own utterances.  The only
communication is the gather of the (tiny) results to the rank that owns the request, done with
``torch.distributed`` object collectives (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence


def shard_utterances(durations: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-first greedy assignment: every rank gets (nearly) equal audio seconds.

    Returns, per rank, the indices of the utterances it decodes (in input order)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(durations)), key=lambda k: (-durations[k], k)):
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += durations[i]
    for r in range(world_size):
        out[r].sort()
    return out


def gather_results(local_indices: Sequence[int], local_results: Sequence, n_total: int, dst: int = 0) -> Optional[list]:
    """Gather per-utterance results to rank `dst` in the original order (None on other ranks)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = [None] * n_total
        for i, r in zip(local_indices, local_results):
            out[i] = r
        return out
    payload = list(zip(local_indices, local_results))
    gathered = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None
    out = [None] * n_total
    for part in gathered:
        for i, r in part:
            out[i] = r
    return out
