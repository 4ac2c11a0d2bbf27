"""The N>1 path on CPU: world_size-2 gloo processes shard a batch and gather results to rank 0."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rhasspy_speech_b200.shard import gather_results, shard_utterances
    durations = [float(1 + (i * 7) % 5) for i in range(37)]
    mine = shard_utterances(durations, world)[rank]
    # stand-in for the per-rank decode: a deterministic function of the utterance index
    results = [[i, i * i] for i in mine]
    out = gather_results(mine, results, len(durations), dst=0)
    # timing reduction as bench.py does it: max over ranks
    t = torch.tensor([0.5 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((out, float(t)))
    else:
        assert out is None
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, tmax = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert out == [[i, i * i] for i in range(37)]
    assert tmax == 1.5
