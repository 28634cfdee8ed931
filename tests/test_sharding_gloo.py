"""N>1 host logic on CPU: world_size-2 gloo run of the stream sharding and the throughput reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdrainer_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_streams(n_streams, rank, world)
        # every rank "processes" its streams: 100 blocks x 2048 samples each, rank 1 is the slow one
        units = len(mine) * 100 * 2048
        seconds = 0.010 * (1 + rank)
        dist.barrier()
        total_units, max_seconds = sharding.aggregate(dist, torch, units, seconds)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            q.put((gathered, total_units, max_seconds))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_and_reduction():
    world, n_streams = 2, 65
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, total_units, max_seconds = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    flat = sorted(s for part in gathered for s in part)
    assert flat == list(range(n_streams))                       # every stream exactly once
    assert all(sharding.owner_of(s, world) == r for r, part in enumerate(gathered) for s in part)
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1        # balanced
    assert total_units == n_streams * 100 * 2048                # whole-job units
    assert abs(max_seconds - 0.020) < 1e-12                     # time = max over ranks


def test_shard_edge_cases():
    assert sharding.shard_streams(0, 0, 4) == []
    assert sharding.shard_streams(3, 3, 4) == []                # more ranks than streams: idle rank
    assert sharding.shard_streams(8, 1, 8) == [1]
    with pytest.raises(ValueError):
        sharding.shard_streams(8, 8, 8)
    assert sharding.aggregate(None, torch, 5, 2.0) == (5.0, 2.0)
