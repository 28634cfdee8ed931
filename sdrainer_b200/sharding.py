"""Multi-GPU partitioning of the hot path: by independent IQ stream, no data-path collective.

rx.Receiver instances share nothing (rx/receiver.go:64-91), so stream s goes to rank s mod world; every rank
owns its own engine, device and pinned rings.  torch.distributed is used only for the barrier around the timed
region and for reducing the per-rank (units, seconds) pair: whole-job throughput = sum(units) / max(seconds).
"""
from __future__ import annotations


def shard_streams(n_streams: int, rank: int, world: int) -> list[int]:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_streams, world))


def owner_of(stream: int, world: int) -> int:
    return stream % world


def aggregate(dist, torch, local_units: float, local_seconds: float, device="cpu"):
    """Returns (total_units, max_seconds) over all ranks; works with gloo (CPU) and nccl (CUDA)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_units), float(local_seconds)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    s = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return float(u.item()), float(s.item())
