"""Multi-GPU sharding of the video hot path: independent sessions, one per GPU, no data-path collective.

The reference has no cross-session state (one capture instance per display, selkies.py:3178-3181; fan-out
to viewers happens after encode, webrtc/contrib/media.py:596-667), so the path shards by session id and the
only cross-rank traffic is the benchmark's barrier and its max-over-ranks timing.
"""
from __future__ import annotations


def session_device(session_id: int, n_gpus: int) -> int:
    """Partitioning rule: session i runs on GPU i mod n (settings.py:162 `gpu_id` is the knob)."""
    if n_gpus <= 0:
        raise ValueError("n_gpus must be positive")
    return session_id % n_gpus


def aggregate_throughput(units_this_rank: float, ms_this_rank: float, device=None) -> tuple[float, float]:
    """Whole-job throughput = sum of units over ranks / max time over ranks.  Works with any initialised
    torch.distributed backend (nccl on the GPU box, gloo in the CPU tests); single process if uninitialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return units_this_rank / (ms_this_rank / 1000.0), ms_this_rank
    t = torch.tensor([ms_this_rank], dtype=torch.float64, device=device)
    u = torch.tensor([units_this_rank], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / (float(t.item()) / 1000.0), float(t.item())


def gather_over_ranks(values: dict, device=None) -> dict:
    """{key: [value on rank 0, rank 1, ...]} for a flat dict of numbers (same keys on every rank): the per-rank breakdown
    bench.py prints next to the aggregate, so that a lagging rank names the wait that ate its time."""
    import torch
    import torch.distributed as dist
    keys = sorted(values)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: [float(values[k])] for k in keys}
    world, rank = dist.get_world_size(), dist.get_rank()
    t = torch.zeros(world, len(keys), dtype=torch.float64, device=device)
    t[rank] = torch.tensor([float(values[k]) for k in keys], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: [float(v) for v in t[:, i].tolist()] for i, k in enumerate(keys)}
