"""The N>1 path on CPU: two gloo ranks, whole-job throughput = sum(frames) / max(time)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from selkies_b200.multi_gpu import aggregate_throughput, session_device
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert session_device(rank, world) == rank
        # rank 0: 320 frames in 100 ms, rank 1: 320 frames in 160 ms -> job = 640 / 0.160 s
        fps, ms = aggregate_throughput(320.0, 100.0 if rank == 0 else 160.0)
        dist.barrier()
        q.put((rank, fps, ms))
    finally:
        dist.destroy_process_group()


def test_two_rank_aggregate():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, fps, ms in res:
        assert ms == pytest.approx(160.0)
        assert fps == pytest.approx(4000.0)


def _worker_gather(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from selkies_b200.multi_gpu import gather_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, gather_over_ranks({"wait_event_us": 40.0 + rank, "submit_us": 20.0 * (rank + 1)})))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_per_rank_breakdown():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_gather, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, d in res:
        assert d == {"submit_us": [20.0, 40.0], "wait_event_us": [40.0, 41.0]}
