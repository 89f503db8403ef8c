"""CPU, world_size 2 over gloo: the N > 1 logic of bench.py — independent replicas, a barrier on both sides of the
timed region, the MAX over ranks of the elapsed time, whole-job throughput on rank 0 (replicas only,
no data-path collective)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local_ms = 10.0 * (rank + 1)                        # rank 1 is the slow replica
        ms = bench.max_over_ranks(local_ms, dist, torch.device("cpu"))
        value = bench.whole_job_throughput(n_tokens=1000, steps=4, ms=ms, world=world)
        q.put((rank, ms, value))
    finally:
        dist.destroy_process_group()


def test_replicas_max_over_ranks_and_aggregate_value():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, value in res:
        assert ms == 20.0                                   # the slowest rank sets the time
        assert value == 2 * 1000 * 4 / 20e-3                # all ranks' tokens over that time
