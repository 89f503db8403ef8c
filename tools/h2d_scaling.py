"""Host <-> device copy rates of the bench's end-to-end step (321 MB in, 74 MB out per step, pinned buffers), per rank
ALONE and with all ranks TOGETHER — what bounds `e2e` at N GPUs (profiles/r02_scaling.md).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_scaling.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (pin_rank_to_gpu_cpus)

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
aff = bench.pin_rank_to_gpu_cpus(local) if os.environ.get("FF_PIN", "1") == "1" else None
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_in, n_out = 321 << 20, 74 << 20
h_in, h_out = torch.empty(n_in, dtype=torch.uint8).pin_memory(), torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n_in, dtype=torch.uint8, device=dev), torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def step():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def measure(iters=8):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


def barrier():
    if world > 1:
        dist.barrier()


alone = None
for r in range(world):
    barrier()
    if r == rank:
        alone = measure()
barrier()
together = measure()
barrier()
res = {"rank": rank, "cpus": sorted(os.sched_getaffinity(0))[:4] + ["..."] if aff else "unpinned",
       "alone_ms": alone * 1e3, "alone_GBs": (n_in + n_out) / alone / 1e9,
       "together_ms": together * 1e3, "together_GBs": (n_in + n_out) / together / 1e9}
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    for o in out:
        print(json.dumps(o))
    print(json.dumps({"world": world, "sum_alone_GBs": sum(o["alone_GBs"] for o in out), "sum_together_GBs": sum(o["together_GBs"] for o in out),
                      "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
