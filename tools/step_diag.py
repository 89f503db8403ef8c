"""Development aid: what the measurement hooks of bench.py cost a device-resident step (20-step regions, CUDA events)."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import synth
from framefusion_b200.main import FrameFusion
from framefusion_b200.utils import scaled_dot_product_attention
import bench
c = synth.CONFIGS["C2"]
wl = synth.to_device(synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0), "cuda")
q, k = synth.make_attention_inputs(wl.seq_len, 28, 4, 128, c["dtype"], seed=0)
q_last, k = q[:, :, -1:, :].contiguous().cuda(), k.cuda()
ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
imp = lambda qq, kk: scaled_dot_product_attention(qq, kk, None, num=1, is_causal=True, enable_gqa=True)
step = lambda: bench.run_step(ff, wl, wl.hidden, wl.cos, wl.sin, q_last, k, imp)
ff.reserve_kernel_events(64)


def region(n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def med(label, setup=None, reps=7, ctx=None):
    ts = []
    for _ in range(reps):
        if setup:
            setup()
        if ctx:
            with ctx():
                ts.append(region())
        else:
            ts.append(region())
    ff.kernel_events = None
    ff.kernel_events_len = None
    print(f"{label:40s} median {statistics.median(ts):7.1f} us  min {min(ts):7.1f}  max {max(ts):7.1f}", flush=True)


for _ in range(5):
    step()
med("plain")
def ev_all():
    ff.kernel_events = []; ff.kernel_events_len = None
def ev_first():
    ff.kernel_events = []; ff.kernel_events_len = wl.seq_len
med("kernel events, every merge call", ev_all)
med("kernel events, call #0 only", ev_first)
med("clock sampler (nvml thread)", None, ctx=lambda: bench.ClockSampler(0))
med("plain again")
for n in (5, 20, 100):
    print(n, "steps per region:", " ".join(f"{region(n):.1f}" for _ in range(4)), flush=True)
