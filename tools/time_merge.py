"""Quick device timing of one merge-stage call (development tool; bench.py is the judged entry point)."""
import argparse
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import synth
from framefusion_b200.main import FrameFusion


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C2")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--frame", type=int, default=1, help="0: multi-kernel path for call #0 as well, 2: the frame-pipelined kernel forced")
    ap.add_argument("--calls", type=int, default=1)
    a = ap.parse_args()
    c = synth.CONFIGS[a.cfg]
    wl = synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0)
    wl = synth.to_device(wl, "cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times = []
    ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
    ff.use_frame = "force" if a.frame == 2 else bool(a.frame)
    ker = []
    for it in range(a.iters + 3):
        ff.prepare(*wl.prepare_args())
        ff.kernel_events = []
        pos = [wl.cos, wl.sin]
        h = wl.hidden
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.calls):
            if ff.finish_merging:
                break
            h, pos, _m = ff(h, pos, None)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if it >= 3:
            times.append((e0.elapsed_time(e1), (t1 - t0) * 1e3))
            ker.append(ff.kernel_events[0][2].elapsed_time(ff.kernel_events[0][3]))
    dev = sorted(t[0] for t in times)[len(times) // 2]
    wall = sorted(t[1] for t in times)[len(times) // 2]
    s_keep = h.shape[1]
    nbytes = synth.algorithmic_bytes(wl.seq_len, s_keep, c["hidden"], wl.hidden.element_size())
    k_us = sorted(ker)[len(ker) // 2] * 1e3
    print(f"{a.cfg} frame={a.frame}: S={wl.seq_len} -> {s_keep}  ff_merge_layer {k_us:.1f} us = {nbytes/k_us/1e3:.0f} GB/s | whole call: device {dev*1e3:.1f} us  wall {wall*1e3:.1f} us  "
          f"alg {nbytes/1e6:.1f} MB -> {nbytes/dev/1e6:.0f} GB/s  {wl.n_vision/dev*1e3:.3e} vision tok/s")


if __name__ == "__main__":
    main()
