"""Runs merge call #0 through the frame-pipelined kernel many times and compares every output with the multi-kernel
path on the same input: S_keep, hidden_states, cos / sin, patch_type, the keep mask and the similarities.
Development tool (a kernel whose roles hand over through shared-memory counters has to be shown free of races)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import synth
from framefusion_b200.main import FrameFusion


def run(ff, wl):
    ff.prepare(*wl.prepare_args())
    h, pos, _m = ff(wl.hidden, [wl.cos, wl.sin], None)
    torch.cuda.synchronize()
    return h, pos, ff.patch_type, ff.last_trace


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C2")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--frames", type=int, default=0)
    a = ap.parse_args()
    c = dict(synth.CONFIGS[a.cfg])
    if a.frames:
        c["frames"] = a.frames
    wl = synth.to_device(synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0), "cuda")
    ref = FrameFusion(c["cost"], c["slb"], c["rlb"])
    ref.use_frame = False
    ref.debug_trace = True
    h0, pos0, pt0, tr0 = run(ref, wl)
    ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
    ff.debug_trace = True
    bad = 0
    P, F = c["patch_num"], c["frames"]
    for it in range(a.iters):
        h, pos, pt, tr = run(ff, wl)
        ok = h.shape == h0.shape and torch.equal(h, h0) and torch.equal(pos[0], pos0[0]) and torch.equal(pos[1], pos0[1]) and torch.equal(pt, pt0)
        if not ok:
            bad += 1
            km, km0 = tr["keep_mask"], tr0["keep_mask"]
            d = np.nonzero(km != km0)[0]
            first = int(np.nonzero(wl.patch_type.reshape(-1).cpu().numpy() >= 0)[0][0])
            where = [((int(r) - first) // P, (int(r) - first) % P) for r in d[:8]]
            sd = np.nonzero(tr["sim_values"] != tr0["sim_values"])[0]
            nrow = 0
            if h.shape == h0.shape:
                nrow = int((h != h0).any(dim=-1).sum())
            print(f"iter {it}: S_keep {h.shape[1]} vs {h0.shape[1]}; keep mask differs at {len(d)} rows (frame, patch) {where}; "
                  f"sim differs at {len(sd)} by-patch positions {[(int(j) // F, int(j) % F, float(tr['sim_values'][j]), float(tr0['sim_values'][j])) for j in sd[:6]]} (patch, frame, got, want); "
                  f"{nrow} output rows differ", flush=True)
    print(f"{a.cfg} x{a.iters}: {bad} runs differ from the multi-kernel path (S_keep {h0.shape[1]})")


if __name__ == "__main__":
    main()
