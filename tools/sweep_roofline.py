"""Roofline sweep of the merge stage (BASELINE config 3): frames x patch layouts, one merge call each.
Prints one line per point: algorithmic GB/s of ff_merge_layer (CUDA events) for the frame-pipelined kernel (the default for
the first call of a uniform video), and the multi-kernel path (profiles/r02_sweep.jsonl also holds the read-once kernel of r02, since removed)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import synth
from framefusion_b200.main import FrameFusion

def measure(frames, patch, hidden, cost, mode, iters=7):
    wl = synth.to_device(synth.make_workload(frames, patch, hidden, torch.bfloat16, seed=0), "cuda")
    ff = FrameFusion(cost, 0.6, 0.1); ff.use_frame = "force" if mode == "frame" else False
    ts = []
    for it in range(iters):
        ff.prepare(*wl.prepare_args()); ff.kernel_events = []
        h, _p, _m = ff(wl.hidden, [wl.cos, wl.sin], None)
        torch.cuda.synchronize()
        ts.append(ff.kernel_events[0][2].elapsed_time(ff.kernel_events[0][3]))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    nbytes = synth.algorithmic_bytes(wl.seq_len, h.shape[1], hidden, 2)
    return dict(frames=frames, patch_num=patch, hidden=hidden, seq_len=wl.seq_len, kept=h.shape[1], path=mode,
                us=round(ms * 1e3, 1), alg_mb=round(nbytes / 1e6, 1), alg_gbs=round(nbytes / ms / 1e6))

for (patch, hidden, cost) in ((576, 3584, 0.5), (729, 4096, 0.3), (210, 3584, 0.3)):
    for frames in (16, 32, 64, 128, 256):
        if frames * patch * hidden * 2 > 1.3e9: continue
        for mode in ("frame", "multi-kernel"):
            print(json.dumps(measure(frames, patch, hidden, cost, mode)), flush=True)
