"""Development aid: per-tile time stamps of the read-once merge kernel (a separate library built with -DFF_FUSED_TRACE).

    python tools/trace_fused.py [--cfg C2]      -> gpurun_out/trace_fused_<cfg>.npy + a summary on stdout
Stamps per tile (ns, globaltimer): 0 iteration start, 1 rows of warp 0 arrived, 2 barrier (A) passed, 3 count posted and
scan item pushed, 4 predecessor state known (warp 0), 5 iteration end (warp 0), 6 scan warp dequeued the tile, 7 its
look-back resolved (kept states published), 8 links written, 9 / 10 / 11 latest warp: predecessor state known / rows
arrived / iteration end."""
import argparse, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser(); ap.add_argument("--cfg", default="C2"); a = ap.parse_args()
lib = os.path.join(ROOT, "gpurun_out", "libff_trace.so")
os.makedirs(os.path.dirname(lib), exist_ok=True)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DFF_FUSED_TRACE", "-Xcompiler", "-fPIC",
                       "-shared", "-I", os.path.join(ROOT, "include"), "-o", lib, os.path.join(ROOT, "framefusion_b200", "csrc", "ff_api.cu")])
os.environ["FF_LIB_PATH"] = lib
out = os.path.join(ROOT, "gpurun_out", f"trace_fused_{a.cfg}.bin")
import numpy as np, torch
from framefusion_b200 import synth
from framefusion_b200.main import FrameFusion
c = synth.CONFIGS[a.cfg]
wl = synth.to_device(synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0), "cuda")
ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
for it in range(4):
    if it == 3: os.environ["FF_FUSED_TRACE_FILE"] = out
    ff.prepare(*wl.prepare_args())
    ff(wl.hidden, [wl.cos, wl.sin], None)
    torch.cuda.synchronize()
t = np.fromfile(out, dtype=np.int64).reshape(-1, 16).astype(np.float64)
np.save(out[:-4] + ".npy", t)
t0 = t[:, 0][t[:, 0] > 0].min()
us = lambda x: (x - t0) / 1e3
n = t.shape[0]
print(f"{n} tiles; kernel span {us(t[:, [5, 8, 11]].max()):.1f} us")
d = lambda i, j: np.median((t[:, i] - t[:, j])[(t[:, i] > 0) & (t[:, j] > 0)]) / 1e3
p90 = lambda i, j: np.percentile((t[:, i] - t[:, j])[(t[:, i] > 0) & (t[:, j] > 0)], 90) / 1e3
for name, i, j in [("rows arrive (warp 0)", 1, 0), ("rows arrive (last warp)", 10, 0), ("barrier A after last rows", 2, 10), ("post+push", 3, 2),
                   ("pred state known, warp 0, after A", 4, 2), ("pred state known, last warp, after A", 9, 2), ("iteration (warp 0)", 5, 0),
                   ("iteration (last warp)", 11, 0), ("scan dequeue after push", 6, 3), ("look-back", 7, 6), ("scan tail (pushes)", 8, 7),
                   ("publish after iteration start", 7, 0)]:
    print(f"  {name:40s} median {d(i, j):7.2f} us   p90 {p90(i, j):7.2f} us")
print(f"  deferred emission steps: {int(t[:, 12].sum())} of {n * 8} rows")
for q in (0.1, 0.25, 0.5, 0.75, 1.0):
    k = min(int(q * n), n - 1)
    print(f"  tile {k:5d}: start {us(t[k, 0]):7.1f}  published {us(t[k, 7]):7.1f}  end {us(t[k, 11]):7.1f}")
