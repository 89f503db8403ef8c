"""Development aid: per-tile time stamps of the read-once merge kernel (a separate library built with -DFF_FUSED_TRACE).

    python tools/trace_fused.py [--cfg C2]      -> gpurun_out/trace_fused_<cfg>.npy + a summary on stdout
Stamps per tile (ns, globaltimer): 0 tile taken, 10 / 9 / 11 latest warp: rows arrived / predecessor flag known / front
step done, 6 scan warp has the flags, 7 look-back resolved and destinations published, 8 tile handed to the workers."""
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
print(f"{n} tiles; kernel span {us(t[:, [8, 11, 15]].max()):.1f} us")
def dd(i, j):
    m = (t[:, i] > 0) & (t[:, j] > 0)
    if not m.any():
        return "no samples"
    x = (t[:, i] - t[:, j])[m] / 1e3
    return f"median {np.median(x):7.2f} us   p90 {np.percentile(x, 90):7.2f} us   max {x.max():7.2f} us"
# stamps: 0 tile taken (thread 0), 10 rows arrived (latest warp), 9 predecessor flag known (latest warp), 11 front step done
# (latest warp), 6 scan warp has the tile's flags, 7 look-back resolved / destinations published, 8 tile handed to the workers
for name, i, j in [("rows arrive (latest warp)", 10, 0), ("predecessor flag known (latest warp)", 9, 0), ("front step done (latest warp)", 11, 0),
                   ("scan has the flags, after front done", 6, 11), ("look-back", 7, 6), ("hand-over to workers", 8, 7),
                   ("destinations after tile taken", 7, 0), ("worker takes the tile, after hand-over", 12, 8),
                   ("worker: walks and destinations", 3, 12), ("worker: copies, aux rows, runs", 15, 3)]:
    print(f"  {name:40s} {dd(i, j)}")
print("  last tile warp, one iteration:")
for name, i, j in [("tile known after its ticket was taken", 2, 0), ("link known, TMA issued", 4, 2), ("rows arrived", 5, 4), ("similarity done", 14, 5)]:
    print(f"    {name:38s} {dd(i, j)}")
# per CTA: iteration period, and the wait between one tile's end and the next tile's start (last tile warp)
cta = t[:, 13].astype(int)
per, gap = [], []
for c in np.unique(cta[cta > 0])[:64]:
    m = cta == c
    o = np.argsort(t[m, 2])
    per += list(np.diff(t[m, 2][o]) / 1e3)
    gap += list((t[m, 2][o][1:] - t[m, 14][o][:-1]) / 1e3)
print(f"    {'wait for the next tile':38s} median {np.median(gap):7.2f} us   p90 {np.percentile(gap, 90):7.2f} us")
print(f"  iteration period of a CTA              median {np.median(per):7.2f} us   p90 {np.percentile(per, 90):7.2f} us")
st = np.sort(t[:, 0][t[:, 0] > 0])
print(f"  tiles taken per us (middle half): {(len(st) // 2) / ((st[3 * len(st) // 4] - st[len(st) // 4]) / 1e3):.1f}")
for q in (0.1, 0.25, 0.5, 0.75, 1.0):
    k = min(int(q * n), n - 1)
    print(f"  tile {k:5d}: taken {us(t[k, 0]):7.1f}  front done {us(t[k, 11]):7.1f}  destinations {us(t[k, 7]):7.1f}  handed over {us(t[k, 8]):7.1f}")
