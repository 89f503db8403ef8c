"""Development aid: per-chain, per-row time stamps of the single-pass kernel -> gpurun_out/trace_<tag>.npy
stamps: 0 TMA issued, 1 sim task starts, 2 row arrived, 3 flag published, 4 gap counted, 5 merge starts, 6 merge done"""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import synth, _lib
from framefusion_b200.main import FrameFusion

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="C2")
ap.add_argument("--tag", default="x")
a = ap.parse_args()
c = synth.CONFIGS[a.cfg]
wl = synth.to_device(synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0), "cuda")
ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
for it in range(3):
    ff.prepare(*wl.prepare_args())
    ff(wl.hidden, [wl.cos, wl.sin], None)
st = ff._state(wl.hidden.device)
n_ids = c["patch_num"]
buf = torch.zeros(n_ids * 96 * 8, dtype=torch.int64, device="cuda")
_lib.check(st.lib.ff_debug_trace(st.ctx, buf.data_ptr(), buf.numel() * 8, n_ids))
ff.prepare(*wl.prepare_args())
ff.kernel_events = []
ff(wl.hidden, [wl.cos, wl.sin], None)
torch.cuda.synchronize()
print("ff_merge_layer us", ff.kernel_events[0][2].elapsed_time(ff.kernel_events[0][3]) * 1e3)
_lib.check(st.lib.ff_debug_trace(st.ctx, None, 0, 0))
tr = buf.cpu().numpy().reshape(n_ids, 96, 8)
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/trace_{a.tag}.npy", tr)
t0 = tr[tr > 0].min()
print("kernel span us", (tr.max() - t0) / 1e3)
