"""Development aid: run one golden case through the CUDA path and print where the output departs from the oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from _harness import load_case, build_inputs, t2f, DT
from oracle import ff_oracle as orc
from framefusion_b200.main import FrameFusion
name = sys.argv[1]
z, spec = load_case(name)
wl, pos, mask, dtype = build_inputs(spec)
ff = FrameFusion(spec["cost"], spec["slb"], spec["rlb"]); ff.debug_trace = True
args = list(wl.prepare_args()); args[0] = args[0].cuda(); ff.prepare(*args)
o = orc.OracleFrameFusion(spec["cost"], spec["slb"], spec["rlb"], dtype)
o.prepare(wl.patch_type.numpy(), wl.patch_num, *wl.prepare_args()[2:])
oh, op, _ = o.forward(t2f(wl.hidden[0]), [t2f(pos[0][0]), t2f(pos[1][0])], None)
keep = np.nonzero(o.last["keep_mask"])[0]
n_bad_runs = 0
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    ff = FrameFusion(spec["cost"], spec["slb"], spec["rlb"])
    ff.prepare(*args)
    h = wl.hidden.clone().cuda(); p = [x.cuda() for x in pos]
    out, p2, _ = ff(h, p, None)
    got = t2f(out[0])
    bad = np.nonzero((got != oh).any(axis=1))[0] if got.shape == oh.shape else np.arange(1)
    if len(bad):
        n_bad_runs += 1
        break
print("shapes", got.shape, oh.shape, "kernel", int(ff._state(h.device).status[8]), "bad after reps", rep, n_bad_runs)
print("rows differing:", len(bad), bad[:20])
for r in bad[:6]:
    src = keep[r]
    d = np.nonzero(got[r] != oh[r])[0]
    # is the row equal to some other oracle row (position error) ?
    match = np.nonzero((oh == got[r]).all(axis=1))[0]
    print(f"out row {r} (src seq {src}, patch {wl.patch_type[0, src].item()}): {len(d)} elems differ, first {d[:4]}, got {got[r][d[:3]]} want {oh[r][d[:3]]}; equals oracle rows {match[:4]}")
