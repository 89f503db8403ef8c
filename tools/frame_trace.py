"""Time stamps of one frame-pipelined merge launch (ff_debug_frame_trace): where a frame's time goes.
Development tool; bench.py is the judged entry point."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import _lib, synth
from framefusion_b200.main import FrameFusion

NAMES = ["load requested", "similarity done", "destinations known", "rows out", "aux rows out"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C2")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    c = synth.CONFIGS[a.cfg]
    wl = synth.to_device(synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0), "cuda")
    ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
    grid, frames, K = 148, c["frames"], 8
    buf = torch.zeros(grid * frames * K, dtype=torch.int64, device="cuda")
    for it in range(3):
        ff.prepare(*wl.prepare_args())
        st = ff._state(wl.hidden.device)
        if it == 2:
            _lib.check(st.lib.ff_debug_frame_trace(st.ctx, buf.data_ptr(), buf.numel() * 8))
        h, pos, _m = ff(wl.hidden, [wl.cos, wl.sin], None)
        torch.cuda.synchronize()
    n_cta = -(-c["patch_num"] // -(-c["patch_num"] // 148))
    t = buf.cpu().numpy().reshape(-1)[: n_cta * frames * K].reshape(n_cta, frames, K).astype(np.float64)
    t0 = t[:, :, 0][t[:, :, 0] > 0].min()
    t = np.where(t > 0, (t - t0) / 1e3, np.nan)            # us
    print(f"{a.cfg}: {n_cta} CTAs x {frames} frames; us since the first load request; median / min / max over CTAs")
    for f in list(range(0, min(frames, 12))) + list(range(12, frames, max(1, frames // 12))) + [frames - 1]:
        row = "  ".join(f"{NAMES[k][:10]:>10s} {np.nanmedian(t[:, f, k]):7.2f} [{np.nanmin(t[:, f, k]):6.2f},{np.nanmax(t[:, f, k]):6.2f}]" for k in range(5))
        print(f"f={f:3d}  {row}")
    d = np.diff(np.nanmedian(t[:, :, 3], axis=0))
    print(f"rows-out period per frame: median {np.nanmedian(d):.2f} us, mean {np.nanmean(d):.2f} us; span {np.nanmax(t):.1f} us")
    for k0, k1 in ((0, 1), (1, 2), (2, 3), (0, 3)):
        dd = t[:, :, k1] - t[:, :, k0]
        print(f"{NAMES[k0]} -> {NAMES[k1]}: median {np.nanmedian(dd):.2f}  p95 {np.nanpercentile(dd, 95):.2f}")
    if a.out:
        np.save(a.out, t)


if __name__ == "__main__":
    main()
