"""(needs a library built with tracing: tools/build_variant.sh trace -DFR_TRACE=1; FF_LIB_PATH=framefusion_b200/variants/libff_trace.so)
Time stamps of one frame-pipelined merge launch (ff_debug_frame_trace): where a frame's time goes.
Development tool; bench.py is the judged entry point."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import _lib, synth
from framefusion_b200.main import FrameFusion

NAMES = ["load requested", "similarity done", "destinations known", "rows out", "aux rows out", "rows in shared memory"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="C2")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    c = synth.CONFIGS[a.cfg]
    wl = synth.to_device(synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0), "cuda")
    ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
    frames, K = c["frames"], 8
    n_cta = -(-c["patch_num"] // -(-c["patch_num"] // 148))           # the library's grid: ceil(P / ceil(P / #SMs))
    buf = torch.zeros(n_cta * frames * K, dtype=torch.int64, device="cuda")
    for it in range(3):
        ff.prepare(*wl.prepare_args())
        st = ff._state(wl.hidden.device)
        if it == 2:
            _lib.check(st.lib.ff_debug_frame_trace(st.ctx, buf.data_ptr(), buf.numel() * 8))
        h, pos, _m = ff(wl.hidden, [wl.cos, wl.sin], None)
        torch.cuda.synchronize()
    raw = buf.cpu().numpy().reshape(-1)[: n_cta * frames * K].reshape(n_cta, frames, K).astype(np.float64)
    notes = raw[:, :6, 7].copy()                            # cycle counts, not time stamps
    raw[:, :, 7] = 0
    t = raw
    t0 = t[:, :, 0][t[:, :, 0] > 0].min()
    t = np.where(t > 0, (t - t0) / 1e3, np.nan)            # us
    print(f"{a.cfg}: {n_cta} CTAs x {frames} frames; us since the first load request; median / min / max over CTAs")
    for f in list(range(0, min(frames, 9))) + list(range(12, frames, max(1, frames // 8))) + [frames - 1]:
        row = "  ".join(f"{NAMES[k][:7]:>7s} {np.nanmedian(t[:, f, k]):6.2f} [{np.nanmin(t[:, f, k]):6.2f},{np.nanmax(t[:, f, k]):6.2f}]" for k in (0, 5, 1, 2, 3, 4))
        print(f"f={f:3d}  {row}")
    print(f"kernel under way (after pdl_wait + init): median {np.nanmedian(t[:, 0, 6]):.2f} [{np.nanmin(t[:, 0, 6]):.2f}, {np.nanmax(t[:, 0, 6]):.2f}]")
    print(f"CTA's chains through: median {np.nanmedian(t[:, 1, 6]):.2f} [{np.nanmin(t[:, 1, 6]):.2f}, {np.nanmax(t[:, 1, 6]):.2f}]")
    print(f"grid through: median {np.nanmedian(t[:, 2, 6]):.2f} [{np.nanmin(t[:, 2, 6]):.2f}, {np.nanmax(t[:, 2, 6]):.2f}];  status block out: {np.nanmax(t[:, 3, 6]):.2f}")
    mhz = 1.92e3
    print(f"cycles of S warp 0 / frame: busy {np.median(notes[:, 0]) / frames:.0f} ({np.median(notes[:, 0]) / frames / mhz:.2f} us at 1.92 GHz), waiting {np.median(notes[:, 1]) / frames:.0f}; "
          f"[S: loop {np.median(notes[:, 4]) / frames:.0f}, the two warp reductions {np.median(notes[:, 5]) / frames:.0f}, rounding chain + flag + global writes {(np.median(notes[:, 0]) - np.median(notes[:, 4]) - np.median(notes[:, 5])) / frames:.0f}] "
          f"G warp 0 / frame: busy {np.median(notes[:, 2]) / frames:.0f} ({np.median(notes[:, 2]) / frames / mhz:.2f} us), waiting {np.median(notes[:, 3]) / frames:.0f}")
    d = np.diff(np.nanmedian(t[:, :, 3], axis=0))
    print(f"rows-out period per frame: median {np.nanmedian(d):.2f} us, mean {np.nanmean(d):.2f} us; span {np.nanmax(t):.1f} us")
    for k0, k1 in ((0, 5), (5, 1), (1, 2), (2, 3), (0, 3), (3, 4)):
        dd = t[:, :, k1] - t[:, :, k0]
        print(f"{NAMES[k0]} -> {NAMES[k1]}: median {np.nanmedian(dd):.2f}  p95 {np.nanpercentile(dd, 95):.2f}")
    if a.out:
        np.save(a.out, t)


if __name__ == "__main__":
    main()
