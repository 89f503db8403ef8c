"""Development aid: where the host time of one bench step goes (cProfile over 200 steps)."""
import cProfile, pstats, sys, os, io, time
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
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100): step()
torch.cuda.synchronize()
print("step wall us", (time.perf_counter() - t0) / 100 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
