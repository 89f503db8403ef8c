"""Yardsticks on this GPU: contiguous copy, row gather (index_select) of the C2 shapes, plain read (sum)."""
import torch
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
S, H, K = 36898, 3584, 22290
x = torch.randn(S, H, device="cuda").to(torch.bfloat16)
out = torch.empty(K, H, device="cuda", dtype=torch.bfloat16)
idx = torch.sort(torch.randperm(S, device="cuda")[:K]).values
big = torch.empty(S, H, device="cuda", dtype=torch.bfloat16)
us = t(lambda: big.copy_(x)); print(f"contiguous copy 264 MB: {us:.1f} us = {2*x.numel()*2/us/1e3:.0f} GB/s (R+W)")
us = t(lambda: out.copy_(x[:K])); print(f"contiguous copy 160 MB: {us:.1f} us = {2*out.numel()*2/us/1e3:.0f} GB/s (R+W)")
us = t(lambda: torch.index_select(x, 0, idx, out=out)); print(f"index_select 22290 rows: {us:.1f} us = {2*out.numel()*2/us/1e3:.0f} GB/s (R+W)")
us = t(lambda: torch.sum(x, dtype=torch.float32)); print(f"sum (read 264 MB): {us:.1f} us = {x.numel()*2/us/1e3:.0f} GB/s")
