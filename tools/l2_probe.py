"""How much of a just-read buffer does the L2 still hold?  Time a read-only pass over n MB repeated back to back."""
import torch
x = torch.randn(160 * 1024 * 1024 // 2, device="cuda").to(torch.bfloat16)   # 160 MB... bf16 -> 2 bytes/elem
x = torch.empty(256 * 1024 * 1024 // 2, device="cuda", dtype=torch.bfloat16).normal_()
def t(n_mb, reps=10):
    v = x[: n_mb * 1024 * 1024 // 2]
    for _ in range(3): torch.sum(v, dtype=torch.float32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): torch.sum(v, dtype=torch.float32)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n in (8, 16, 32, 48, 64, 96, 128, 192, 256):
    us = t(n)
    print(f"{n:4d} MB repeated read: {us:7.1f} us  -> {n*1.048576/us*1e3:7.0f} GB/s")
