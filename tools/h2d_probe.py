import torch, time
n = 321 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def two():
    m = n // 2
    with torch.cuda.stream(s1): d[:m].copy_(h[:m], non_blocking=True)
    with torch.cuda.stream(s2): d[m:].copy_(h[m:], non_blocking=True)
def four():
    m = n // 4
    for k, s in enumerate((s1, s2, s1, s2)):
        with torch.cuda.stream(s): d[k*m:(k+1)*m].copy_(h[k*m:(k+1)*m], non_blocking=True)
for name, fn in (("one stream", one), ("two streams", two), ("four chunks / two streams", four)):
    for _ in range(2): fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {dt*1e3:.2f} ms  {n/dt/1e9:.1f} GB/s")
