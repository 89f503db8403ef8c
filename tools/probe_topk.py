"""Which elements does torch.topk pick on CUDA when the k-th value is tied?"""
import torch
torch.manual_seed(0)
for n, k in ((36864, 25804), (22000, 10393), (4096, 1000)):
    v = torch.randn(n).mul(0.3).add(0.5).to(torch.bfloat16)
    for dev in ("cpu", "cuda"):
        x = v.to(dev)[None]
        idx = torch.topk(x, k).indices[0].sort().values.cpu()
        kth = v[idx].float().min()
        ties = (v.float() == kth).nonzero().flatten()
        sel = torch.isin(ties, idx)
        n_sel = int(sel.sum())
        lowest = bool(sel[:n_sel].all())
        highest = bool(sel[-n_sel:].all()) if n_sel else True
        print(f"n={n} k={k} {dev}: ties at kth={len(ties)} selected={n_sel} lowest-first={lowest} highest-first={highest}")
