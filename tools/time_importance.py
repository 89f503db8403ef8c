"""Device timing of the importance path (last-query attention probabilities) at the C2 prune-stage shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from framefusion_b200 import synth
from framefusion_b200.utils import scaled_dot_product_attention
for S, num in ((22290, 1), (22290, 4), (44354, 1)):
    q, k = synth.make_attention_inputs(S, 28, 4, 128, torch.bfloat16, seed=0)
    q, k = q[:, :, -num:, :].contiguous().cuda(), k.cuda()
    for _ in range(3):
        scaled_dot_product_attention(q, k, None, num=num, is_causal=True, enable_gqa=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        scaled_dot_product_attention(q, k, None, num=num, is_causal=True, enable_gqa=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"S={S} num={num}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per call (logits + softmax, incl. host launch gaps)")
