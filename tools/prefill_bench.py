#!/usr/bin/env python
"""Prefill of a random-init LLaVA-Video-7B-shape decoder (Qwen2: 28 layers, hidden 3584, 28 heads / 4 KV heads,
MLP 18944, RoPE theta 1e6) on synthetic video embeddings, with the FrameFusion hooks installed — BASELINE config 5.

    python tools/prefill_bench.py [--gpus N] [--impl b200|torch|dense] [--frames 64] [--patch 576] [--layers 28]

One process drives N GPUs of one box (layers split in contiguous blocks, activations hop over NVLink at the block
boundaries — framefusion_b200/dispatch.py), which is how the reference runs it (device_map="auto").  A single
request is a sequential pipeline: more GPUs hold the model, they do not speed the prefill up.
impl: b200 = this repository's operator; torch = the same hooks driving the torch-op restatement of the reference
operator on the GPU (oracle/ff_torch_port.py — baseline infrastructure, the "reference torch path"); dense = no
FrameFusion at all.  Prints one JSON line: prefill tokens/s = original sequence length / prefill time."""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from framefusion_b200 import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--impl", default="b200", choices=["b200", "torch", "dense"])
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--patch", type=int, default=576)
    ap.add_argument("--layers", type=int, default=28)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--cost", type=float, default=0.3)
    a = ap.parse_args()
    from transformers import Qwen2Config, Qwen2ForCausalLM
    cfg = Qwen2Config(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_hidden_layers=a.layers,
                      num_attention_heads=28, num_key_value_heads=4, max_position_embeddings=131072, rope_theta=1e6)
    cfg._attn_implementation = "sdpa"
    devices = [torch.device("cuda", k) for k in range(a.gpus)]
    torch.manual_seed(0)
    with torch.device(devices[0]):
        model = Qwen2ForCausalLM(cfg).to(torch.bfloat16).eval()
    llm = model.model
    if a.impl != "dense":
        from framefusion_b200.interface import apply_framefusion
        apply_framefusion(model, cost=a.cost, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
        if a.impl == "torch":
            import framefusion_b200.hooks.qwen2 as hk
            from oracle import ff_torch_port as port
            class TorchOperator(torch.nn.Module):            # the port behind the attribute surface the hooks read
                def __init__(self):
                    super().__init__()
                    object.__setattr__(self, "p", port.TorchPortFrameFusion(a.cost, 0.6, 0.1))
                    self.p.trace = False
                def prepare(self, *args): self.p.prepare(*args)
                def forward(self, h, pos, mask, attn=None): return self.p(h, pos, mask, attn)
                finish_merging = property(lambda s: s.p.finish_merging)
                finish_pruning = property(lambda s: s.p.finish_pruning)
                sparsity_list = property(lambda s: s.p.sparsity_list)
            op = TorchOperator()
            for m in [model, llm] + list(llm.layers) + [l.self_attn for l in llm.layers]:
                m.framefusion = op
            hk.scaled_dot_product_attention = lambda q, k, v, num=1, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False: \
                port.last_query_attention(q, k, num=num, is_causal=is_causal, scale=scale)
    if a.gpus > 1:
        from framefusion_b200.dispatch import split_layers
        split_layers(llm, devices)
    wl = synth.make_workload(a.frames, a.patch, 3584, torch.bfloat16, seed=0)
    embeds = wl.hidden.to(devices[0])
    pt = wl.patch_type.to(devices[0])
    ff = getattr(model, "framefusion", None)

    def prefill():
        if ff is not None:
            ff.prepare(pt, wl.patch_num, wl.n_pre, wl.n_pre + wl.n_vision - 1, wl.n_vision, wl.seq_len)
        with torch.no_grad():
            return llm(inputs_embeds=embeds, use_cache=True)

    times, kept = [], None
    for it in range(a.iters + 1):
        for d in devices: torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        out = prefill()
        for d in devices: torch.cuda.synchronize(d)
        if it > 0: times.append(time.perf_counter() - t0)
        kept = out.last_hidden_state.shape[1]
        del out
    t = sorted(times)[len(times) // 2]
    print(json.dumps({"metric": "prefill tokens/s, random-init Qwen2-7B-shape decoder with FrameFusion hooks", "impl": a.impl,
                      "n_gpus": a.gpus, "layers": a.layers, "seq_len": wl.seq_len, "tokens_after_last_layer": kept,
                      "prefill_ms": t * 1e3, "value": wl.seq_len / t, "unit": "tokens/s",
                      "sparsity_list": getattr(ff, "sparsity_list", None) if ff is not None else None}))


if __name__ == "__main__":
    main()
