"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE.  Usage (build container only — ``/root/reference`` does not exist on the GPU box)::

    python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden]

The reference operator (``framefusion/main.py``) is loaded by file path and driven through
``prepare`` + repeated ``forward`` calls on seeded synthetic inputs (``framefusion_b200/synth.py``).
A thin subclass taps the values that flow between its static methods (similarity, order, merge
index, keep mask, prune indices) without changing any arithmetic.  ``utils.scaled_dot_product_attention``
cannot be imported here (``utils.py`` needs matplotlib), so its source text is extracted with ``ast``
and executed as is.

Every fixture stores the generator spec, a checksum of the regenerated inputs, and per call the
selection results exactly plus order-sensitive row checksums of the tensors that come out.
"""
from __future__ import annotations

import argparse
import ast
import importlib.util
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from framefusion_b200 import synth  # noqa: E402

DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


# ------------------------------------------------------------------------------------------
# helpers shared with the tests (tests import them from here)
# ------------------------------------------------------------------------------------------
def raw_bits(t: torch.Tensor) -> np.ndarray:
    """Storage bits of a tensor as an unsigned numpy array (uint16 / uint32 / uint64 / uint8)."""
    t = t.detach().cpu().contiguous()
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.view(torch.int16).numpy().view(np.uint16)
    if t.dtype == torch.float32:
        return t.view(torch.int32).numpy().view(np.uint32)
    if t.dtype == torch.int64:
        return t.numpy().view(np.uint64)
    if t.dtype == torch.bool:
        return t.numpy().astype(np.uint8)
    raise TypeError(t.dtype)


def row_checksums(t: torch.Tensor) -> np.ndarray:
    """``[rows, 2]`` uint64: plain and position-weighted sums of the storage bits of each row."""
    b = raw_bits(t)
    b = b.reshape(-1, b.shape[-1]).astype(np.uint64)
    w = np.arange(1, b.shape[1] + 1, dtype=np.uint64)
    return np.stack([b.sum(axis=1), (b * w).sum(axis=1)], axis=1)


def tensor_checksum(t: torch.Tensor) -> np.ndarray:
    c = row_checksums(t.reshape(1, -1) if t.ndim == 1 else t.reshape(-1, t.shape[-1]))
    w = np.arange(1, c.shape[0] + 1, dtype=np.uint64)[:, None]
    return (c * w).sum(axis=0)


def load_reference(ref_dir: str):
    path = os.path.join(ref_dir, "framefusion", "main.py")
    spec = importlib.util.spec_from_file_location("_ref_framefusion_main", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_sdpa(ref_dir: str):
    """Extract ``scaled_dot_product_attention`` from utils.py without importing the module."""
    path = os.path.join(ref_dir, "framefusion", "utils.py")
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "scaled_dot_product_attention":
            code = ast.get_source_segment(src, node)
            ns = {"torch": torch, "math": math}
            exec(compile(code, path, "exec"), ns)
            return ns["scaled_dot_product_attention"]
    raise RuntimeError("function not found")


# ------------------------------------------------------------------------------------------
# fixture cases
# ------------------------------------------------------------------------------------------
def _wl(**kw):
    return kw


CASES = {
    # name: dict(workload kwargs, operator kwargs, calls, position container, mask)
    "C1_seed0": dict(wl=_wl(frames=8, patch_num=196, hidden=1024, dtype="f32", seed=0), cost=0.3, slb=0.6, rlb=0.1, calls=1),
    "C1_seed1": dict(wl=_wl(frames=8, patch_num=196, hidden=1024, dtype="f32", seed=1), cost=0.3, slb=0.6, rlb=0.1, calls=1),
    "C1_seed2_multi": dict(wl=_wl(frames=8, patch_num=196, hidden=1024, dtype="f32", seed=2), cost=0.3, slb=0.6, rlb=0.1, calls=6),
    "C2r_seed0": dict(wl=_wl(frames=12, patch_num=48, hidden=3584, dtype="bf16", seed=0), cost=0.3, slb=0.6, rlb=0.1, calls=1),
    "C2r_seed1_multi": dict(wl=_wl(frames=12, patch_num=48, hidden=3584, dtype="bf16", seed=1), cost=0.3, slb=0.6, rlb=0.1, calls=8),
    "C2r_topk": dict(wl=_wl(frames=12, patch_num=48, hidden=3584, dtype="bf16", seed=2, r_lo=0.8, r_hi=1.0), cost=0.3, slb=0.6, rlb=0.1, calls=2),
    "C2r_lowsim_prune": dict(wl=_wl(frames=12, patch_num=48, hidden=3584, dtype="bf16", seed=3, r_lo=0.0, r_hi=0.5), cost=0.3, slb=0.6, rlb=0.1, calls=3),
    "C2r_longruns": dict(wl=_wl(frames=16, patch_num=40, hidden=3584, dtype="bf16", seed=4, per_patch_r=True), cost=0.5, slb=0.6, rlb=0.1, calls=2),
    "C2r_slb07": dict(wl=_wl(frames=12, patch_num=48, hidden=3584, dtype="bf16", seed=5), cost=0.3, slb=0.7, rlb=0.1, calls=2),
    "C2r_drift": dict(wl=_wl(frames=12, patch_num=48, hidden=3584, dtype="bf16", seed=11), cost=0.3, slb=0.6, rlb=0.1, calls=6, drift=0.7),
    "C1_drift": dict(wl=_wl(frames=8, patch_num=196, hidden=1024, dtype="f32", seed=12), cost=0.3, slb=0.6, rlb=0.1, calls=6, drift=0.7),
    "C3r_seed0": dict(wl=_wl(frames=20, patch_num=36, hidden=3584, dtype="bf16", seed=0), cost=0.5, slb=0.6, rlb=0.1, calls=4),
    "C4r_seed0": dict(wl=_wl(frames=10, patch_num=81, hidden=4096, dtype="bf16", seed=0), cost=0.3, slb=0.6, rlb=0.1, calls=3),
    "C4r_lowsim": dict(wl=_wl(frames=10, patch_num=81, hidden=4096, dtype="bf16", seed=1, r_lo=0.0, r_hi=0.5), cost=0.3, slb=0.6, rlb=0.1, calls=3),
    "f16_seed0": dict(wl=_wl(frames=10, patch_num=30, hidden=1536, dtype="f16", seed=0), cost=0.3, slb=0.6, rlb=0.1, calls=4),
    "mrope4d": dict(wl=_wl(frames=8, patch_num=24, hidden=512, dtype="bf16", seed=6), cost=0.3, slb=0.5, rlb=0.1, calls=4, pos="list4d", num=4),
    "posids2d": dict(wl=_wl(frames=8, patch_num=24, hidden=512, dtype="bf16", seed=7), cost=0.3, slb=0.5, rlb=0.1, calls=4, pos="tensor2d"),
    "mask4d": dict(wl=_wl(frames=6, patch_num=16, hidden=256, dtype="bf16", seed=8, r_lo=0.0, r_hi=0.55), cost=0.3, slb=0.6, rlb=0.15, calls=4, mask=True),
    "oddhidden": dict(wl=_wl(frames=7, patch_num=13, hidden=1000, dtype="bf16", seed=9, n_pre=3, n_post=5), cost=0.4, slb=0.6, rlb=0.1, calls=3),
    # SURVEY H9 corner cases
    "H9_run300": dict(wl=_wl(frames=301, patch_num=4, hidden=64, dtype="bf16", seed=20, r_lo=0.0, r_hi=0.5, frozen_patches=1), cost=0.3, slb=0.6, rlb=0.1, calls=1),
    "H9_zero_thr": dict(wl=_wl(frames=8, patch_num=24, hidden=256, dtype="bf16", seed=21, zero_rows=[[2, 3], [3, 3], [5, 7], [0, 11], [7, 20]]), cost=0.3, slb=0.6, rlb=0.1, calls=2),
    "H9_zero_topk": dict(wl=_wl(frames=8, patch_num=24, hidden=256, dtype="bf16", seed=22, r_lo=0.8, r_hi=1.0, zero_rows=[[2, 3], [3, 3], [5, 7], [0, 11], [7, 20]]), cost=0.3, slb=0.6, rlb=0.1, calls=1),
    "H9_topk_sentinels": dict(wl=_wl(frames=8, patch_num=16, hidden=256, dtype="bf16", seed=23), cost=0.05, slb=-2.0, rlb=0.1, calls=1),
    "floatpatchnum": dict(wl=_wl(frames=6, patch_num=20, hidden=384, dtype="bf16", seed=10), cost=0.3, slb=0.6, rlb=0.1, calls=2, patch_num_float=True),
}


def build_inputs(case: dict):
    kw = dict(case["wl"])
    dtype = kw.pop("dtype")
    wl = synth.make_workload(dtype=DT[dtype], **kw)
    pos_kind = case.get("pos", "list3d")
    if pos_kind == "list3d":
        pos = [wl.cos.clone(), wl.sin.clone()]
    elif pos_kind == "list4d":
        g = torch.Generator().manual_seed(1234 + kw["seed"])
        pos = [torch.randn(3, 1, wl.seq_len, 64, generator=g).to(DT[dtype]) for _ in range(2)]
    elif pos_kind == "tensor2d":
        pos = torch.arange(wl.seq_len, dtype=torch.int64)[None] * 3 + 1
    else:
        raise ValueError(pos_kind)
    mask = None
    if case.get("mask"):
        s = wl.seq_len
        mask = torch.full((s, s), float("-inf")).triu(1).to(DT[dtype])[None, None]
    return wl, pos, mask, dtype


def input_checksum(wl, pos, mask) -> np.ndarray:
    parts = [tensor_checksum(wl.hidden[0]), tensor_checksum(wl.patch_type)]
    for p in (pos if isinstance(pos, list) else [pos]):
        parts.append(tensor_checksum(p.reshape(-1, p.shape[-1])))
    if mask is not None:
        parts.append(tensor_checksum(mask[0, 0]))
    return np.concatenate(parts)


def run_case(ref, name: str, case: dict) -> dict:
    wl, pos, mask, dtype = build_inputs(case)
    num = case.get("num", 1)
    tap = {}

    class Tap(ref.FrameFusion):
        def compute_similarity_and_token_index_by_patch(self, hidden_states, token_patch_type, patch_num):
            sim, order = ref.FrameFusion.compute_similarity_and_token_index_by_patch(hidden_states, token_patch_type, patch_num)
            tap["sim"], tap["order"] = sim.clone(), order.clone()
            return sim, order

        def merge_tokens_and_get_mask(self, hidden_states, sim, order, merge_index):
            tap["merge_index"] = merge_index.clone()
            out, keep = ref.FrameFusion.merge_tokens_and_get_mask(hidden_states, sim, order, merge_index)
            tap["keep_mask"] = keep.clone()
            return out, keep

        def position_embedding_handler_at_pruning(self, position_embeddings, keep_indexs):
            tap["prune_keep"] = keep_indexs.clone()
            return super().position_embedding_handler_at_pruning(position_embeddings, keep_indexs)

    ff = Tap(case["cost"], case["slb"], case["rlb"])
    args = list(wl.prepare_args())
    if case.get("patch_num_float"):
        args[1] = float(args[1])
    ff.prepare(*args)

    out = {"spec": np.array(json.dumps({k: v for k, v in case.items()})), "input_checksum": input_checksum(wl, pos, mask)}
    hidden = wl.hidden.clone()
    n_done = 0
    for c in range(case["calls"]):
        tap.clear()
        attn = None
        if c > 0 and case.get("drift"):
            hidden = synth.apply_drift(hidden, case["drift"], case["wl"]["seed"], c)
        if ff.finish_merging and not ff.finish_pruning:
            attn = synth.make_attention_row(hidden.shape[1], n_heads=28, num=num, dtype=DT[dtype], seed=wl.hidden.shape[1] + c)
        hidden, pos, mask = ff(hidden, pos, mask, attn)
        p = f"c{c}_"
        stage = ("prune" if "prune_keep" in tap else "") + ("merge" if "sim" in tap else "")
        out[p + "stage"] = np.array(stage or "none")
        if "sim" in tap:
            out[p + "sim"] = raw_bits(tap["sim"][0])
            out[p + "order"] = tap["order"][0].numpy().astype(np.int32)
            out[p + "merge_index"] = tap["merge_index"].numpy().astype(np.int32)
            out[p + "keep_mask"] = np.packbits(tap["keep_mask"][0].numpy())
        if "prune_keep" in tap:
            out[p + "prune_keep"] = tap["prune_keep"].numpy().astype(np.int32)
            out[p + "attn_checksum"] = tensor_checksum(attn[0].reshape(-1, attn.shape[-1]))
        out[p + "seq_len"] = np.array(hidden.shape[1])
        out[p + "hidden_rows"] = row_checksums(hidden[0])
        plist = pos if isinstance(pos, list) else [pos]
        for i, t in enumerate(plist):
            out[p + f"pos{i}"] = tensor_checksum(t.reshape(-1, t.shape[-1]))
            out[p + f"pos{i}_shape"] = np.array(t.shape)
        if mask is not None:
            out[p + "mask"] = tensor_checksum(mask[0, 0])
        out[p + "patch_type"] = ff.patch_type[0].numpy().astype(np.int32)
        out[p + "flags"] = np.array([ff.finish_merging, ff.finish_pruning])
        out[p + "sparsity_list"] = np.array(ff.sparsity_list, dtype=np.float64)
        n_done += 1
    out["n_calls"] = np.array(n_done)
    return out


def run_statics(ref) -> dict:
    """Direct calls of the reference's helper functions on tiny inputs (incl. its own docstring KAT)."""
    out = {}
    kat_in = torch.tensor([[0, 1, 1, 1, 0, 0, 1, 1]])
    out["runs_kat_in"] = kat_in.numpy()
    out["runs_kat_out"] = ref.find_contigious_latter_index(kat_in).numpy()
    g = torch.Generator().manual_seed(0)
    rnd = (torch.rand(4, 97, generator=g) < 0.55).to(torch.int64)
    out["runs_rand_in"] = rnd.numpy()
    out["runs_rand_out"] = np.stack([ref.find_contigious_latter_index(r[None])[0].numpy() for r in rnd])
    rows = []
    for cost in (0.2, 0.3, 0.5, 0.7, 0.9):
        for sl in ([], [0.39], [0.39, 0.2], [0.5, 0.4, 0.3, 0.05], [0.0, 0.0]):
            try:
                v = ref.FrameFusion._compute_pruning_ratio(sl, cost)
            except ValueError:
                v = float("nan")
            rows.append((cost, len(sl), v))
    out["budget"] = np.array(rows, dtype=np.float64)
    for dtype in ("bf16", "f16", "f32"):
        a = torch.randn(32, 515, generator=g).to(DT[dtype])
        b = (0.7 * a.float() + 0.7 * torch.randn(32, 515, generator=g)).to(DT[dtype])
        out[f"cos_{dtype}_a"] = raw_bits(a)
        out[f"cos_{dtype}_b"] = raw_bits(b)
        out[f"cos_{dtype}_out"] = raw_bits(ref.cosine_similarity(a[None], b[None])[0])
    return out


def run_importance(ref_dir: str) -> dict:
    sdpa = load_reference_sdpa(ref_dir)
    out = {}
    for tag, (s_len, num, causal, gqa, dtype) in {
        "a": (333, 1, True, False, "bf16"),
        "b": (1000, 4, True, False, "bf16"),
        "c": (257, 1, False, False, "bf16"),
        "d": (300, 1, True, True, "bf16"),
        "e": (129, 4, True, False, "f32"),
        "f": (200, 1, True, False, "f16"),
    }.items():
        q, k = synth.make_attention_inputs(s_len, n_heads=28, n_kv_heads=4, head_dim=128, dtype=DT[dtype], seed=s_len)
        kk = k if gqa else k.repeat_interleave(7, dim=1)
        w = sdpa(q, kk, kk, num=num, attn_mask=None, dropout_p=0.0, is_causal=causal, enable_gqa=gqa)
        out[f"imp_{tag}_spec"] = np.array(json.dumps(dict(s_len=s_len, num=num, causal=causal, gqa=gqa, dtype=dtype)))
        out[f"imp_{tag}_out"] = raw_bits(w[0])
        out[f"imp_{tag}_mean"] = raw_bits(torch.mean(w, dim=(1, 2))[0])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("FF_REFERENCE_DIR", "/root/reference"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    ref = load_reference(a.ref)
    os.makedirs(a.out, exist_ok=True)
    if a.only is None:
        np.savez_compressed(os.path.join(a.out, "statics.npz"), **run_statics(ref))
        np.savez_compressed(os.path.join(a.out, "importance.npz"), **run_importance(a.ref))
    for name, case in CASES.items():
        if a.only and a.only != name:
            continue
        res = run_case(ref, name, case)
        np.savez_compressed(os.path.join(a.out, f"case_{name}.npz"), **res)
        stages = [str(res[f"c{c}_stage"]) for c in range(int(res["n_calls"]))]
        lens = [int(res[f"c{c}_seq_len"]) for c in range(int(res["n_calls"]))]
        print(f"{name}: stages={stages} seq_len={lens} sparsity={res[f'c{int(res['n_calls'])-1}_sparsity_list']}")


if __name__ == "__main__":
    main()
