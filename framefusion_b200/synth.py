"""Seeded synthetic inputs for the FrameFusion token-reduction path.

A sequence is ``[n_pre text rows | F frames x P patches of vision rows | n_post text rows]``.
Vision rows follow an AR(1) process along the frame axis of every patch column,
``x_f = r * x_{f-1} + sqrt(1 - r^2) * eps``, so that ``E[cos(x_{f-1}, x_f)] ~= r``.
The correlation ``r`` is drawn per (frame, patch) from ``U(r_lo, r_hi)`` unless
``per_patch_r`` is set (one value per patch column: runs that span all frames).

Everything is drawn on the CPU from ``torch.Generator().manual_seed(seed)`` and cast
to the requested dtype afterwards, so the oracle, the reference and the CUDA path see
identical bits when the same tensors are copied to a device.

The layout mirrors what the reference adapters hand to ``FrameFusion.prepare``
(/root/reference/framefusion/models/qwenvl/modeling_qwen2_vl.py:117-138):
``patch_type = [-1]*n_pre + list(range(P))*F + [-1]*n_post``.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch

TEXT_TOKEN = -1


@dataclasses.dataclass
class Workload:
    """One prefill-sized input of the token-reduction path."""

    hidden: torch.Tensor          # [1, S, H]
    cos: torch.Tensor             # [1, S, D]
    sin: torch.Tensor             # [1, S, D]
    patch_type: torch.Tensor      # [1, S] int64
    patch_num: int
    frames: int
    n_pre: int
    n_post: int

    @property
    def seq_len(self) -> int:
        return self.hidden.shape[1]

    @property
    def n_vision(self) -> int:
        return self.frames * self.patch_num

    def prepare_args(self):
        """Positional arguments of ``FrameFusion.prepare`` (main.py:15-38)."""
        n = self.n_vision
        return (self.patch_type, self.patch_num, self.n_pre, self.n_pre + n - 1, n, self.seq_len)


# name -> (frames, patches/frame, hidden, dtype, cost, similarity_lower_bound, ratio_lower_bound)
CONFIGS = {
    "C1": dict(frames=8, patch_num=196, hidden=1024, dtype=torch.float32, cost=0.3, slb=0.6, rlb=0.1),
    "C2": dict(frames=64, patch_num=576, hidden=3584, dtype=torch.bfloat16, cost=0.3, slb=0.6, rlb=0.1),
    "C3": dict(frames=128, patch_num=576, hidden=3584, dtype=torch.bfloat16, cost=0.5, slb=0.6, rlb=0.1),
    "C4": dict(frames=64, patch_num=729, hidden=4096, dtype=torch.bfloat16, cost=0.3, slb=0.6, rlb=0.1),
}


def make_workload(
    frames: int,
    patch_num: int,
    hidden: int,
    dtype: torch.dtype = torch.bfloat16,
    seed: int = 0,
    r_lo: float = 0.0,
    r_hi: float = 1.0,
    per_patch_r: bool = False,
    n_pre: int = 14,
    n_post: int = 20,
    rot_dim: int = 128,
    device: Optional[torch.device] = None,
    frozen_patches: int = 0,
    zero_rows=(),
) -> Workload:
    """``frozen_patches``: the first k patch columns repeat their frame-0 row in every frame (one run over all frames);
    ``zero_rows``: ``(frame, patch)`` pairs whose row is all zeros (zero norm -> NaN similarity).  Both are
    applied after the random draws, so the other rows are the same as without them."""
    g = torch.Generator().manual_seed(seed)
    F, P, H = frames, patch_num, hidden
    S = n_pre + F * P + n_post
    out = torch.empty(S, H, dtype=dtype)
    out[:n_pre] = torch.randn(n_pre, H, generator=g).to(dtype)
    if per_patch_r:
        r = (torch.rand(1, P, generator=g) * (r_hi - r_lo) + r_lo).expand(F, P)
    else:
        r = torch.rand(F, P, generator=g) * (r_hi - r_lo) + r_lo
    if frozen_patches:
        r = r.clone()
        r[:, :frozen_patches] = 1.0
    x = torch.randn(P, H, generator=g)
    for f in range(F):
        if f > 0:
            rf = r[f].unsqueeze(1)
            eps = torch.randn(P, H, generator=g)
            x = rf * x + torch.sqrt(1.0 - rf * rf) * eps
        out[n_pre + f * P: n_pre + (f + 1) * P] = x.to(dtype)
    for (f, p) in zero_rows:
        out[n_pre + f * P + p] = 0
    out[n_pre + F * P:] = torch.randn(n_post, H, generator=g).to(dtype)
    cos = torch.randn(1, S, rot_dim, generator=g).to(dtype)
    sin = torch.randn(1, S, rot_dim, generator=g).to(dtype)
    pt = torch.tensor([[TEXT_TOKEN] * n_pre + list(range(P)) * F + [TEXT_TOKEN] * n_post], dtype=torch.int64)
    wl = Workload(out.unsqueeze(0), cos, sin, pt, P, F, n_pre, n_post)
    if device is not None:
        wl = to_device(wl, device)
    return wl


def apply_drift(hidden: torch.Tensor, amount: float, seed: int, call: int) -> torch.Tensor:
    """Stand-in for "a decoder layer ran" between two FrameFusion calls: adds ``amount`` times one fixed
    random direction to every row (raises all adjacent-token similarities, so later calls merge on the
    ragged chains the earlier calls left behind).  Deterministic in (seed, call)."""
    g = torch.Generator().manual_seed(1_000_003 * (seed + 1) + call)
    v = torch.randn(hidden.shape[-1], generator=g)
    return (hidden.float().cpu() + amount * v).to(hidden.dtype).to(hidden.device)


def to_device(wl: Workload, device) -> Workload:
    return dataclasses.replace(
        wl,
        hidden=wl.hidden.to(device),
        cos=wl.cos.to(device),
        sin=wl.sin.to(device),
        patch_type=wl.patch_type.to(device),
    )


def make_attention_inputs(seq_len: int, n_heads: int = 28, n_kv_heads: int = 4, head_dim: int = 128,
                          dtype: torch.dtype = torch.bfloat16, seed: int = 0):
    """Random q / k for the importance kernel: q ``[1, n_heads, S, D]``, k ``[1, n_kv_heads, S, D]``."""
    g = torch.Generator().manual_seed(seed + 7919)
    q = torch.randn(1, n_heads, seq_len, head_dim, generator=g).to(dtype)
    k = torch.randn(1, n_kv_heads, seq_len, head_dim, generator=g).to(dtype)
    return q, k


def make_attention_row(seq_len: int, n_heads: int = 28, num: int = 1, dtype: torch.dtype = torch.bfloat16,
                       seed: int = 0) -> torch.Tensor:
    """A synthetic last-``num``-query attention-probability tensor ``[1, n_heads, num, S]``."""
    g = torch.Generator().manual_seed(seed + 104729)
    logits = torch.randn(1, n_heads, num, seq_len, generator=g) * 2.0
    return torch.softmax(logits, dim=-1).to(dtype)


def algorithmic_bytes(seq_len: int, kept: int, hidden: int, elem_bytes: int, rot_dim: int = 128) -> int:
    """Algorithmic HBM bytes of one merge-stage call."""
    return (seq_len * hidden * elem_bytes + kept * hidden * elem_bytes
            + 2 * (seq_len + kept) * rot_dim * elem_bytes + 8 * (seq_len + kept))
