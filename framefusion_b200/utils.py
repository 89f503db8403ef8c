"""Helpers with the reference's names (``/root/reference/framefusion/utils.py:10-57``).

``scaled_dot_product_attention`` returns the attention *probabilities* of the last ``num`` queries against
all keys — the importance signal the prune stage consumes — computed by ``ff_importance`` (one pass over K,
GQA aware) instead of ``repeat_kv`` + matmul + softmax.  The debug image dumps of the reference
(utils.py:59-101) are out of scope.
"""
from __future__ import annotations

import math
from typing import Any

import torch

from . import _lib
from .main import _dtype_code, _stream, FrameFusion

# meta
TEXT_TOKEN = -1
IGNORE_TOKEN = -2


def get_attr_by_name(obj: Any, name: str) -> Any:
    """
    Get an attribute from an object using a dot notation string.
    e.g., get_attr_by_name(model, "layers.0.self_attn.q_proj") will return model.layers[0].self_attn.q_proj
    """
    current = obj
    for level in name.split('.'):
        current = current[int(level)] if level.isdigit() else getattr(current, level)
    return current


def scaled_dot_product_attention(query, key, value, num=1, attn_mask=None, dropout_p=0.0,
                                 is_causal=False, scale=None, enable_gqa=False) -> torch.Tensor:
    """query ``[1, Hq, S, D]``, key ``[1, Hk, S, D]`` -> probabilities ``[1, Hq, num, S]`` in the query dtype.

    Same signature as the reference.  ``value`` is unused there too (only the weights are returned).  A key
    with fewer heads than the query is treated as grouped (what ``enable_gqa`` + ``repeat_interleave`` do in
    the reference) — callers can pass K before ``repeat_kv`` and save 7x of its bytes."""
    if not query.is_cuda:
        raise RuntimeError("framefusion_b200 runs on CUDA tensors only (there is no CPU fallback)")
    if attn_mask is not None:
        raise NotImplementedError("attn_mask is not supported (the reference hooks always pass None)")
    if dropout_p != 0.0:
        raise NotImplementedError("dropout_p must be 0 (inference)")
    bsz, n_q, q_len, d = query.shape
    assert bsz == 1, "Only support batch size 1"
    n_kv, s_len = key.shape[1], key.shape[2]
    if n_q % n_kv != 0:
        raise ValueError(f"{n_q} query heads are not a multiple of {n_kv} key heads")
    if n_kv != n_q and not enable_gqa:
        raise RuntimeError("key has fewer heads than query: pass enable_gqa=True")
    num = min(num, q_len)
    if query.stride(-1) != 1:
        query = query.contiguous()
    if key.stride(-1) != 1 or key.dtype != query.dtype:
        key = key.to(query.dtype).contiguous()
    device = query.device
    st = FrameFusion._static(device)
    scale_factor = 1 / math.sqrt(d) if scale is None else scale
    out = torch.empty((1, n_q, num, s_len), dtype=query.dtype, device=device)
    scratch = torch.empty(n_q * num * s_len, dtype=torch.float32, device=device)
    # the last `num` queries: pass the full tensor with its strides, the kernel offsets to S - num
    _lib.check(st.lib.ff_importance(
        st.ctx, query.data_ptr() + (q_len - s_len) * query.stride(2) * query.element_size() if q_len != s_len else query.data_ptr(),
        key.data_ptr(), _dtype_code(query), n_q, n_kv, s_len, d, num,
        query.stride(1), query.stride(2), key.stride(1), key.stride(2), 1 if is_causal else 0, float(scale_factor),
        out.data_ptr(), scratch.data_ptr(), scratch.numel() * 4, _stream(device)))
    return out
