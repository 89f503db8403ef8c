"""Layer placement for a decoder split over the GPUs of one box — what ``device_map="auto"`` + accelerate hooks do
for the reference (README.md:120, interface.py:204-207), without accelerate.

``split_layers(llm, devices)`` puts contiguous blocks of decoder layers on the given devices (embedding and rotary
module with the first block, final norm with the last) and registers a forward pre-hook on every layer that moves
its inputs — ``hidden_states``, the attention mask, and the two entries of the ``position_embeddings`` LIST that
FrameFusion keeps swapping — to the layer's device.  At a block boundary that is one peer copy of the CURRENT,
already reduced, sequence over NVLink; there is no collective.  The FrameFusion operator follows by itself: it
keeps one context / workspace per device and rebuilds its chain links from ``patch_type`` when the device changes
(``main.py:106`` of the reference does the same ``.to(device)``).
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn


def _to(x, device):
    if isinstance(x, torch.Tensor):
        return x if x.device == device else x.to(device, non_blocking=True)
    if isinstance(x, list):
        for k, v in enumerate(x):
            x[k] = _to(v, device)          # in place: the hooks hand the same list from layer to layer
        return x
    if isinstance(x, tuple):
        return tuple(_to(v, device) for v in x)
    if isinstance(x, dict):
        return {k: _to(v, device) for k, v in x.items()}
    return x


def layer_devices(n_layers: int, devices: Sequence[torch.device]):
    """Contiguous blocks, sizes differing by at most one."""
    per, extra = divmod(n_layers, len(devices))
    out = []
    for k, d in enumerate(devices):
        out += [torch.device(d)] * (per + (1 if k < extra else 0))
    return out


def split_layers(llm: nn.Module, devices: Sequence, decoder_key: str = "layers"):
    """Places ``llm`` (a decoder stack with ``embed_tokens``, ``layers``, ``norm``, ``rotary_emb``) on ``devices``."""
    layers = getattr(llm, decoder_key)
    placement = layer_devices(len(layers), [torch.device(d) for d in devices])
    first, last = placement[0], placement[-1]
    for name in ("embed_tokens", "rotary_emb"):
        if hasattr(llm, name):
            getattr(llm, name).to(first)
    if hasattr(llm, "norm"):
        llm.norm.to(last)
        llm.norm.register_forward_pre_hook(lambda m, args: _to(args, last))
    for layer, dev in zip(layers, placement):
        layer.to(dev)

        def pre(module, args, kwargs, dev=dev):
            return _to(args, dev), _to(kwargs, dev)
        layer.register_forward_pre_hook(pre, with_kwargs=True)
    return placement
