"""``patch_type`` builders: the token layout the embed-stage patches of the reference hand to ``FrameFusion.prepare``.

The reference builds it inside patched third-party functions (``models/qwenvl/modeling_qwen2_vl.py:117-138`` for
Qwen2-VL, ``models/llava_video/modeling_llava_video.py:321-339`` for LLaVA-Video); those functions belong to code
that is not vendored here, so the arithmetic is offered on its own.  Each returns the positional arguments of
``prepare`` — ``(patch_type [1,S] int64, patch_num, image_token_start_index, image_token_end_index,
image_token_length, original_length)`` — with ``patch_type = [-1]*start + list(range(patch_num))*n_frames +
[-1]*tail``.
"""
from __future__ import annotations

import math

import torch

TEXT_TOKEN = -1


def _layout(start: int, patch_num: int, image_token_length: int, original_length: int, device):
    """``[-1]*start + range(patch_num)*n_frames + [-1]*(original_length - end - 1)``, built the way the reference builds it
    (qwenvl/modeling_qwen2_vl.py:125-127).  If ``image_token_length % patch_num != 0`` that list is SHORTER than the
    sequence (the tail shifts left) and the reference fails later at ``patch_type[keep_mask]`` (main.py:132); here
    ``FrameFusion`` raises as soon as it sees the length mismatch, so the same construction is kept."""
    end = start + image_token_length - 1
    n_frames = image_token_length // patch_num
    body = torch.arange(patch_num, dtype=torch.int64).repeat(n_frames)
    tail = original_length - end - 1
    pt = torch.cat([torch.full((start,), TEXT_TOKEN, dtype=torch.int64), body,
                    torch.full((max(tail, 0),), TEXT_TOKEN, dtype=torch.int64)])[None]
    return pt.to(device), patch_num, start, end, image_token_length, original_length


def qwen2vl_prepare_args(input_ids: torch.Tensor, video_token_id: int, video_grid_thw: torch.Tensor,
                         spatial_merge_size: int, device=None):
    """Qwen2-VL: one video whose placeholder tokens sit in ``input_ids`` (reference :118-127).
    ``patch_num = H*W / merge^2`` tokens per frame, the span is first..last video token."""
    ids = input_ids[0]
    where = torch.where(ids == video_token_id)[0]
    if where.numel() == 0:
        raise ValueError("no video tokens in input_ids")
    patch_num = int((int(video_grid_thw[0, 1]) * int(video_grid_thw[0, 2])) / (spatial_merge_size * spatial_merge_size))
    start, end = int(where[0]), int(where[-1])
    return _layout(start, patch_num, end - start + 1, ids.numel(), device if device is not None else input_ids.device)


def llava_video_prepare_args(input_ids: torch.Tensor, image_token_index: int, image_token_length: int,
                             num_patches_per_side: int, spatial_pool_mode: str = "bilinear", device=None):
    """LLaVA-Video: the single image placeholder of ``input_ids`` expands to ``image_token_length`` features; a frame
    is ``ps * (ps + 1)`` tokens — ``ps`` pooled patches per side plus one newline token per row (reference :322-336)."""
    ps = math.ceil(num_patches_per_side / 2) if spatial_pool_mode == "bilinear" else num_patches_per_side // 2
    patch_num = ps * (ps + 1)
    where = torch.where(input_ids[0] == image_token_index)[0]
    if where.numel() != 1:
        raise ValueError("expected exactly one image placeholder (the reference asserts num_images == 1)")
    start = int(where[0])
    original_length = input_ids[0].numel() + image_token_length - 1
    return _layout(start, patch_num, image_token_length, original_length, device if device is not None else input_ids.device)
