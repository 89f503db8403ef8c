"""CPU oracle for the fixed-amount comparison methods (FastV, fixed-sparsity merging) — TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this module (same rule as ``oracle/ff_oracle.py``); nothing under ``framefusion_b200/`` does.

What is restated (``/root/reference/framefusion/models/qwen2/modeling_qwen2_baseline.py``):

* FastV selection, :318-342 — ``mean`` of the last query's attention over the heads, top ``round(L * (1 - r))`` of the vision
  span, text tokens kept, ascending order.  That op sequence is the prune stage of ``main.py:61-92`` with the ratio given,
  so it is built from ``ff_oracle.mean_heads`` / ``prune_keep_indices`` — both PINNED by the golden fixtures of the
  unmodified reference (``tests/test_oracle_golden.py``).
* fixed-sparsity selection, :916-920 and :1003 — ``prune_num = floor(sparsity * n_vis)``, ``torch.topk`` of the by-patch
  similarities — on top of ``ff_oracle.similarity_by_patch`` / ``topk_lowest_index`` / ``merge_tokens_and_get_mask``
  (pinned).  The merge ARITHMETIC is FrameFusion's (main.py:285-317), not the baseline's ``mean`` of normalised states:
  ``framefusion_b200/baselines.py`` says why, and that difference is deliberate.

Parity status of the baseline DRIVERS (which layer calls what): UNPINNED — ``modeling_qwen2_baseline.py`` does not import
under transformers 5.5 (``Qwen2SdpaAttention``, ``SinkCache``, ``QWEN2_INPUTS_DOCSTRING`` are gone) and nothing in the
reference calls it, so no fixture of the unmodified file could be generated; the schedule is restated from the lines cited.
"""
from __future__ import annotations

import math

import numpy as np

from . import ff_oracle as orc


class OracleBaseline:
    """numpy mirror of ``framefusion_b200.baselines.TokenReductionBaseline`` (B = 1): hidden ``[S, H]``, position
    embeddings a list of two ``[S, D]`` arrays, mask ``[S, S]`` or None."""

    def __init__(self, sparsity=None, fastv_k=None, fastv_r=0.5, dtype="bf16"):
        self.sparsity = None if sparsity is None else list(sparsity)
        self.fastv_k = fastv_k
        self.fastv_r = fastv_r
        self.dtype = dtype

    def prepare(self, patch_type, patch_num, image_token_start_index, image_token_end_index, image_token_length,
                original_length):
        self.patch_type = None if patch_type is None else np.asarray(patch_type).reshape(-1).astype(np.int64)
        self.patch_num = patch_num
        self.start = int(image_token_start_index)
        self.image_token_length = int(image_token_length)
        self.original_length = int(original_length)
        self.last = None

    def wants_attention(self, layer_idx):
        return self.fastv_k is not None and layer_idx == self.fastv_k - 1

    @staticmethod
    def _take(pos, sel):
        return [p[sel] for p in pos]

    def merge_at(self, layer_idx, hidden, pos, mask):
        self.last = None
        if self.sparsity is None or hidden.shape[0] <= 1 or layer_idx >= len(self.sparsity):
            return hidden, pos, mask
        n_vis = int((self.patch_type != orc.TEXT_TOKEN).sum())
        k = math.floor(self.sparsity[layer_idx] * n_vis)                      # reference :918
        if k <= 0:
            return hidden, pos, mask
        sr = orc.similarity_by_patch(hidden, self.patch_type, self.patch_num, self.dtype)
        mi = orc.topk_lowest_index(sr.sim, k)                                  # reference :1003
        merged, keep = orc.merge_tokens_and_get_mask(hidden, sr.order, mi, self.dtype)
        self.last = dict(stage="merge", sim=sr, merge_index=mi, keep_mask=keep, k=k)
        self.patch_type = self.patch_type[keep]
        return merged[keep], self._take(pos, keep), None if mask is None else mask[keep][:, keep]

    def fastv_at(self, layer_idx, hidden, pos, mask, attn):
        """attn ``[heads, 1, S]`` in T."""
        self.last = None
        if self.fastv_k is None or layer_idx != self.fastv_k or hidden.shape[0] <= 1:
            return hidden, pos, mask
        q_len = hidden.shape[0]
        length = self.image_token_length - (self.original_length - q_len)     # reference :1451 (== L without merges, :304)
        imp = orc.mean_heads(attn, self.dtype)                                 # reference :321-323
        keep = orc.prune_keep_indices(imp, self.start, length, q_len, self.fastv_r)    # reference :325-331
        self.last = dict(stage="prune", keep=keep, importance=imp, start=self.start, length=length)
        if self.patch_type is not None:
            self.patch_type = self.patch_type[keep]
        return hidden[keep], self._take(pos, keep), None if mask is None else mask[keep][:, keep]
