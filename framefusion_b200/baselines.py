"""The comparison methods of the reference — FastV, fixed-sparsity merging and their combination — on the SAME kernels
as FrameFusion, with the amounts GIVEN instead of derived from a budget.

Reference: ``/root/reference/framefusion/models/qwen2/modeling_qwen2_baseline.py``

* FastV (``replace_Qwen2_fastv`` :175-188; selection :300-342): before decoder layer ``fastv_k`` the vision span keeps the
  ``round(L * (1 - fastv_r))`` tokens the LAST query of layer ``fastv_k - 1`` attends to most (mean over heads), text
  tokens stay, order is preserved; hidden states and position embeddings are compacted, ``position_ids`` become the
  kept indices.  Here: ``ff_importance`` (num = 1) in the attention of layer ``fastv_k - 1`` and ONE ``ff_prune_layer``
  call with that ``k`` — the prune stage of FrameFusion (main.py:61-101) with a given ratio.
* fixed-sparsity merging (``replace_Qwen2_merging`` :860-874; :916-1070): at every prefill layer ``i`` the
  ``floor(sparsity[i] * n_vis)`` adjacent-frame tokens of highest cosine similarity merge into their chain predecessor.
  Here: ONE ``ff_merge_layer`` call per layer on its top-k branch with ``k = int(sparsity[i] * n_vis)``.
* merge-then-FastV (``replace_Qwen2_merge_then_fastv`` :1339-1355): both, the vision length at layer ``fastv_k`` being what
  the merges left (:1451).
* StreamingLLM (:579-617) is an attention-sink + sliding-window ATTENTION kernel from the un-vendored ``minference``
  package (the reference imports it optionally, :13-17); it moves no rows of ``hidden_states`` and is not part of this
  path: ``replace_Qwen2_forward(mode="streamingllm")`` raises ``NotImplementedError`` saying so.

Deliberate differences from the reference baseline, documented in DESIGN.md:

* fixed-sparsity merging uses FrameFusion's arithmetic (main.py:216-238, 285-317: similarity on the residual stream, runs
  summed in chain order with one rounding per add, one division) on the residual stream at the layer's entry.  The
  reference baseline averages the *normalised* hidden states with ``mean`` for q/k/v and merely DROPS the merged rows of
  the residual (:1043-1066, :1176-1180).  The selection rule (top ``floor(s * n_vis)`` by-patch similarities) is the same.
* a 4-D ``attention_mask`` is compacted with the tokens (the reference's FastV passes the stale full-length mask on, :311).
* top-k ties go to the lowest index (``torch.topk`` leaves them unspecified).

Batch size 1, CUDA tensors only, no CPU fallback — like ``FrameFusion``.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import _lib
from .main import FrameFusion, TEXT_TOKEN, _stream


def compute_density_overhead(sparsity_list) -> tuple:
    """(average token density over the layers, density after the last layer) of a per-layer sparsity list
    (reference modeling_qwen2_baseline.py:26-39)."""
    alive, total = 1.0, 0.0
    for s in sparsity_list:
        alive *= 1 - s
        total += alive
    return total / len(sparsity_list), alive


class TokenReductionBaseline(FrameFusion):
    """One operator for the fixed-amount methods: ``sparsity`` (per-layer merged fraction of the vision tokens, or None)
    and ``fastv_k`` / ``fastv_r`` (layer index in front of which FastV prunes, fraction pruned; ``fastv_k`` None: off).

    The hooks (``hooks/qwen2_baselines.py``) call ``merge_at(layer_idx, ...)`` on the input of every decoder layer and
    ``fastv_at(layer_idx, ...)`` on the input of layer ``fastv_k``; ``wants_attention(layer_idx)`` tells the attention
    of layer ``fastv_k - 1`` to produce the last query's probabilities."""

    def __init__(self, sparsity: Optional[Sequence[float]] = None, fastv_k: Optional[int] = None, fastv_r: float = 0.5):
        super().__init__(cost=1.0, similarity_lower_bound=-3.0, ratio_lower_bound=0.0)
        self.sparsity = None if sparsity is None else list(sparsity)
        self.fastv_k = fastv_k
        self.fastv_r = fastv_r
        self.frame_token_num = None
        self.keep_indexs = None
        self._keep_src = None

    def prepare(self, patch_type, patch_num, image_token_start_index, image_token_end_index, image_token_length,
                original_length, finish_merging=False, finish_pruning=False, sparsity_list: List[float] = None):
        """Same arguments as ``FrameFusion.prepare`` (the embed-stage patches fill them, interface.py:140-166); FastV alone
        needs ``image_token_start_index`` / ``image_token_length`` / ``original_length`` only (reference :303-304)."""
        super().prepare(patch_type, patch_num, image_token_start_index, image_token_end_index, image_token_length,
                        original_length, finish_merging, finish_pruning, sparsity_list)
        self.frame_token_num = None
        self.keep_indexs = None
        self._keep_src = None

    # ---- what the hooks ask ------------------------------------------------------------------------------------
    def wants_attention(self, layer_idx: int) -> bool:
        return self.fastv_k is not None and layer_idx == self.fastv_k - 1

    def merge_at(self, layer_idx: int, hidden_states, position_embeddings, attention_mask):
        """Fixed-sparsity merge in front of decoder layer ``layer_idx`` (reference :916-920: ``prune_num = floor(sparsity *
        frame_token_num)``, nothing happens for 0)."""
        if self.sparsity is None or hidden_states.shape[1] <= 1 or layer_idx >= len(self.sparsity):
            return hidden_states, position_embeddings, attention_mask
        s = self.sparsity[layer_idx]
        if self.frame_token_num is None:
            # (one host read per prefill; afterwards the count follows from what each call reports)
            self.frame_token_num = int((self.patch_type != TEXT_TOKEN).sum().item())
        if math.floor(s * self.frame_token_num) <= 0:
            return hidden_states, position_embeddings, attention_mask
        out = self._merge(hidden_states, position_embeddings, attention_mask, fixed_sparsity=s)
        st = self._state(hidden_states.device)
        self.frame_token_num -= int(st.status[_lib.ST_NMERGED])
        return out

    def fastv_at(self, layer_idx: int, hidden_states, position_embeddings, attention_mask, last_layer_attention):
        """FastV in front of decoder layer ``layer_idx == fastv_k`` (reference :318-342); ``last_layer_attention`` is the
        ``[1, heads, 1, S]`` output of layer ``fastv_k - 1``'s attention.  Leaves the kept indices in ``keep_indexs``
        (a callable: the index tensor is only built when somebody asks)."""
        if self.fastv_k is None or layer_idx != self.fastv_k or hidden_states.shape[1] <= 1:
            return hidden_states, position_embeddings, attention_mask
        q_len = hidden_states.shape[1]
        out = self._prune(hidden_states, position_embeddings, attention_mask, last_layer_attention, pruning_ratio=self.fastv_r)
        self._keep_src = (hidden_states.device, q_len)
        self.keep_indexs = self._read_keep_indexs
        if self.sparsity is not None and any(s > 0 for s in self.sparsity[layer_idx:]):
            # merges follow: the token layout loses the pruned rows as well
            keep = self._read_keep_indexs()
            self.patch_type = self.patch_type.to(keep.device).reshape(-1)[keep].reshape(1, -1)
            self._links_for = None
            self.frame_token_num = int((self.patch_type != TEXT_TOKEN).sum().item())
        return out

    def _read_keep_indexs(self) -> torch.Tensor:
        """Kept sequence positions of the last FastV call, ascending (reference :329-331)."""
        device, q_len = self._keep_src
        st = self._state(device)
        wp, wb = st.ws_ptr()
        keep = torch.empty(q_len, dtype=torch.uint8, device=device)
        _lib.check(st.lib.ff_debug_read(st.ctx, wp, wb, 0, keep.data_ptr(), q_len, 0, _stream(device)))
        return torch.nonzero(keep, as_tuple=False).reshape(-1)

    def forward(self, *args, **kwargs):
        raise RuntimeError("TokenReductionBaseline is driven through merge_at / fastv_at (hooks/qwen2_baselines.py)")
