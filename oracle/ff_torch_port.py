"""torch-CPU port of the reference token-reduction operator — TEST / BASELINE INFRASTRUCTURE ONLY.

What this is for.  ``bench.py`` has to time "the reference's own CPU implementation of the path" on the GPU
box's host cores, but ``/root/reference`` does not travel to that box and its sources may not be copied into
this repository.  The reference operator is ~380 lines of ATen calls (``/root/reference/framefusion/main.py``);
this file restates the same algorithm with the same ATen arithmetic (so the numbers it produces on a CPU are the
reference's numbers — pinned by ``tests/test_torch_port.py`` against the golden fixtures generated from the
unmodified reference) and with all host threads, so its wall time is a fair stand-in for the reference's
torch-CPU path (``cpu_baseline.kind == "port"``).  It is written independently of the reference's code
structure: by-patch order through one stable ``argsort`` instead of a ``[P, S]`` comparison matrix, runs through
a cumulative maximum instead of per-run-length loops, one norm per row instead of one per gathered copy.  Where
that makes it *faster* than the reference, the baseline it sets is harder to beat, not easier.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs import this module; nothing under
``framefusion_b200/`` does.

Line references are to ``/root/reference/framefusion/main.py``.
"""
from __future__ import annotations

from typing import List, Optional

import torch

TEXT_TOKEN = -1
IGNORE_TOKEN = -2


def pruning_ratio(sparsity_list, cost, num_layers=28):
    """Budget formula (:321-343), Python doubles."""
    s, spent = 1, 0
    for x in sparsity_list:
        s *= (1 - x)
        spent += s
    remain = num_layers * cost - spent
    if remain < 0:
        raise ValueError("The cost is too small")
    share = remain / ((num_layers - len(sparsity_list)) * s)
    return 0 if share > 1 else 1 - share


def by_patch_order(patch_type: torch.Tensor, patch_num) -> torch.Tensor:
    """Sequence indices of tokens with patch id in ``[0, patch_num)``, stably sorted by id (:208-214)."""
    pt = patch_type.reshape(-1)
    n_ids = int(-(-float(patch_num) // 1))
    vis = torch.nonzero((pt >= 0) & (pt < n_ids)).reshape(-1)
    return vis[torch.argsort(pt[vis], stable=True)]


def similarity_by_patch(hidden: torch.Tensor, patch_type: torch.Tensor, patch_num):
    """``(sim [N] in hidden dtype with -2 at chain heads, order [N])`` (:180-241, :345-349).

    Every intermediate stays in the hidden dtype, as in the reference: product tensor, its row sum, the row
    norms, their product, the quotient."""
    h = hidden[0]
    order = by_patch_order(patch_type, patch_num)
    n = order.numel()
    sim = torch.full((n,), IGNORE_TOKEN, dtype=h.dtype, device=h.device)
    if n < 2:
        return sim, order
    rows = h[order]                                        # by-patch rows, gathered once
    norms = torch.norm(rows, dim=-1)                       # one norm per row; both neighbours reuse it
    dots = torch.sum(rows[:-1] * rows[1:], dim=-1)
    body = dots / (norms[:-1] * norms[1:])
    ids = patch_type.reshape(-1)[order]
    same = ids[1:] == ids[:-1]
    sim[1:] = torch.where(same, body, torch.full_like(body, IGNORE_TOKEN))
    return sim, order


def merge_runs_(hidden: torch.Tensor, order: torch.Tensor, merge_index: torch.Tensor) -> torch.Tensor:
    """In-place run merge + keep mask ``[S]`` (:243-319).  A run of flagged by-patch positions folds into the
    unflagged position in front of it: members are added one at a time in ascending order in the hidden dtype
    (``index_add_`` on CPU is sequential), then the sum is divided by the member count + 1."""
    h = hidden[0]
    dev = h.device
    keep = torch.ones(h.shape[0], dtype=torch.bool, device=dev)
    if merge_index.numel() == 0:
        return keep
    n = order.numel()
    flagged = torch.zeros(n, dtype=torch.bool, device=dev)
    flagged[merge_index] = True
    keep[order[merge_index]] = False
    pos = torch.arange(n, device=dev)
    anchor = torch.where(flagged, torch.full_like(pos, -1), pos).cummax(0).values    # last unflagged position <= j
    members = merge_index
    a_of_m = anchor[members]                               # -1 (run at position 0) wraps like python indexing
    a_rows = order[a_of_m]
    # along dim 1 of the [1, S, H] tensor, as the reference calls it: that ATen path adds slice by slice in the
    # hidden dtype (the 2-D dim-0 path of torch 2.11 accumulates differently for runs of 2+)
    hidden.index_add_(1, a_rows, hidden[:, order[members], :])
    anchors, counts = torch.unique_consecutive(a_of_m, return_counts=True)
    rows = order[anchors]
    h[rows] = h[rows] / (counts + 1).to(h.dtype).unsqueeze(-1)
    return keep


class TorchPortFrameFusion:
    """Same constructor / ``prepare`` / call contract as the reference module (:8-140)."""

    def __init__(self, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1):
        self.cost = cost
        self.similarity_lower_bound = similarity_lower_bound
        self.ratio_lower_bound = ratio_lower_bound
        self.trace = True                 # keep what flowed between the stages (the parity harness reads it)

    def prepare(self, patch_type, patch_num, image_token_start_index, image_token_end_index, image_token_length,
                original_length, finish_merging=False, finish_pruning=False, sparsity_list: Optional[List[float]] = None):
        self.patch_type = patch_type
        self.patch_num = patch_num
        self.image_token_start_index = image_token_start_index
        self.image_token_end_index = image_token_end_index
        self.image_token_length = image_token_length
        self.original_length = original_length
        self.finish_merging = finish_merging
        self.finish_pruning = finish_pruning
        self.sparsity_list = [] if sparsity_list is None else sparsity_list
        self.last = None

    @staticmethod
    def _select_pos(pos, sel):
        """``sel`` is a bool mask or an index tensor over the sequence axis (:142-178)."""
        if type(pos) == list:
            assert len(pos) == 2
            for i in range(2):
                if pos[i].ndim == 3:
                    pos[i] = pos[i][:, sel, :]
                elif pos[i].ndim == 4:
                    pos[i] = pos[i][:, :, sel, :]
                else:
                    raise NotImplementedError("Only support 3D or 4D position embeddings")
            return pos
        if type(pos) == torch.Tensor:
            if pos.ndim != 2:
                raise NotImplementedError("Only support 2D position embeddings")
            return pos[:, sel]
        raise NotImplementedError("Only support list or tensor for position embeddings")

    def __call__(self, hidden_states, position_embeddings, attention_mask, self_attn_weights=None):
        bsz, q_len, _ = hidden_states.shape
        self.last = None

        if q_len > 1 and self.finish_merging and not self.finish_pruning:            # prune stage (:61-101)
            start = int(self.image_token_start_index)
            length = int(self.image_token_length - (self.original_length - q_len))
            imp = self_attn_weights.mean(dim=(1, 2))[0]
            ratio = pruning_ratio(self.sparsity_list, self.cost)
            k = round(length * (1 - ratio))
            top = torch.topk(imp[start:start + length], k).indices + start
            dev = hidden_states.device
            keep = torch.cat((torch.arange(start, device=dev), top, torch.arange(start + length, q_len, device=dev))).sort().values
            hidden_states = hidden_states[:, keep, :]
            position_embeddings = self._select_pos(position_embeddings, keep)
            if attention_mask is not None:
                attention_mask = attention_mask[:, :, keep, :][:, :, :, keep]
            self.finish_pruning = True
            if self.trace:
                self.last = dict(stage="prune", keep=keep.cpu().numpy(), importance=imp.float().cpu().numpy(), start=start, length=length)

        if q_len > 1 and not self.finish_merging:                                    # merge stage (:104-138)
            assert bsz == 1, "Only support batch size 1"
            self.patch_type = self.patch_type.to(hidden_states.device)
            bound = pruning_ratio(self.sparsity_list, self.cost)
            sim, order = similarity_by_patch(hidden_states, self.patch_type, self.patch_num)
            n_vis = int((self.patch_type != TEXT_TOKEN).sum())
            merge_index = torch.nonzero(sim >= self.similarity_lower_bound).reshape(-1)
            ratio = merge_index.numel() / n_vis
            if ratio < bound:
                branch = "threshold"
                self.sparsity_list.append(ratio)
                if ratio < self.ratio_lower_bound:
                    self.finish_merging = True
            else:
                branch = "topk"
                merge_index = torch.topk(sim, int(bound * n_vis)).indices.sort().values
                self.finish_merging = True
                self.finish_pruning = True
            keep = merge_runs_(hidden_states, order, merge_index)
            self.patch_type = self.patch_type[:, keep]
            hidden_states = hidden_states[:, keep, :]
            position_embeddings = self._select_pos(position_embeddings, keep)
            if attention_mask is not None:
                attention_mask = attention_mask[:, :, keep, :][:, :, :, keep]
            if self.trace:
                self.last = dict(stage="merge", branch=branch, keep_mask=keep.cpu().numpy(), merge_index=merge_index.cpu().numpy(),
                                 sim_values=sim.float().cpu().numpy(), order=order.cpu().numpy())
        return hidden_states, position_embeddings, attention_mask


def last_query_attention(query, key, num=1, is_causal=False, scale=None):
    """Attention probabilities of the last ``num`` queries (``/root/reference/framefusion/utils.py:27-57``):
    query ``[1, Hq, S, D]``, key ``[1, Hk, S, D]`` -> ``[1, Hq, num, S]`` in the query dtype."""
    hq, hk = query.shape[1], key.shape[1]
    if hk != hq:
        key = key.repeat_interleave(hq // hk, dim=1)
    q = query[:, :, -num:, :]
    length, s_len = q.shape[-2], key.shape[-2]
    scale = q.shape[-1] ** -0.5 if scale is None else scale
    bias = torch.zeros(length, s_len, dtype=q.dtype, device=q.device)
    if is_causal:
        hide = torch.ones(length, s_len, dtype=torch.bool, device=q.device).triu(diagonal=s_len - length + 1)
        bias.masked_fill_(hide, float("-inf"))
    w = q @ key.transpose(-2, -1) * scale
    w += bias
    return torch.softmax(w, dim=-1)
