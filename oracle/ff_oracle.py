"""CPU oracle for the FrameFusion token-reduction path — TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm in
``/root/reference/framefusion/main.py`` (and ``utils.py:27-57``).  It exists to *check*
the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``framefusion_b200/`` imports it and the product path has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function here
against fixtures in ``tests/golden/`` that were produced by importing the unmodified
reference in the build container (``oracle/gen_golden.py``), plus the one known-answer
vector the reference itself carries (``main.py:361-363``).

Numerics.  A tensor of dtype T (bf16 / f16 / f32) is held as a float32 ndarray whose
values are exactly representable in T.  Every reference op that materialises a tensor
in T is restated as "compute in float32/float64, round to T":

* ``cosine_similarity`` (main.py:345-349) is the rounding chain
  ``T(T(sum_f32 T(a*b)) / T(T(|a|) * T(|b|)))`` with ``|a| = T(sqrt(sum_f32 a*a))``;
* the ``>=`` threshold is compared in T (the Python scalar is cast to the tensor dtype);
* ``index_add_`` (main.py:304-311) adds run members one by one, each add rounded to T,
  in ascending chain order; the average is one float32 division rounded to T (:314-317).

The only thing ATen leaves undefined is the *order* of a float32 row sum.  The oracle
sums in float64 and reports, per similarity, whether any float32 ordering could land
on the other side of a T rounding boundary (``SimResult.fragile`` with the bracketing
values ``lo``/``hi``).  A checker requires bit equality where ``fragile`` is False and
``lo <= x <= hi`` elsewhere.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

TEXT_TOKEN = -1
IGNORE_TOKEN = -2

_U32 = np.float32(2.0 ** -24)          # float32 unit round-off
_SUM_SLACK = 16.0                      # |fp32 row sum - exact| <= _SUM_SLACK * u * sum|x|  (any sane order)


# ----------------------------------------------------------------------------------------------
# dtype rounding
# ----------------------------------------------------------------------------------------------
def _round_bf16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    nan = np.isnan(x)
    bias = np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    r = ((u + bias) & np.uint32(0xFFFF0000)).view(np.float32)
    if nan.any():
        r = r.copy()
        r[nan] = np.nan
    return r


def _round_f16(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def _round_f32(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float32)


_ROUND = {"bf16": _round_bf16, "f16": _round_f16, "f32": _round_f32}
ELEM_BYTES = {"bf16": 2, "f16": 2, "f32": 4}


def round_to(x, dtype: str) -> np.ndarray:
    """Round float32/float64 values to dtype T (returned as float32)."""
    with np.errstate(over="ignore", invalid="ignore"):
        return _ROUND[dtype](np.asarray(x, dtype=np.float32))


def bits_to_f32(bits: np.ndarray, dtype: str) -> np.ndarray:
    """Raw storage (uint16 for bf16/f16, float32 for f32) -> float32 values."""
    if dtype == "bf16":
        return (bits.astype(np.uint32) << np.uint32(16)).view(np.float32)
    if dtype == "f16":
        return bits.view(np.float16).astype(np.float32)
    bits = np.asarray(bits)
    return bits.view(np.float32) if bits.dtype == np.uint32 else bits.astype(np.float32)


def f32_to_bits(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "bf16":
        return (np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) >> np.uint32(16)).astype(np.uint16)
    if dtype == "f16":
        return np.asarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)
    return np.asarray(x, dtype=np.float32)


# ----------------------------------------------------------------------------------------------
# budget formula  (main.py:321-343)
# ----------------------------------------------------------------------------------------------
def compute_pruning_ratio(sparsity_list: Sequence[float], cost: float, num_layers: int = 28) -> float:
    s = 1
    total = 0
    for x in sparsity_list:
        s *= (1 - x)
        total += s
    remain = num_layers * cost - total
    if remain < 0:
        raise ValueError("The cost is too small")
    q = remain / ((num_layers - len(sparsity_list)) * s)
    if q > 1:
        return 0
    return 1 - q


# ----------------------------------------------------------------------------------------------
# run-length helper  (main.py:351-380) — known-answer vector in its docstring
# ----------------------------------------------------------------------------------------------
def find_contiguous_latter_index(flags: np.ndarray) -> np.ndarray:
    """``[0,1,1,1,0,0,1,1] -> [0,0,0,3,0,0,0,2]`` for a 1-D 0/1 array."""
    f = np.asarray(flags).astype(np.int64)
    prev = np.concatenate(([0], f[:-1]))
    nxt = np.concatenate((f[1:], [0]))
    starts = np.nonzero((f == 1) & (prev == 0))[0]
    ends = np.nonzero((f == 1) & (nxt == 0))[0]
    out = np.zeros_like(f)
    out[ends] = ends - starts + 1
    return out


# ----------------------------------------------------------------------------------------------
# by-patch order and similarity  (main.py:180-241, 345-349)
# ----------------------------------------------------------------------------------------------
def order_by_patch(patch_type: np.ndarray, patch_num) -> np.ndarray:
    """Indices of tokens whose patch id is in ``[0, patch_num)``, stably sorted by patch id (:208-214)."""
    pt = np.asarray(patch_type).reshape(-1).astype(np.int64)
    n_ids = int(math.ceil(float(patch_num)))          # torch.arange(float) semantics (nvila passes a float)
    idx = np.nonzero((pt >= 0) & (pt < n_ids))[0]
    return idx[np.argsort(pt[idx], kind="stable")]


@dataclasses.dataclass
class SimResult:
    sim: np.ndarray        # [N] float32 holding T values; -2 at chain heads (by-patch order)
    order: np.ndarray      # [N] int64 sequence index of each by-patch position
    fragile: np.ndarray    # [N] bool: a float32 summation order could change sim[j]
    lo: np.ndarray         # [N] lower / upper bracket of the admissible value (== sim where not fragile)
    hi: np.ndarray


def _row_sums(hidden: np.ndarray, rows_a: np.ndarray, rows_b: np.ndarray, dtype: str, chunk: int = 4096):
    """float64 ``sum T(a*b)`` and ``sum |T(a*b)|`` for row pairs."""
    n = rows_a.shape[0]
    d = np.empty(n, np.float64)
    a_abs = np.empty(n, np.float64)
    for s in range(0, n, chunk):
        a = hidden[rows_a[s:s + chunk]]
        b = hidden[rows_b[s:s + chunk]]
        prod = round_to(a * b, dtype).astype(np.float64)
        d[s:s + chunk] = prod.sum(axis=1)
        a_abs[s:s + chunk] = np.abs(prod).sum(axis=1)
    return d, a_abs


def _sq_sums(hidden: np.ndarray, rows: np.ndarray, chunk: int = 4096) -> np.ndarray:
    q = np.empty(rows.shape[0], np.float64)
    for s in range(0, rows.shape[0], chunk):
        a = hidden[rows[s:s + chunk]].astype(np.float64)
        q[s:s + chunk] = (a * a).sum(axis=1)
    return q


def _chain(dot_t, n1_t, n2_t, dtype):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        den = round_to(n1_t.astype(np.float32) * n2_t.astype(np.float32), dtype)
        return round_to(dot_t.astype(np.float32) / den, dtype)


def similarity_by_patch(hidden: np.ndarray, patch_type: np.ndarray, patch_num, dtype: str) -> SimResult:
    """Restates ``compute_similarity_and_token_index_by_patch`` (main.py:180-241)."""
    hidden = np.asarray(hidden, dtype=np.float32)
    pt = np.asarray(patch_type).reshape(-1).astype(np.int64)
    order = order_by_patch(pt, patch_num)
    n = order.shape[0]
    sim = np.full(n, np.float32(IGNORE_TOKEN), np.float32)
    fragile = np.zeros(n, bool)
    lo = sim.copy()
    hi = sim.copy()
    if n < 2:
        return SimResult(sim, order, fragile, lo, hi)

    q = _sq_sums(hidden, order)                                   # exact |row|^2 per by-patch position
    d, a_abs = _row_sums(hidden, order[:-1], order[1:], dtype)

    with np.errstate(invalid="ignore"):
        slack_q = _SUM_SLACK * float(_U32) * q
        nrm = round_to(np.sqrt(q.astype(np.float32)), dtype)
        nrm_lo = round_to(np.sqrt(np.maximum(q - slack_q, 0).astype(np.float32)), dtype)
        nrm_hi = round_to(np.sqrt((q + slack_q).astype(np.float32)), dtype)
        slack_d = _SUM_SLACK * float(_U32) * a_abs
        dot = round_to(d.astype(np.float32), dtype)
        dot_lo = round_to((d - slack_d).astype(np.float32), dtype)
        dot_hi = round_to((d + slack_d).astype(np.float32), dtype)

    body = _chain(dot, nrm[:-1], nrm[1:], dtype)
    cands = []
    for dd in (dot_lo, dot_hi):
        for n1 in (nrm_lo[:-1], nrm_hi[:-1]):
            for n2 in (nrm_lo[1:], nrm_hi[1:]):
                cands.append(_chain(dd, n1, n2, dtype))
    cands = np.stack(cands)
    with np.errstate(invalid="ignore"):
        b_lo = np.nanmin(np.where(np.isnan(cands), np.inf, cands), axis=0).astype(np.float32)
        b_hi = np.nanmax(np.where(np.isnan(cands), -np.inf, cands), axis=0).astype(np.float32)
    if dtype == "f32":
        # nothing is rounded to a coarser grid: bracket by the float32 slack itself
        tol = np.float32(1e-5)
        b_lo, b_hi = body - tol, body + tol
        frag = np.ones(n - 1, bool)
    else:
        frag = (b_lo != b_hi) & ~np.isnan(body)

    same = pt[order[:-1]] == pt[order[1:]]
    sim[1:] = np.where(same, body, np.float32(IGNORE_TOKEN))
    lo[1:] = np.where(same, np.where(frag, b_lo, body), np.float32(IGNORE_TOKEN))
    hi[1:] = np.where(same, np.where(frag, b_hi, body), np.float32(IGNORE_TOKEN))
    fragile[1:] = same & frag
    return SimResult(sim, order, fragile, lo, hi)


def threshold_in_dtype(similarity_lower_bound: float, dtype: str) -> np.float32:
    """``sim >= python_float`` compares in the tensor dtype (wrapped-number rule); SURVEY H2."""
    return round_to(np.float32(similarity_lower_bound), dtype).reshape(())[()]


# ----------------------------------------------------------------------------------------------
# selection  (main.py:109-127)
# ----------------------------------------------------------------------------------------------
def topk_lowest_index(values: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k largest values, NaN ranked highest, ties at the k-th value broken by the
    LOWEST index.  Returned ascending.  (torch.topk leaves the tie choice unspecified — SURVEY H3;
    this is the rule the CUDA path implements, and a checker against torch compares tie-agnostically.)"""
    v = np.asarray(values, dtype=np.float32)
    if k <= 0:
        return np.zeros(0, np.int64)
    nan = np.isnan(v)
    key = np.where(nan, np.float32(0), v).astype(np.float64)
    # lexsort: last key is primary -> NaN first, then descending value, then ascending index
    rank = np.lexsort((np.arange(v.shape[0]), -key, ~nan))
    return np.sort(rank[:k]).astype(np.int64)


@dataclasses.dataclass
class Selection:
    merge_index: np.ndarray     # ascending by-patch positions to merge
    branch: str                 # "threshold" | "topk"
    ratio: float                # above_k_ratio
    count: int
    n_vis: int


def select_merge_index(sim: np.ndarray, patch_type: np.ndarray, similarity_lower_bound: float,
                       bound: float, dtype: str) -> Selection:
    pt = np.asarray(patch_type).reshape(-1)
    n_vis = int((pt != TEXT_TOKEN).sum())
    thr = threshold_in_dtype(similarity_lower_bound, dtype)
    with np.errstate(invalid="ignore"):
        idx = np.nonzero(sim >= thr)[0].astype(np.int64)
    ratio = idx.shape[0] / n_vis
    if ratio < bound:
        return Selection(idx, "threshold", ratio, int(idx.shape[0]), n_vis)
    k = int(bound * n_vis)
    return Selection(topk_lowest_index(sim, k), "topk", ratio, int(idx.shape[0]), n_vis)


# ----------------------------------------------------------------------------------------------
# merge + keep mask  (main.py:243-319)
# ----------------------------------------------------------------------------------------------
def merge_tokens_and_get_mask(hidden: np.ndarray, order: np.ndarray, merge_index: np.ndarray,
                              dtype: str) -> Tuple[np.ndarray, np.ndarray]:
    """Returns ``(hidden_after, keep_mask)``; ``hidden`` is not modified (the reference works in place)."""
    hidden = np.array(hidden, dtype=np.float32, copy=True)
    s_len = hidden.shape[0]
    keep = np.ones(s_len, bool)
    m = np.asarray(merge_index, dtype=np.int64)
    if m.shape[0] == 0:
        return hidden, keep
    n = order.shape[0]
    flag = np.zeros(n, bool)
    flag[m] = True
    keep[order[m]] = False
    pos = np.arange(n)
    anchor = np.maximum.accumulate(np.where(~flag, pos, -1))      # last unflagged position <= j
    src = hidden.copy()                                           # sources are gathered before the adds (:306-311)
    members = np.nonzero(flag)[0]
    off = members - anchor[members]                               # 1..L within the run
    run_len = np.zeros(n + 1, np.int64)                           # indexed by anchor position (+1 so that -1 -> slot n)
    np.add.at(run_len, anchor[members], 1)
    for t in range(1, int(off.max()) + 1):
        sel = members[off == t]
        arow = order[anchor[sel]]                                 # anchor -1 wraps to the last by-patch position
        hidden[arow] = round_to(hidden[arow] + src[order[sel]], dtype)
    anchors = np.unique(anchor[members])
    arow = order[anchors]
    # the divisor tensor is cast to T before the division (in-place op on a T tensor, main.py:314-317): exact up to 256
    # rows in bf16, T(301) = 300 beyond
    div = round_to((run_len[anchors] + 1).astype(np.float32), dtype)[:, None]
    hidden[arow] = round_to(hidden[arow] / div, dtype)
    return hidden, keep


# ----------------------------------------------------------------------------------------------
# importance  (utils.py:27-57) and prune selection (main.py:61-92)
# ----------------------------------------------------------------------------------------------
def last_query_attention(q: np.ndarray, k: np.ndarray, num: int, dtype: str, is_causal: bool = False,
                         scale: Optional[float] = None) -> np.ndarray:
    """q ``[Hq, S, D]``, k ``[Hk, S, D]`` (Hq % Hk == 0) -> probabilities ``[Hq, num, S]`` in T."""
    hq, s_len, d = q.shape
    hk = k.shape[0]
    g = hq // hk
    qs = q[:, -num:, :].astype(np.float64)
    scale_f = np.float32(1 / math.sqrt(d) if scale is None else scale)
    out = np.empty((hq, qs.shape[1], s_len), np.float32)
    for h in range(hq):
        kk = k[h // g].astype(np.float64)
        logits = round_to((qs[h] @ kk.T).astype(np.float32), dtype)          # bf16 matmul, fp32 accumulate
        logits = round_to(logits * round_to(scale_f, "f32"), dtype)          # * python scalar
        bias = np.zeros((qs.shape[1], s_len), np.float32)
        if is_causal:
            l_q = qs.shape[1]
            mask = np.triu(np.ones((l_q, s_len), bool), k=s_len - l_q + 1)
            bias[mask] = -np.inf
        logits = round_to(logits + bias, dtype)
        mx = logits.max(axis=1, keepdims=True)
        e = np.exp((logits - mx).astype(np.float64))
        out[h] = round_to((e / e.sum(axis=1, keepdims=True)).astype(np.float32), dtype)
    return out


def mean_heads(attn: np.ndarray, dtype: str) -> np.ndarray:
    """``torch.mean(attn, dim=(1,2))[0]``: ``T(sum_f32 / count)``, one rounding (main.py:69-70; probed on torch 2.11 CPU)."""
    hq, num, s_len = attn.shape
    tot = attn.astype(np.float64).sum(axis=(0, 1)).astype(np.float32)
    return round_to(tot / np.float32(hq * num), dtype)


def prune_keep_indices(importance: np.ndarray, start: int, length: int, q_len: int, ratio: float) -> np.ndarray:
    k = round(length * (1 - ratio))
    top = topk_lowest_index(importance[start:start + length], k) + start
    return np.sort(np.concatenate((np.arange(start), top, np.arange(start + length, q_len)))).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# the state machine  (main.py:8-140)
# ----------------------------------------------------------------------------------------------
def _to_int(x) -> int:
    return int(x.item()) if hasattr(x, "item") else int(x)


class OracleFrameFusion:
    """numpy mirror of ``FrameFusion`` (B = 1).  hidden ``[S,H]``; position container is a list of two
    arrays with the sequence on axis ``pos_axis`` or one 1-D/2-D array of position ids; mask ``[S,S]`` or None."""

    def __init__(self, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1, dtype="bf16"):
        self.cost = cost
        self.similarity_lower_bound = similarity_lower_bound
        self.ratio_lower_bound = ratio_lower_bound
        self.dtype = dtype

    def prepare(self, patch_type, patch_num, image_token_start_index, image_token_end_index,
                image_token_length, original_length, finish_merging=False, finish_pruning=False,
                sparsity_list=None):
        self.patch_type = np.asarray(patch_type).reshape(-1).astype(np.int64)
        self.patch_num = patch_num
        self.image_token_start_index = image_token_start_index
        self.image_token_end_index = image_token_end_index
        self.image_token_length = image_token_length
        self.original_length = original_length
        self.finish_merging = finish_merging
        self.finish_pruning = finish_pruning
        self.sparsity_list = [] if sparsity_list is None else sparsity_list
        self.trace: List[dict] = []
        self.last = None

    @staticmethod
    def _take(pos, idx_or_mask):
        if isinstance(pos, list):
            assert len(pos) == 2
            return [np.take(p, idx_or_mask, axis=p.ndim - 2) if idx_or_mask.dtype != bool
                    else np.compress(idx_or_mask, p, axis=p.ndim - 2) for p in pos]
        if isinstance(pos, np.ndarray):
            if pos.ndim != 2:
                raise NotImplementedError("Only support 2D position embeddings")
            return pos[:, idx_or_mask]
        raise NotImplementedError("Only support list or tensor for position embeddings")

    def forward(self, hidden, position_embeddings, attention_mask, self_attn_weights=None):
        q_len = hidden.shape[0]
        self.last = None
        if q_len > 1 and self.finish_merging and not self.finish_pruning:
            start = _to_int(self.image_token_start_index)
            length = _to_int(self.image_token_length - (self.original_length - q_len))
            imp = mean_heads(self_attn_weights, self.dtype)
            ratio = compute_pruning_ratio(self.sparsity_list, self.cost)
            keep = prune_keep_indices(imp, start, length, q_len, ratio)
            hidden = hidden[keep]
            position_embeddings = self._take(position_embeddings, keep)
            if attention_mask is not None:
                attention_mask = attention_mask[keep][:, keep]
            self.finish_pruning = True
            self.last = dict(stage="prune", keep=keep, importance=imp, start=start, length=length)
            self.trace.append(self.last)

        if q_len > 1 and not self.finish_merging:
            bound = compute_pruning_ratio(self.sparsity_list, self.cost)
            sr = similarity_by_patch(hidden, self.patch_type, self.patch_num, self.dtype)
            sel = select_merge_index(sr.sim, self.patch_type, self.similarity_lower_bound, bound, self.dtype)
            if sel.branch == "threshold":
                self.sparsity_list.append(sel.ratio)
                if sel.ratio < self.ratio_lower_bound:
                    self.finish_merging = True
            else:
                self.finish_merging = True
                self.finish_pruning = True
            hidden, keep_mask = merge_tokens_and_get_mask(hidden, sr.order, sel.merge_index, self.dtype)
            self.patch_type = self.patch_type[keep_mask]
            hidden = hidden[keep_mask]
            position_embeddings = self._take(position_embeddings, keep_mask)
            if attention_mask is not None:
                attention_mask = attention_mask[keep_mask][:, keep_mask]
            self.last = dict(stage="merge", sim=sr, sel=sel, keep_mask=keep_mask, sim_values=sr.sim, order=sr.order,
                             merge_index=sel.merge_index, branch=sel.branch)
            self.trace.append(self.last)
        return hidden, position_embeddings, attention_mask
