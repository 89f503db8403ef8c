"""``FrameFusion`` — the token-reduction operator, host side.

Mirror of ``/root/reference/framefusion/main.py`` (class ``FrameFusion``, lines 8-343): same constructor,
``prepare`` / ``forward`` / static helpers, same attributes, same exceptions.  The state machine and the
budget arithmetic stay in Python doubles exactly as the reference has them (main.py:109-127, 321-343); every
tensor operation is one call into ``libframefusion_b200.so`` through the C ABI (``include/framefusion_b200.h``):

* merge stage (main.py:104-138)  -> ``ff_build_links`` (first call of a prefill) + ``ff_merge_layer``
* prune stage (main.py:61-101)   -> ``ff_prune_layer``
* 4-D mask compaction            -> ``ff_compact_mask``

A reducing call synchronises the stream exactly once (to learn ``S_keep``); the reference needs >= 10 syncs.
There is no CPU / eager fallback: CPU tensors raise.

Differences from the reference, all invisible to its callers (which rebind every returned value,
models/qwen2/modeling_qwen2.py:46,67):
* ``forward`` does not write merged anchors back into the *input* ``hidden_states`` (the reference mutates it
  in place before compacting, main.py:304-317); the static ``merge_tokens_and_get_mask`` keeps the in-place
  contract.
* top-k ties are broken by the lowest by-patch index (``torch.topk`` leaves the choice unspecified).
* the returned tensors are views of buffers sized for the incoming sequence (only ``S_keep`` rows are used).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List

import numpy as np
import torch
from torch import nn

from . import _lib

TEXT_TOKEN = -1
IGNORE_TOKEN = -2

_DT = {torch.bfloat16: _lib.FF_BF16, torch.float16: _lib.FF_F16, torch.float32: _lib.FF_F32}


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"framefusion_b200 supports bfloat16 / float16 / float32 hidden states, got {t.dtype}")


_THR_CACHE = {}


def _threshold_in(value, dtype) -> float:
    """``sim >= python_scalar`` compares in the tensor dtype: the scalar rounded to T, as a Python float."""
    key = (float(value), dtype)
    v = _THR_CACHE.get(key)
    if v is None:
        v = _THR_CACHE[key] = torch.tensor(value, dtype=dtype).item()
    return v


def _stream(device) -> int:
    """Handle of torch's current stream on ``device`` (the raw getter: ``torch.cuda.current_stream`` builds a Python
    object per call, ~9 us, and a reducing call asks several times)."""
    index = device.index
    if index is None:
        index = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(index)


class _DeviceState:
    """Context + workspace of one FrameFusion object on one device."""

    def __init__(self, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        h = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(self.lib.ff_ctx_create(idx, C.byref(h)))
        self.ctx = h
        self.status = np.ctypeslib.as_array(self.lib.ff_ctx_status(h), shape=(_lib.ST_SLOTS,))
        self.ws = None
        self.ws_cap = (-1, -1)

    def workspace(self, seq_len: int, n_ids: int) -> torch.Tensor:
        if self.ws is None or seq_len > self.ws_cap[0] or n_ids > self.ws_cap[1]:
            cap = (max(seq_len, self.ws_cap[0]), max(n_ids, self.ws_cap[1]))
            nbytes = self.lib.ff_workspace_bytes(cap[0], cap[1])
            self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            self.ws_cap = cap
        return self.ws

    def ws_ptr(self):
        p = self.ws.data_ptr()
        a = (p + 255) // 256 * 256
        return a, self.ws.numel() - (a - p)

    def __del__(self):
        try:
            self.lib.ff_ctx_destroy(self.ctx)
        except Exception:
            pass


def _aux_of(t: torch.Tensor, seq_dim: int):
    """(src contiguous, dst buffer at full capacity, planes, plane stride bytes, row bytes) for a tensor that is
    compacted along ``seq_dim``; everything before seq_dim is a 'plane', everything after it a 'row'."""
    src = t.contiguous()
    planes = 1
    for d in src.shape[:seq_dim]:
        planes *= d
    row = src.element_size()
    for d in src.shape[seq_dim + 1:]:
        row *= d
    dst = torch.empty_like(src)
    stride = src.shape[seq_dim] * row
    return src, dst, planes, stride, row


class FrameFusion(nn.Module):
    # per-prefill / per-call state: plain Python attributes
    _PLAIN = frozenset((
        "cost", "similarity_lower_bound", "ratio_lower_bound", "patch_type", "patch_num", "image_token_start_index",
        "image_token_end_index", "image_token_length", "original_length", "finish_merging", "finish_pruning",
        "sparsity_list", "use_frame", "debug_trace", "last_trace", "kernel_events", "kernel_events_len", "_links_for", "_have_order",
        "_dev"))

    def __setattr__(self, name, value):
        # the operator has no parameters, buffers or sub-modules: skip nn.Module's bookkeeping (it costs ~4 us per
        # assignment, the isinstance check against nn.Parameter alone ~1.5 us, and forward() assigns a dozen state
        # attributes per call)
        if name in FrameFusion._PLAIN or not isinstance(value, (nn.Module, nn.Parameter)):
            object.__setattr__(self, name, value)
        else:
            super().__setattr__(name, value)

    def __init__(self, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1):
        super(FrameFusion, self).__init__()
        self.cost = cost
        self.similarity_lower_bound = similarity_lower_bound
        self.ratio_lower_bound = ratio_lower_bound
        self._dev = {}                  # torch.device -> _DeviceState
        self._links_for = None          # (patch_type tensor, its _version, device) the workspace links describe
        self._have_order = False        # the workspace holds the compact by-patch order (the multi-kernel path needs it)
        # the first merge call of a prefill on a uniform video runs as ONE launch (csrc/ff_frame.cuh: rows travel HBM ->
        # shared memory -> HBM once); the library checks the layout on the device and this class redoes the call on the
        # multi-kernel path if it says no.  False: never ask for it; "force": also on shapes where it is the slower one.
        self.use_frame = True
        self.debug_trace = False        # tests: keep what flowed between the stages of the last call
        self.last_trace = None
        self.kernel_events = None       # bench: a list collects (name, start, end) CUDA events around ff_* launches
        self.kernel_events_len = None   # ... of the merge calls whose sequence has this length only (None: of every merge call)

    # ---------------------------------------------------------------------------------------------
    def prepare(
        self,
        patch_type: torch.Tensor,
        patch_num: int,
        image_token_start_index: torch.Tensor,
        image_token_end_index: torch.Tensor,
        image_token_length: torch.Tensor,
        original_length: int,
        finish_merging: bool = False,
        finish_pruning: bool = False,
        sparsity_list: List[float] = None,
    ):
        """Per-prefill state reset (main.py:15-38)."""
        self.patch_type = patch_type
        self.patch_num = patch_num
        self.image_token_start_index = image_token_start_index
        self.image_token_end_index = image_token_end_index
        self.image_token_length = image_token_length
        self.original_length = original_length
        self.finish_merging = finish_merging
        self.finish_pruning = finish_pruning
        if sparsity_list is None:
            self.sparsity_list = []
        else:
            self.sparsity_list = sparsity_list
        self._links_for = None

    # ---------------------------------------------------------------------------------------------
    def _state(self, device) -> _DeviceState:
        st = self._dev.get(device)
        if st is None:
            st = self._dev[device] = _DeviceState(device)
        return st

    @staticmethod
    def _n_ids(patch_num) -> int:
        # torch.arange(patch_num) semantics: nvila passes a float, minicpmv a 0-d tensor
        if isinstance(patch_num, torch.Tensor):
            patch_num = patch_num.item()
        return int(math.ceil(float(patch_num)))

    def _ensure_links(self, st: _DeviceState, q_len: int, need_order: bool = False, next_row_bytes: int = 0, next_flags: int = 0):
        """Links of the current ``patch_type`` in the workspace.  ``next_row_bytes`` / ``next_flags`` describe the merge call
        that follows: if the library will serve it with the frame-pipelined kernel it only counts the tokens and checks the
        layout here (``ff_build_links_for``: one small kernel instead of the counting sort's four)."""
        pt = self.patch_type
        key = self._links_for
        if key is not None and key[0] is pt and key[1] == pt._version and key[2] == st.device \
                and (self._have_order or not need_order):
            return
        if pt.numel() != q_len:
            raise RuntimeError(f"patch_type has {pt.numel()} entries for a sequence of {q_len} tokens")
        n_ids = self._n_ids(self.patch_num)
        st.workspace(q_len, n_ids)
        ptc = pt.reshape(-1).to(torch.int64).contiguous()
        wp, wb = st.ws_ptr()
        _lib.check(st.lib.ff_build_links_for(st.ctx, wp, wb, ptc.data_ptr(), q_len, n_ids, next_row_bytes, next_flags,
                                             _stream(st.device)))
        self._links_for = (pt, pt._version, st.device)
        self._have_order = True      # (or the library keeps what its own next kernel needs: it refuses anything else)

    def _pos_aux(self, position_embeddings, auxes):
        """Registers the position container's tensors for compaction; returns a closure that rebuilds it."""
        if type(position_embeddings) == list:
            assert len(position_embeddings) == 2
            seq_dim = 2 if position_embeddings[0].ndim == 4 else 1
            i0 = len(auxes)
            auxes.append(_aux_of(position_embeddings[0], seq_dim) + (seq_dim,))
            auxes.append(_aux_of(position_embeddings[1], seq_dim) + (seq_dim,))

            def rebuild(outs):
                position_embeddings[0] = outs[i0]
                position_embeddings[1] = outs[i0 + 1]
                return position_embeddings
            return rebuild
        elif type(position_embeddings) == torch.Tensor:
            if position_embeddings.ndim == 2:
                i0 = len(auxes)
                auxes.append(_aux_of(position_embeddings, 1) + (1,))
                return lambda outs: outs[i0]
            else:
                raise NotImplementedError("Only support 2D position embeddings")
        else:
            raise NotImplementedError("Only support list or tensor for position embeddings")

    @staticmethod
    def _pack_aux(auxes):
        arr = (_lib.FFAux * max(len(auxes), 1))()
        for i, (src, dst, planes, stride, row, _sd) in enumerate(auxes):
            arr[i].src = src.data_ptr()
            arr[i].dst = dst.data_ptr()
            arr[i].planes = planes
            arr[i].src_plane_stride = stride
            arr[i].dst_plane_stride = stride
            arr[i].row_bytes = row
        return arr

    @staticmethod
    def _narrow(auxes, s_keep):
        outs = []
        for (_src, dst, planes, _stride, _row, seq_dim) in auxes:
            v = dst.narrow(seq_dim, 0, s_keep)
            outs.append(v if planes == 1 else v.contiguous())
        return outs

    def _compact_mask(self, st, attention_mask, q_len, s_keep):
        if attention_mask.ndim != 4 or attention_mask.shape[-1] != q_len or attention_mask.shape[-2] != q_len \
                or attention_mask.shape[0] != 1 or attention_mask.shape[1] != 1:
            raise NotImplementedError("attention_mask must be [1, 1, S, S]")
        m = attention_mask.contiguous()
        out = torch.empty((1, 1, s_keep, s_keep), dtype=m.dtype, device=m.device)
        wp, wb = st.ws_ptr()
        _lib.check(st.lib.ff_compact_mask(st.ctx, wp, wb, m.data_ptr(), out.data_ptr(), q_len, s_keep,
                                          m.element_size(), _stream(st.device)))
        return out

    # ---------------------------------------------------------------------------------------------
    def forward(self, hidden_states, position_embeddings, attention_mask, self_attn_weights=None):
        """Same contract as the reference forward (main.py:40-140): returns
        ``(hidden_states, position_embeddings, attention_mask)`` after at most one prune and one merge stage."""
        bsz, q_len, hidden_size = hidden_states.size()
        device = hidden_states.device
        self.last_trace = None

        # pruning (main.py:61-101)
        if q_len > 1 and self.finish_merging == True and self.finish_pruning == False:
            hidden_states, position_embeddings, attention_mask = self._prune(
                hidden_states, position_embeddings, attention_mask, self_attn_weights)
            self.finish_pruning = True

        # merging (main.py:104-138)
        if q_len > 1 and (not self.finish_merging):
            hidden_states, position_embeddings, attention_mask = self._merge(
                hidden_states, position_embeddings, attention_mask)

        return hidden_states, position_embeddings, attention_mask

    # ---------------------------------------------------------------------------------------------
    def _require_cuda(self, t):
        if not t.is_cuda:
            raise RuntimeError("framefusion_b200 runs on CUDA tensors only (there is no CPU fallback)")

    def _prune(self, hidden_states, position_embeddings, attention_mask, self_attn_weights, pruning_ratio=None):
        """One prune stage.  ``pruning_ratio`` None: the budget formula (main.py:76-78); a number: that fraction of the
        vision span goes (the fixed-ratio baselines, ``baselines.py``)."""
        self._require_cuda(hidden_states)
        bsz, q_len, hidden_size = hidden_states.size()
        assert bsz == 1, "Only support batch size 1"
        device = hidden_states.device
        st = self._state(device)

        def to_int(x):
            return x.item() if isinstance(x, torch.Tensor) else int(x)
        start = to_int(self.image_token_start_index)
        length = to_int(self.image_token_length - (self.original_length - q_len))
        if self_attn_weights is None:
            raise TypeError("the prune stage needs self_attn_weights (last-query attention probabilities)")
        attn = self_attn_weights[0]
        if attn.shape[-1] != q_len:
            raise RuntimeError(f"self_attn_weights covers {attn.shape[-1]} keys, hidden_states has {q_len} tokens")
        attn = attn.reshape(-1, q_len).to(hidden_states.dtype).contiguous()
        if pruning_ratio is None:
            pruning_ratio = self._compute_pruning_ratio(self.sparsity_list, self.cost)
        k = round(length * (1 - pruning_ratio))
        if k < 0 or k > length or start < 0 or start + length > q_len:
            raise RuntimeError(f"selected index k out of range (k={k}, vision span [{start}, {start + length}) of {q_len})")

        auxes = []
        rebuild = self._pos_aux(position_embeddings, auxes)
        hidden = hidden_states.contiguous()
        out = torch.empty_like(hidden)
        imp = torch.empty(q_len, dtype=hidden.dtype, device=device) if self.debug_trace else None
        st.workspace(q_len, 0)                  # only ever grows (a layer-split model may prune on another device)
        wp, wb = st.ws_ptr()
        stream = _stream(device)
        _lib.check(st.lib.ff_prune_layer(
            st.ctx, wp, wb, attn.data_ptr(), attn.shape[0], hidden.data_ptr(), out.data_ptr(), _dtype_code(hidden),
            q_len, hidden_size, start, length, k, self._pack_aux(auxes), len(auxes),
            imp.data_ptr() if imp is not None else None, stream))
        # the selection kernel writes S_keep before the gather runs: the host goes on while the rows move
        _lib.check(st.lib.ff_status_wait(st.ctx, stream))
        if int(st.status[_lib.ST_INTERNAL]) != 0:
            st.status[_lib.ST_INTERNAL] = 0
            raise _lib.FFError("framefusion_b200: a wait inside the single-launch merge kernel of an earlier call timed out")
        s_keep = int(st.status[_lib.ST_SEQ_KEEP])
        outs = self._narrow(auxes, s_keep)
        position_embeddings = rebuild(outs)
        hidden_states = out.narrow(1, 0, s_keep)
        if attention_mask != None:
            attention_mask = self._compact_mask(st, attention_mask, q_len, s_keep)
        if self.debug_trace:
            keep = torch.empty(q_len, dtype=torch.uint8, device=device)
            _lib.check(st.lib.ff_debug_read(st.ctx, wp, wb, 0, keep.data_ptr(), q_len, 0, _stream(device)))
            self.last_trace = dict(stage="prune", keep=np.nonzero(keep.cpu().numpy())[0], importance=imp.float().cpu().numpy(),
                                   start=start, length=length)
        return hidden_states, position_embeddings, attention_mask

    def _merge(self, hidden_states, position_embeddings, attention_mask, fixed_sparsity=None):
        """One merge stage.  ``fixed_sparsity`` None: threshold / budget as in main.py:109-127; a number s: exactly the
        ``int(s * n_vis)`` most similar tokens merge (the top-k branch with a given k — the fixed-sparsity baseline,
        ``baselines.py``): the threshold sits below the chain-head sentinel, so every token counts and the device takes the
        top-k branch with ``bound = s``; the operator's own state (flags, ``sparsity_list``) is left alone."""
        self._require_cuda(hidden_states)
        bsz, q_len, hidden_size = hidden_states.size()
        assert bsz == 1, "Only support batch size 1"
        device = hidden_states.device
        st = self._state(device)

        # align devices (main.py:106)
        self.patch_type = self.patch_type.to(device)
        fixed = fixed_sparsity is not None
        sparsity_upper_bound = float(fixed_sparsity) if fixed else self._compute_pruning_ratio(self.sparsity_list, self.cost)
        flags = (0 if self.use_frame else 2) | (4 if self.use_frame == "force" else 0)
        if fixed:
            flags = 2                                                        # the top-k branch lives on the multi-kernel path
        dt = hidden_states.dtype
        thr = -3.0 if fixed else _threshold_in(self.similarity_lower_bound, dt)   # the scalar is compared in T
        # tell the library what call follows: if it is going to be the frame-pipelined kernel, the counting sort of the
        # links is not needed (debug_trace reads the by-patch order back, and a threshold at the sentinel never takes it)
        lite_row_bytes = 0 if (self.debug_trace or (flags & 2) or not thr > -2.0) else hidden_size * hidden_states.element_size()
        self._ensure_links(st, q_len, need_order=True, next_row_bytes=lite_row_bytes, next_flags=flags)
        hidden = hidden_states.contiguous()
        out = torch.empty_like(hidden)
        auxes = [_aux_of(self.patch_type.reshape(1, -1).to(torch.int64), 1) + (1,)]
        rebuild = self._pos_aux(position_embeddings, auxes)
        packed = self._pack_aux(auxes)
        wp, wb = st.ws_ptr()
        stream = _stream(device)
        code = _dtype_code(hidden)

        def launch(flags):
            ev = self.kernel_events
            if ev is not None and self.kernel_events_len not in (None, q_len):
                ev = None                                                    # only calls on a sequence of that length are timed
            if ev is not None:
                # the library records the two events right around its own launches (ff_ctx_timing): GPU time of the
                # call's kernels, without the host's way to the first launch
                self.reserve_kernel_events(len(ev) + 1)
                e0, e1 = self._event_pool[len(ev)]                           # the n-th timed call of a list uses the n-th pair
                _lib.check(st.lib.ff_ctx_timing(st.ctx, e0.cuda_event, e1.cuda_event))
            try:
                _lib.check(st.lib.ff_merge_layer(st.ctx, wp, wb, hidden.data_ptr(), out.data_ptr(), code, q_len, hidden_size,
                                                 thr, float(sparsity_upper_bound), packed, len(auxes), flags, stream))
            finally:
                if ev is not None:
                    st.lib.ff_ctx_timing(st.ctx, None, None)
            if ev is not None:
                ev.append(("ff_merge_layer", q_len, e0, e1))
            # the deciding kernel writes the status block before the rows have all moved (the scan kernel before the gather
            # runs, the frame-pipelined kernel when its last frame is decided), and the host goes on meanwhile: everything it
            # enqueues next is ordered behind them
            _lib.check(st.lib.ff_status_wait(st.ctx, stream))

        try:
            launch(flags)
        except ValueError:
            if not lite_row_bytes:
                raise
            # the links were left for the frame-pipelined kernel only and the library cannot take it for these tensors
            # (alignment): build them in full and let it choose again
            lite_row_bytes = 0
            self._links_for = None
            self._ensure_links(st, q_len, need_order=True)
            launch(flags)
        status = st.status
        ran_frame = int(status[_lib.ST_FUSED]) == 2        # the library took the frame-pipelined kernel (first call of a prefill)
        if int(status[_lib.ST_INTERNAL]) != 0:
            # (the frame-pipelined kernel publishes the status block before its last rows are out: a wait that gave up
            # after that shows here at the latest at the next call)
            status[_lib.ST_INTERNAL] = 0
            raise _lib.FFError("framefusion_b200: a wait inside the single-launch merge kernel timed out")
        if ran_frame and int(status[_lib.ST_ERROR]) == 3:
            # the single-launch kernel speculates on the threshold branch and on a uniform video layout; the device says
            # otherwise: redo with the multi-kernel path (the input is untouched)
            self._links_for = None
            self._ensure_links(st, q_len, need_order=True)
            launch(2)
            ran_frame = False
        err = int(status[_lib.ST_ERROR])
        if err == 1:
            raise ZeroDivisionError("division by zero")                      # frame_token_num == 0 (main.py:114)
        if err == 2:
            raise RuntimeError("selected index k out of range")              # torch.topk's complaint (main.py:122)
        count, frame_token_num = int(status[_lib.ST_COUNT]), int(status[_lib.ST_NVIS])
        s_keep, branch = int(status[_lib.ST_SEQ_KEEP]), int(status[_lib.ST_BRANCH])
        above_k_ratio = count / frame_token_num
        assert (above_k_ratio < sparsity_upper_bound) == (branch == 0), "device / host branch decision disagree"
        if fixed:
            pass                                                             # a given k: no budget bookkeeping
        elif above_k_ratio < sparsity_upper_bound:
            self.sparsity_list.append(above_k_ratio)
            if above_k_ratio < self.ratio_lower_bound:
                self.finish_merging = True
        else:
            self.finish_merging = True
            self.finish_pruning = True

        self._have_order = True
        if self.debug_trace:
            self._record_merge_trace(st, hidden, q_len, int(status[_lib.ST_NCHAIN]), branch)

        if not ran_frame and int(status[_lib.ST_NMERGED]) == 0:
            # nothing was merged: the sequence is unchanged and the gather kernel did not run — hand the inputs back
            # (the reference returns copies with the same values; its callers rebind them, modeling_qwen2.py:46,67)
            self._links_for = (self.patch_type, self.patch_type._version, device)
            return hidden_states, position_embeddings, attention_mask

        outs = self._narrow(auxes, s_keep)
        self.patch_type = outs[0].reshape(bsz, -1)
        self._links_for = (self.patch_type, self.patch_type._version, device)
        hidden_states = out.narrow(1, 0, s_keep)
        position_embeddings = rebuild(outs)
        if attention_mask is not None:
            attention_mask = self._compact_mask(st, attention_mask, q_len, s_keep)
        return hidden_states, position_embeddings, attention_mask

    def reserve_kernel_events(self, n: int):
        """Create the CUDA event pairs the ``kernel_events`` hook hands to ``ff_ctx_timing`` ahead of time (an event gets
        its handle when it is first recorded), so that a timed region pays nothing for them."""
        pool = self.__dict__.setdefault("_event_pool", [])
        while len(pool) < n:
            pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            pair[0].record()
            pair[1].record()
            pool.append(pair)

    def _record_merge_trace(self, st, hidden, q_len, n_chain, branch):
        device = hidden.device
        wp, wb = st.ws_ptr()
        stream = _stream(device)
        keep = torch.empty(q_len, dtype=torch.uint8, device=device)
        flags = torch.empty(max(n_chain, 1), dtype=torch.uint8, device=device)
        sim = torch.empty(max(n_chain, 1), dtype=hidden.dtype, device=device)
        order = torch.empty(max(n_chain, 1), dtype=torch.int64, device=device)
        _lib.check(st.lib.ff_debug_read(st.ctx, wp, wb, 0, keep.data_ptr(), q_len, 0, stream))
        _lib.check(st.lib.ff_debug_read(st.ctx, wp, wb, 1, flags.data_ptr(), n_chain, 0, stream))
        _lib.check(st.lib.ff_debug_read(st.ctx, wp, wb, 2, sim.data_ptr(), n_chain, _dtype_code(hidden), stream))
        _lib.check(st.lib.ff_debug_read(st.ctx, wp, wb, 3, order.data_ptr(), n_chain, 0, stream))
        torch.cuda.current_stream(device).synchronize()
        self.last_trace = dict(
            stage="merge", branch="topk" if branch else "threshold",
            keep_mask=keep[:q_len].cpu().numpy().astype(bool),
            merge_index=np.nonzero(flags[:n_chain].cpu().numpy())[0],
            sim_values=sim[:n_chain].float().cpu().numpy(),
            order=order[:n_chain].cpu().numpy())

    # ---------------------------------------------------------------------------------------------
    # static helpers with the reference's signatures (main.py:180-343)
    # ---------------------------------------------------------------------------------------------
    _static_state = {}

    @classmethod
    def _static(cls, device) -> _DeviceState:
        st = cls._static_state.get(device)
        if st is None:
            st = cls._static_state[device] = _DeviceState(device)
        return st

    @staticmethod
    def compute_similarity_and_token_index_by_patch(hidden_states, token_patch_type, patch_num):
        """(similarity_by_patch [1,N] in the hidden dtype, token_index_by_patch [1,N] int64) — main.py:180-241."""
        bsz, q_len, hidden_size = hidden_states.size()
        device = hidden_states.device
        assert bsz == 1, "Only support batch size 1"
        if not hidden_states.is_cuda:
            raise RuntimeError("framefusion_b200 runs on CUDA tensors only (there is no CPU fallback)")
        st = FrameFusion._static(device)
        n_ids = FrameFusion._n_ids(patch_num)
        st.workspace(q_len, n_ids)
        wp, wb = st.ws_ptr()
        stream = _stream(device)
        pt = token_patch_type.to(device).reshape(-1).to(torch.int64).contiguous()
        _lib.check(st.lib.ff_build_links(st.ctx, wp, wb, pt.data_ptr(), q_len, n_ids, stream))
        torch.cuda.current_stream(device).synchronize()
        n = int(st.status[_lib.ST_NCHAIN])
        sim = torch.empty((bsz, n), dtype=hidden_states.dtype, device=device)
        order = torch.empty((bsz, n), dtype=torch.int64, device=device)
        hidden = hidden_states.contiguous()
        _lib.check(st.lib.ff_similarity(st.ctx, wp, wb, hidden.data_ptr(), _dtype_code(hidden), q_len, hidden_size,
                                        2.0, sim.data_ptr(), order.data_ptr(), stream))
        return sim, order

    @staticmethod
    def merge_tokens_and_get_mask(hidden_states: torch.Tensor, similarity_by_patch, token_index_by_patch, merge_index_by_patch):
        """In-place merge of ``hidden_states`` + keep mask ``[1,S]`` bool — main.py:243-319."""
        device = hidden_states.device
        if merge_index_by_patch.shape[0] == 0:
            keep_mask = torch.ones(hidden_states.shape[:-1], dtype=torch.bool, device=device)
            return hidden_states, keep_mask
        if not hidden_states.is_cuda:
            raise RuntimeError("framefusion_b200 runs on CUDA tensors only (there is no CPU fallback)")
        bsz, q_len, hidden_size = hidden_states.size()
        assert bsz == 1, "Only support batch size 1"
        if not hidden_states.is_contiguous():
            raise RuntimeError("merge_tokens_and_get_mask works in place and needs contiguous hidden_states")
        st = FrameFusion._static(device)
        st.workspace(q_len, 0)
        wp, wb = st.ws_ptr()
        order = token_index_by_patch.reshape(-1).to(device=device, dtype=torch.int64).contiguous()
        idx = merge_index_by_patch.reshape(-1).to(device=device, dtype=torch.int64).contiguous()
        keep = torch.empty(q_len, dtype=torch.uint8, device=device)
        _lib.check(st.lib.ff_merge_apply(st.ctx, wp, wb, hidden_states.data_ptr(), _dtype_code(hidden_states), q_len,
                                         hidden_size, order.data_ptr(), order.numel(), idx.data_ptr(), idx.numel(),
                                         keep.data_ptr(), _stream(device)))
        return hidden_states, keep.to(torch.bool).reshape(bsz, q_len)

    @staticmethod
    def _compute_pruning_ratio(sparsity_list, cost, num_layers=28):
        """Budget formula in host doubles (main.py:321-343): what fraction of the tokens still has to go so that the
        remaining layers fit into ``num_layers * cost`` token-layers, given the fractions merged so far."""
        alive, spent = 1, 0
        for merged in sparsity_list:
            alive *= (1 - merged)
            spent += alive
        budget_left = num_layers * cost - spent
        if budget_left < 0:
            raise ValueError("The cost is too small")
        share = budget_left / ((num_layers - len(sparsity_list)) * alive)
        return 0 if share > 1 else 1 - share


def cosine_similarity(mat1, mat2):
    """Exported helper of the reference (main.py:345-349), in torch ops; the operator itself uses the kernels."""
    return (mat1 * mat2).sum(dim=-1) / (mat1.norm(dim=-1) * mat2.norm(dim=-1))


def find_contigious_latter_index(index_tensor: torch.LongTensor) -> torch.Tensor:
    """Run lengths at the last element of every run of ones, zeros elsewhere (main.py:351-380);
    ``[0,1,1,1,0,0,1,1] -> [0,0,0,3,0,0,0,2]``.  Exported helper; the kernels carry runs implicitly."""
    bsz, n = index_tensor.shape
    ones = index_tensor == 1
    idx = torch.arange(n, device=index_tensor.device).expand(bsz, n)
    # position of the most recent non-one element, via a running maximum
    last_zero = torch.where(ones, torch.full_like(idx, -1), idx).cummax(dim=1).values
    run = (idx - last_zero).to(index_tensor.dtype)
    nxt = torch.cat([ones[:, 1:], torch.zeros((bsz, 1), dtype=torch.bool, device=index_tensor.device)], dim=1)
    return torch.where(ones & ~nxt, run, torch.zeros_like(run))
