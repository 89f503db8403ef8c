"""Shared parity harness: drives an implementation through the call sequence of a golden fixture
(``tests/golden/case_*.npz``, produced from the unmodified reference by ``oracle/gen_golden.py``) and
compares every recorded quantity.  Used by the CPU tests (implementation = the numpy oracle) and by the
GPU tests (implementation = the CUDA path through the C-ABI)."""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from framefusion_b200 import synth  # noqa: E402
from oracle import ff_oracle as orc  # noqa: E402
from oracle.gen_golden import (DT, build_inputs, input_checksum, raw_bits, row_checksums,  # noqa: E402
                               tensor_checksum)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz"))
    spec = json.loads(str(z["spec"]))
    return z, spec


def t2f(t: torch.Tensor) -> np.ndarray:
    """torch tensor of dtype T -> float32 ndarray (exact)."""
    return t.detach().float().cpu().numpy()


def f2t(a: np.ndarray, dtype: str) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DT[dtype])


class OracleAdapter:
    """Presents ``oracle.ff_oracle.OracleFrameFusion`` with the torch-tensor interface of ``FrameFusion``."""

    def __init__(self, cost, slb, rlb, dtype):
        self.o = orc.OracleFrameFusion(cost, slb, rlb, dtype)
        self.dtype = dtype

    def prepare(self, patch_type, patch_num, start, end, length, original_length):
        self.o.prepare(patch_type.numpy(), patch_num, start, end, length, original_length)

    def __call__(self, hidden, pos, mask, attn=None):
        dt = self.dtype
        if isinstance(pos, list):
            p = [t2f(x) for x in pos]
            p = [x[0] if x.ndim == 3 else x for x in p]            # [S,D] or [3,1,S,D]
        else:
            p = pos.numpy()
        m = None if mask is None else t2f(mask[0, 0])
        a = None if attn is None else t2f(attn[0])
        h, p, m = self.o.forward(t2f(hidden[0]), p, m, a)
        hidden = f2t(h, dt)[None]
        if isinstance(pos, list):
            pos[0] = f2t(p[0], dt) if p[0].ndim == 4 else f2t(p[0], dt)[None]
            pos[1] = f2t(p[1], dt) if p[1].ndim == 4 else f2t(p[1], dt)[None]
        else:
            pos = torch.from_numpy(p)
        if m is not None:
            mask = f2t(m, dt)[None, None]
        return hidden, pos, mask

    # introspection used by the harness
    @property
    def finish_merging(self): return self.o.finish_merging
    @property
    def finish_pruning(self): return self.o.finish_pruning
    @property
    def sparsity_list(self): return self.o.sparsity_list
    @property
    def patch_type(self): return torch.from_numpy(self.o.patch_type)[None]
    def last_trace(self): return self.o.last


def same_selection_modulo_ties(values: np.ndarray, got: np.ndarray, want: np.ndarray) -> bool:
    """Two top-k index sets agree up to the (unspecified) choice among elements equal to the k-th value."""
    if got.shape != want.shape:
        return False
    if got.shape[0] == 0:
        return True
    v = np.where(np.isnan(values), np.inf, values)
    kth = np.sort(v[want])[0]
    sg, sw = set(got.tolist()), set(want.tolist())
    diff = sg ^ sw
    if any(v[i] != kth for i in diff):
        return False
    must = set(np.nonzero(v > kth)[0].tolist())
    return must <= sg and must <= sw


def run_and_compare(name, make_impl, device="cpu", check_sim_with_oracle=True):
    """make_impl(cost, slb, rlb, dtype) -> object with prepare/__call__/flags; returns a small report dict."""
    z, spec = load_case(name)
    wl, pos, mask, dtype = build_inputs(spec)
    assert np.array_equal(input_checksum(wl, pos, mask), z["input_checksum"]), \
        "regenerated inputs differ from the ones the golden fixture was made from"
    num = spec.get("num", 1)
    impl = make_impl(spec["cost"], spec["slb"], spec["rlb"], dtype)
    args = list(wl.prepare_args())
    if spec.get("patch_num_float"):
        args[1] = float(args[1])
    args[0] = args[0].to(device)
    impl.prepare(*args)
    hidden = wl.hidden.clone().to(device)
    pos = [p.to(device) for p in pos] if isinstance(pos, list) else pos.to(device)
    mask = None if mask is None else mask.to(device)
    report = dict(fragile=0, sim_mismatch_in_fragile=0, n_sim=0)
    for c in range(int(z["n_calls"])):
        p = f"c{c}_"
        attn = None
        if c > 0 and spec.get("drift"):
            hidden = synth.apply_drift(hidden, spec["drift"], spec["wl"]["seed"], c)
        if impl.finish_merging and not impl.finish_pruning:
            attn = synth.make_attention_row(hidden.shape[1], n_heads=28, num=num, dtype=DT[dtype],
                                            seed=wl.hidden.shape[1] + c).to(device)
        hidden_in = hidden.clone()
        pt_in = impl.patch_type.clone()
        hidden, pos, mask = impl(hidden, pos, mask, attn)
        stage = str(z[p + "stage"])
        tr = impl.last_trace()
        got_stage = "none" if tr is None else tr["stage"]
        assert got_stage == stage, f"call {c}: stage {got_stage!r}, reference did {stage!r}"
        if stage == "none":
            assert hidden.shape[1] == int(z[p + "seq_len"])
        if "merge" in stage:
            assert tr is not None and "merge" in got_stage, f"call {c}: expected a merge stage"
            g_sim = orc.bits_to_f32(z[p + "sim"], dtype)
            sim = np.asarray(tr["sim_values"], dtype=np.float32)
            order = np.asarray(tr["order"])
            assert np.array_equal(order, z[p + "order"]), f"call {c}: by-patch order differs"
            # similarity: bit-equal unless a float32 summation order can change it (oracle brackets)
            sr = orc.similarity_by_patch(t2f(hidden_in[0]), pt_in.cpu().numpy().reshape(-1), args[1], dtype)
            for label, val in (("golden", g_sim), ("impl", sim)):
                neq = ~((val == sr.sim) | (np.isnan(val) & np.isnan(sr.sim)))
                bad = neq & ~sr.fragile
                assert not bad.any(), f"call {c}: {label} sim differs from oracle outside the fragile set at {np.nonzero(bad)[0][:8]}"
                inb = (val >= sr.lo) & (val <= sr.hi)
                assert inb[neq].all(), f"call {c}: {label} sim outside the oracle bracket"
            report["fragile"] += int(sr.fragile.sum()) if dtype != "f32" else 0
            report["sim_mismatch_in_fragile"] += int((~(sim == g_sim)).sum())
            report["n_sim"] += sim.shape[0]
            mi = np.asarray(tr["merge_index"])
            g_mi = z[p + "merge_index"]
            if tr["branch"] == "topk":
                assert same_selection_modulo_ties(g_sim, mi, g_mi), f"call {c}: top-k selection differs beyond ties"
                report["topk_tie_diff"] = int(len(set(mi.tolist()) ^ set(g_mi.tolist())))
            else:
                assert np.array_equal(mi, g_mi), f"call {c}: merge index differs"
            keep = np.unpackbits(z[p + "keep_mask"])[: hidden_in.shape[1]].astype(bool)
            if tr["branch"] != "topk" or report.get("topk_tie_diff", 0) == 0:
                assert np.array_equal(np.asarray(tr["keep_mask"]), keep), f"call {c}: keep mask differs"
        if "prune" in stage:
            assert "prune" in got_stage, f"call {c}: expected a prune stage"
            trp = tr
            imp = np.asarray(trp["importance"], dtype=np.float32)
            got_keep = np.asarray(trp["keep"])
            want_keep = z[p + "prune_keep"]
            assert got_keep.shape == want_keep.shape
            if not np.array_equal(got_keep, want_keep):
                st, ln = trp["start"], trp["length"]
                in_span = lambda k: k[(k >= st) & (k < st + ln)] - st
                assert np.array_equal(got_keep[got_keep < st], want_keep[want_keep < st])
                assert np.array_equal(got_keep[got_keep >= st + ln], want_keep[want_keep >= st + ln])
                assert same_selection_modulo_ties(imp[st:st + ln], in_span(got_keep), in_span(want_keep)), \
                    f"call {c}: prune selection differs beyond ties"
                report["prune_tie_diff"] = int(len(set(got_keep.tolist()) ^ set(want_keep.tolist())))
        tie_diff = report.get("topk_tie_diff", 0) + report.get("prune_tie_diff", 0)
        assert hidden.shape[1] == int(z[p + "seq_len"]), f"call {c}: sequence length {hidden.shape[1]} != {int(z[p + 'seq_len'])}"
        assert [bool(impl.finish_merging), bool(impl.finish_pruning)] == z[p + "flags"].tolist(), f"call {c}: flags"
        assert np.array_equal(np.array(impl.sparsity_list, dtype=np.float64), z[p + "sparsity_list"]), f"call {c}: sparsity_list"
        if tie_diff == 0:
            assert np.array_equal(row_checksums(hidden[0]), z[p + "hidden_rows"]), f"call {c}: hidden rows differ"
            plist = pos if isinstance(pos, list) else [pos]
            for i, t in enumerate(plist):
                assert list(t.shape) == z[p + f"pos{i}_shape"].tolist(), f"call {c}: pos{i} shape"
                assert np.array_equal(tensor_checksum(t.reshape(-1, t.shape[-1])), z[p + f"pos{i}"]), f"call {c}: pos{i}"
            if mask is not None:
                assert np.array_equal(tensor_checksum(mask[0, 0]), z[p + "mask"]), f"call {c}: mask"
            assert np.array_equal(impl.patch_type[0].cpu().numpy().astype(np.int32), z[p + "patch_type"]), f"call {c}: patch_type"
    return report


# merge-stage kernel choice of framefusion_b200.main.FrameFusion: "frame" = the first merge call of a prefill on a uniform
# video runs the frame-pipelined kernel (forced: the tests' shapes are mostly ones the library would not pick it for),
# everything else the multi-kernel path; "multi" = the multi-kernel path for every call
MODES = ["frame", "multi"]


def set_mode(ff, mode):
    ff.use_frame = "force" if mode == "frame" else False
