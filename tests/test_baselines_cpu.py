"""CPU: the fixed-amount comparison methods (FastV, fixed-sparsity merging, the combinations) — the meta interface and the
hook plumbing on a tiny random-init decoder, with the numpy oracle standing in for the CUDA operator (which raises on CPU
tensors).  Reference: models/qwen2/modeling_qwen2_baseline.py:26-39, 45-109, 175-188, 300-342, 860-874, 916-920, 1339-1355,
2055-2069."""
import math

import numpy as np
import pytest
import torch

from _harness import f2t, t2f
from oracle import ff_oracle as orc
from oracle import ff_torch_port as port
from oracle.ff_baselines_oracle import OracleBaseline
from framefusion_b200 import synth
from framefusion_b200.baselines import TokenReductionBaseline, compute_density_overhead
from framefusion_b200.hooks import qwen2_baselines as qb
from test_hooks_cpu import tiny_model, workload


class OracleBaselineOperator(torch.nn.Module):
    """``OracleBaseline`` behind the torch-tensor surface the hooks use."""

    def __init__(self, sparsity=None, fastv_k=None, fastv_r=0.5):
        super().__init__()
        object.__setattr__(self, "o", OracleBaseline(sparsity, fastv_k, fastv_r, "f32"))
        self.calls = []
        self._keep = None

    fastv_k = property(lambda s: s.o.fastv_k)

    def prepare(self, pt, pn, start, end, length, original_length):
        self.o.prepare(pt.numpy(), pn, start, end, length, original_length)

    def wants_attention(self, i):
        return self.o.wants_attention(i)

    def _wrap(self, fn, i, hidden, pos, mask, *extra):
        h, p, m = fn(i, t2f(hidden[0]), [t2f(pos[0][0]), t2f(pos[1][0])], None if mask is None else t2f(mask[0, 0]), *extra)
        pos[0], pos[1] = f2t(p[0], "f32")[None], f2t(p[1], "f32")[None]
        return f2t(h, "f32")[None], pos, None if m is None else f2t(m, "f32")[None, None]

    def merge_at(self, i, hidden, pos, mask):
        out = self._wrap(self.o.merge_at, i, hidden, pos, mask)
        self.calls.append(("merge", i, hidden.shape[1], out[0].shape[1]))
        return out

    def fastv_at(self, i, hidden, pos, mask, attn):
        out = self._wrap(self.o.fastv_at, i, hidden, pos, mask, t2f(attn[0]))
        self._keep = torch.from_numpy(self.o.last["keep"])
        self.calls.append(("fastv", i, hidden.shape[1], out[0].shape[1]))
        return out

    def keep_indexs(self):
        return self._keep


def swap_operator(model, op):
    for m in [model, model.model] + list(model.model.layers) + [l.self_attn for l in model.model.layers]:
        assert isinstance(m.baseline, TokenReductionBaseline)
        m.baseline = op


@pytest.fixture()
def patched_importance(monkeypatch):
    import framefusion_b200.hooks.qwen2 as hk

    def cpu_importance(q, k, v, num=1, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False):
        return port.last_query_attention(q, k, num=num, is_causal=is_causal, scale=scale)
    monkeypatch.setattr(hk, "scaled_dot_product_attention", cpu_importance)


def manual_reference(model, wl, op):
    """The schedule of the baseline hooks spelled out with stock modules: merge / FastV on the INPUT of a layer, the
    attention of layer fastv_k - 1 hands over its last query's probabilities (reference :308-342, :916-920)."""
    from transformers.models.qwen2.modeling_qwen2 import apply_rotary_pos_emb
    m = model.model
    h = wl.hidden.clone()
    pe = list(m.rotary_emb(h, torch.arange(h.shape[1])[None]))
    op.prepare(*wl.prepare_args())
    w = None
    for li, layer in enumerate(m.layers):
        if op.fastv_k is not None and li == op.fastv_k:
            h, pe, _ = op.fastv_at(li, h, pe, None, w)
        h, pe, _ = op.merge_at(li, h, pe, None)
        x = layer.input_layernorm(h)
        att = layer.self_attn
        shp = (*x.shape[:-1], -1, att.head_dim)
        q = att.q_proj(x).view(shp).transpose(1, 2)
        k = att.k_proj(x).view(shp).transpose(1, 2)
        v = att.v_proj(x).view(shp).transpose(1, 2)
        q, k = apply_rotary_pos_emb(q, k, pe[0], pe[1])
        w = port.last_query_attention(q, k, num=1, is_causal=True, scale=att.scaling) if op.wants_attention(li) else None
        kk = k.repeat_interleave(att.num_key_value_groups, dim=1)
        vv = v.repeat_interleave(att.num_key_value_groups, dim=1)
        o = torch.nn.functional.scaled_dot_product_attention(q, kk, vv, is_causal=True, scale=att.scaling)
        h = h + att.o_proj(o.transpose(1, 2).reshape(*x.shape[:-1], -1))
        h = h + layer.mlp(layer.post_attention_layernorm(h))
    return m.norm(h)


CONFIGS = {
    "fastv": dict(mode="fastv", fastv_k=2, fastv_r=0.5),
    "prefill_merge": dict(mode="prefill_merge", sparsity=[0.2, 0.0, 0.1, 0.3]),
    "merge_then_fastv": dict(mode="merge_then_fastv", sparsity=[0.1, 0.1, 0.1, 0.1], fastv_k=2, fastv_r=0.25),
    "fastv_then_merge": dict(mode="fastv_then_merge", fastv_k=1, fastv_r=0.75, merging_sparsity=0.3),
}


def operator_settings(name):
    c = dict(CONFIGS[name])
    c.pop("mode")
    if name == "fastv_then_merge":
        return dict(sparsity=[0.0] * (c["fastv_k"] + 1) + [c["merging_sparsity"]], fastv_k=c["fastv_k"], fastv_r=c["fastv_r"])
    return dict(sparsity=c.get("sparsity"), fastv_k=c.get("fastv_k"), fastv_r=c.get("fastv_r", 0.5))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_hook_schedule_matches_manual_reference(patched_importance, name, capsys):
    model = tiny_model()
    wl = workload(frames=6, patches=12)
    cfg = dict(CONFIGS[name])
    installed = qb.replace_Qwen2_forward(model, **cfg)
    assert isinstance(installed, TokenReductionBaseline) and model.baseline is installed
    s = operator_settings(name)
    assert (installed.sparsity, installed.fastv_k, installed.fastv_r) == (s["sparsity"], s["fastv_k"], s["fastv_r"])
    assert f"mode: {cfg['mode']}" in capsys.readouterr().out
    op = OracleBaselineOperator(**s)
    swap_operator(model, op)
    with torch.no_grad():
        want = manual_reference(model, wl, OracleBaselineOperator(**s))
        op.prepare(*wl.prepare_args())
        out = model.model(inputs_embeds=wl.hidden.clone(), use_cache=True)
    got = out.last_hidden_state
    assert got.shape == want.shape and got.shape[1] < wl.seq_len
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    reducing = [c for c in op.calls if c[2] != c[3]]
    assert reducing, "nothing was reduced"
    if s["fastv_k"] is not None:
        fv = [c for c in op.calls if c[0] == "fastv"]
        assert len(fv) == 1 and fv[0][1] == s["fastv_k"] and fv[0][3] < fv[0][2]
    # every layer's KV cache keeps the length that layer saw: non-increasing, layer 0 possibly already reduced
    lens = [out.past_key_values.get_seq_length(i) for i in range(len(model.model.layers))]
    assert all(a >= b for a, b in zip(lens, lens[1:])) and lens[-1] == got.shape[1]
    # a decode step passes through untouched
    with torch.no_grad():
        n_calls = len(op.calls)
        res = model.model(inputs_embeds=torch.randn(1, 1, 64), past_key_values=out.past_key_values, use_cache=True)
    assert res.last_hidden_state.shape == (1, 1, 64) and torch.isfinite(res.last_hidden_state).all()
    assert all(c[2] == c[3] for c in op.calls[n_calls:])
    assert [res.past_key_values.get_seq_length(i) for i in range(len(lens))] == [l + 1 for l in lens]


def test_fastv_counts_and_kept_text_tokens(patched_importance):
    """round(L * (1 - r)) vision tokens survive, every text token does (reference :325-329)."""
    model = tiny_model()
    wl = workload(frames=5, patches=11)
    qb.replace_Qwen2_fastv(model, fastv_k=3, fastv_r=0.4)
    assert (model.fastv_k, model.fastv_r) == (3, 0.4)
    op = OracleBaselineOperator(None, 3, 0.4)
    swap_operator(model, op)
    with torch.no_grad():
        op.prepare(*wl.prepare_args())
        out = model.model(inputs_embeds=wl.hidden.clone(), use_cache=True)
    L = 5 * 11
    assert out.last_hidden_state.shape[1] == wl.seq_len - L + round(L * (1 - 0.4))
    keep = op.keep_indexs().numpy()
    pt = wl.patch_type.numpy().reshape(-1)
    assert set(np.nonzero(pt == -1)[0].tolist()) <= set(keep.tolist()) and np.all(np.diff(keep) > 0)


def test_fastv_needs_the_cache():
    model = tiny_model()
    qb.replace_Qwen2_fastv(model)
    with pytest.raises(NotImplementedError, match="use_cache"):
        model.model(inputs_embeds=torch.randn(1, 8, 64), use_cache=False)


def test_meta_interface_errors():
    model = tiny_model()
    with pytest.raises(NotImplementedError, match="minference"):
        qb.replace_Qwen2_forward(model, mode="streamingllm")
    with pytest.raises(NotImplementedError, match="not implemented"):
        qb.replace_Qwen2_forward(model, mode="merge_then_fastv_cost_given")       # the reference's own default raises, :108-109

    class NotQwen(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.Linear(2, 2)
    with pytest.raises(TypeError, match="not Qwen2"):
        qb.replace_Qwen2_fastv(NotQwen())


def test_compute_density_overhead():
    """reference :26-39"""
    cost, remaining = compute_density_overhead([0.1] * 28)
    want = sum(0.9 ** (i + 1) for i in range(28)) / 28
    assert math.isclose(cost, want, rel_tol=1e-12) and math.isclose(remaining, 0.9 ** 28, rel_tol=1e-12)
    assert compute_density_overhead([0.0] * 4) == (1.0, 1.0)
    assert compute_density_overhead([0.5, 0.0]) == (0.5, 0.5)


def test_operator_has_no_cpu_fallback():
    op = TokenReductionBaseline([0.5], 1, 0.5)
    wl = workload()
    op.prepare(*wl.prepare_args())
    pe = [torch.zeros(1, wl.seq_len, 16), torch.zeros(1, wl.seq_len, 16)]
    with pytest.raises(RuntimeError, match="CUDA"):
        op.merge_at(0, wl.hidden, pe, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        op.fastv_at(1, wl.hidden, pe, None, torch.rand(1, 4, 1, wl.seq_len))
    with pytest.raises(RuntimeError, match="merge_at"):
        op(wl.hidden, pe, None)


def test_fixed_sparsity_oracle_counts():
    """floor(s * n_vis) tokens go at every layer, text tokens never (reference :916-920)."""
    wl = workload(frames=8, patches=10, lo=0.0, hi=1.0)
    o = OracleBaseline([0.25, 0.0, 0.5], None, 0.5, "f32")
    o.prepare(wl.patch_type.numpy(), *wl.prepare_args()[1:])
    h, pe = t2f(wl.hidden[0]), [np.zeros((wl.seq_len, 4), np.float32)] * 2
    n_vis = 80
    for li, s in enumerate([0.25, 0.0, 0.5]):
        before = h.shape[0]
        h, pe, _ = o.merge_at(li, h, pe, None)
        k = math.floor(s * n_vis)
        assert before - h.shape[0] == k
        n_vis -= k
        assert int((o.patch_type != -1).sum()) == n_vis and int((o.patch_type == -1).sum()) == wl.seq_len - 80


def test_wrapper_models_hold_the_decoder_under_llm(patched_importance, capsys):
    """MiniCPM-V / NVILA keep the Qwen2 stack under ``model.llm.model`` (reference :190-218, meta interfaces :111-164)."""
    class Wrapper(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.llm = tiny_model()
    for installer, kw in ((qb.replace_minicpmv_forward, dict(mode="fastv", fastv_k=2, fastv_r=0.5)),
                          (qb.replace_nvila_forward, dict(mode="fastv", fastv_k=2, fastv_r=0.5))):
        model = Wrapper()
        op_cuda = installer(model, **kw)
        assert isinstance(op_cuda, TokenReductionBaseline) and model.baseline is op_cuda and (model.fastv_k, model.fastv_r) == (2, 0.5)
        assert installer.__name__ in capsys.readouterr().out
        inner = model.llm.model
        assert inner.baseline is op_cuda and all(l.baseline is op_cuda and l.self_attn.baseline is op_cuda for l in inner.layers)
        op = OracleBaselineOperator(None, 2, 0.5)
        for m in [model, inner] + list(inner.layers) + [l.self_attn for l in inner.layers]:
            m.baseline = op
        wl = workload(frames=6, patches=12)
        with torch.no_grad():
            want = manual_reference(model.llm, wl, OracleBaselineOperator(None, 2, 0.5))
            op.prepare(*wl.prepare_args())
            out = inner(inputs_embeds=wl.hidden.clone(), use_cache=True)
        assert torch.allclose(out.last_hidden_state, want, atol=1e-5, rtol=1e-5) and out.last_hidden_state.shape[1] < wl.seq_len
    with pytest.raises(NotImplementedError, match="minference"):
        qb.replace_minicpmv_forward(Wrapper(), mode="streamingllm")
    with pytest.raises(NotImplementedError, match="not implemented"):
        qb.replace_nvila_forward(Wrapper())                                  # the reference's default mode raises too (:163-164)

    class NotQwen(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.llm = torch.nn.Module()
            self.llm.model = torch.nn.Linear(2, 2)
    with pytest.raises(TypeError, match="not Qwen2"):
        qb.replace_minicpmv_fastv(NotQwen())
