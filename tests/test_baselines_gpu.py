"""GPU: the fixed-amount comparison methods through the CUDA operator (``framefusion_b200.baselines``), checked against the
numpy oracle on the same inputs — FastV (reference modeling_qwen2_baseline.py:318-342), fixed-sparsity merging (:916-920,
:1003) and the combination inside a patched tiny decoder, call by call."""
import math

import numpy as np
import pytest
import torch

from _harness import t2f
from oracle import ff_oracle as orc
from oracle.ff_baselines_oracle import OracleBaseline
from framefusion_b200 import synth
from framefusion_b200.baselines import TokenReductionBaseline

pytestmark = pytest.mark.gpu


def check_merge_call(op, h_in, pt_in, patch_num, s, out, dtype="bf16"):
    """One fixed-sparsity call of the CUDA operator against the oracle: the branch, the amount, the similarities (bit-equal
    outside the oracle's fragile set, bracketed inside), the selection (exactly the rule: top-k of the similarities the device
    computed, lowest index first), the merged rows and the compaction bit for bit."""
    tr = op.last_trace
    n_vis = int((pt_in != -1).sum())
    k = math.floor(s * n_vis)
    assert tr["stage"] == "merge" and tr["branch"] == "topk"
    sim = np.asarray(tr["sim_values"], np.float32)
    mi = np.asarray(tr["merge_index"])
    assert mi.shape[0] == k
    sr = orc.similarity_by_patch(h_in, pt_in, patch_num, dtype)
    assert np.array_equal(np.asarray(tr["order"]), sr.order)
    neq = ~((sim == sr.sim) | (np.isnan(sim) & np.isnan(sr.sim)))
    assert not (neq & ~sr.fragile).any() and ((sim >= sr.lo) & (sim <= sr.hi))[neq].all()
    assert np.array_equal(mi, orc.topk_lowest_index(sim, k))
    merged, keep = orc.merge_tokens_and_get_mask(h_in, sr.order, mi, dtype)
    assert np.array_equal(np.asarray(tr["keep_mask"]), keep)
    assert np.array_equal(t2f(out[0][0]), merged[keep])
    return keep


@pytest.mark.parametrize("hidden", [1024, 3584])
def test_fixed_sparsity_merge_against_oracle(hidden):
    wl = synth.make_workload(12, 40, hidden, torch.bfloat16, seed=3, r_lo=0.0, r_hi=1.0, n_pre=5, n_post=9, rot_dim=64)
    dev = synth.to_device(wl, "cuda")
    sparsity = [0.3, 0.0, 0.2, 0.05]
    op = TokenReductionBaseline(sparsity)
    op.debug_trace = True
    op.prepare(*dev.prepare_args())
    h, pos, mask = dev.hidden, [dev.cos, dev.sin], None
    pt = wl.patch_type.numpy().reshape(-1)
    n_vis = 12 * 40
    for li, s in enumerate(sparsity):
        h_in, p_in = t2f(h[0]), [t2f(pos[0][0]), t2f(pos[1][0])]
        out = op.merge_at(li, h, pos, mask)
        k = math.floor(s * n_vis)
        assert out[0].shape[1] == h_in.shape[0] - k
        if k:
            keep = check_merge_call(op, h_in, pt, wl.patch_num, s, out)
            assert np.array_equal(t2f(out[1][0][0]), p_in[0][keep]) and np.array_equal(t2f(out[1][1][0]), p_in[1][keep])
            pt = pt[keep]
            assert np.array_equal(op.patch_type.reshape(-1).cpu().numpy(), pt)
        n_vis -= k
        assert op.frame_token_num == n_vis
        h, pos, mask = out
    # the operator's budget state is not involved
    assert op.sparsity_list == [] and not op.finish_merging and not op.finish_pruning


@pytest.mark.parametrize("r", [0.5, 0.9, 0.0])
def test_fastv_against_oracle(r):
    wl = synth.make_workload(16, 96, 1024, torch.bfloat16, seed=7, n_pre=14, n_post=20, rot_dim=128)
    dev = synth.to_device(wl, "cuda")
    S = wl.seq_len
    attn = synth.make_attention_row(S, n_heads=28, num=1, dtype=torch.bfloat16, seed=99)
    mask = torch.triu(torch.full((S, S), float("-inf")), 1).to(torch.bfloat16)[None, None]
    op = TokenReductionBaseline(None, fastv_k=3, fastv_r=r)
    op.prepare(*dev.prepare_args())
    o = OracleBaseline(None, 3, r, "bf16")
    o.prepare(None, None, *wl.prepare_args()[2:])
    h_in, p_in = t2f(wl.hidden[0]), [t2f(wl.cos[0]), t2f(wl.sin[0])]
    # not the layer FastV acts on: untouched
    same = op.fastv_at(2, dev.hidden, [dev.cos, dev.sin], None, attn.cuda())
    assert same[0] is dev.hidden
    got = op.fastv_at(3, dev.hidden, [dev.cos, dev.sin], mask.cuda(), attn.cuda())
    want_h, want_p, want_m = o.fastv_at(3, h_in, p_in, t2f(mask[0, 0]), t2f(attn[0]))
    L = 16 * 96
    assert got[0].shape[1] == S - L + round(L * (1 - r))
    keep = op.keep_indexs().cpu().numpy()
    assert np.array_equal(keep, o.last["keep"])             # (ties at the k-th value: lowest index on both sides)
    assert np.array_equal(t2f(got[0][0]), want_h)
    assert np.array_equal(t2f(got[1][0][0]), want_p[0]) and np.array_equal(t2f(got[1][1][0]), want_p[1])
    assert np.array_equal(t2f(got[2][0, 0]), want_m)


def tiny_model():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(0)
    cfg = Qwen2Config(vocab_size=128, hidden_size=256, intermediate_size=512, num_hidden_layers=6, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=8192, rope_theta=1e6)
    cfg._attn_implementation = "sdpa"
    return Qwen2ForCausalLM(cfg).eval().to(torch.bfloat16).cuda()


@pytest.mark.parametrize("mode,kw", [
    ("fastv", dict(fastv_k=2, fastv_r=0.5)),
    ("prefill_merge", dict(sparsity=[0.2, 0.1, 0.0, 0.1, 0.0, 0.3])),
    ("merge_then_fastv", dict(sparsity=[0.1] * 6, fastv_k=3, fastv_r=0.5)),
    ("fastv_then_merge", dict(fastv_k=2, fastv_r=0.75, merging_sparsity=0.3)),
])
def test_patched_prefill_call_by_call(mode, kw):
    """The patched decoder on the GPU: every reducing call of the operator is replayed through the oracle on the very same
    inputs; FastV's importance comes from ``ff_importance`` inside the attention of layer fastv_k - 1."""
    from framefusion_b200.hooks.qwen2_baselines import replace_Qwen2_forward
    model = tiny_model()
    op = replace_Qwen2_forward(model, mode=mode, **kw)
    op.debug_trace = True
    wl = synth.make_workload(10, 24, 256, torch.bfloat16, seed=11, n_pre=5, n_post=7, rot_dim=64)
    o = OracleBaseline(op.sparsity, op.fastv_k, op.fastv_r, "bf16")
    o.prepare(wl.patch_type.numpy(), *wl.prepare_args()[1:])
    calls = []
    inner_merge, inner_fastv = op.merge_at, op.fastv_at

    def merge_at(i, hidden, pos, mask):
        h_in = t2f(hidden[0])
        pt_in = op.patch_type.reshape(-1).cpu().numpy().copy()
        if op.sparsity is not None and any(s > 0 for s in op.sparsity[i:]):
            assert np.array_equal(pt_in, o.patch_type)      # (the layout is only kept up while merges are still to come)
        out = inner_merge(i, hidden, pos, mask)
        if out[0].shape[1] != h_in.shape[0]:
            keep = check_merge_call(op, h_in, pt_in, wl.patch_num, op.sparsity[i], out)
            o.patch_type = o.patch_type[keep]
            calls.append(("merge", i, h_in.shape[0], out[0].shape[1]))
        return out

    def fastv_at(i, hidden, pos, mask, attn):
        h_in, p_in = t2f(hidden[0]), [t2f(pos[0][0]), t2f(pos[1][0])]
        assert attn is not None and attn.shape == (1, 4, 1, hidden.shape[1])
        a_in = t2f(attn[0])
        out = inner_fastv(i, hidden, pos, mask, attn)
        want_h, want_p, _ = o.fastv_at(i, h_in, p_in, None, a_in)
        assert np.array_equal(op.keep_indexs().cpu().numpy(), o.last["keep"])
        assert np.array_equal(t2f(out[0][0]), want_h) and np.array_equal(t2f(out[1][0][0]), want_p[0])
        calls.append(("fastv", i, h_in.shape[0], out[0].shape[1]))
        return out

    op.merge_at, op.fastv_at = merge_at, fastv_at
    with torch.no_grad():
        op.prepare(*synth.to_device(wl, "cuda").prepare_args())
        res = model.model(inputs_embeds=wl.hidden.cuda(), use_cache=True)
        assert torch.isfinite(res.last_hidden_state.float()).all()
        assert res.last_hidden_state.shape[1] == calls[-1][3] < wl.seq_len
        if op.fastv_k is not None:
            assert [c[1] for c in calls if c[0] == "fastv"] == [op.fastv_k]
        if op.sparsity is not None:
            assert [c[1] for c in calls if c[0] == "merge"] == [i for i, s in enumerate(op.sparsity) if s > 0]
        # decode after the reduced prefill: per-layer ragged caches, one more key each
        lens = [res.past_key_values.get_seq_length(i) for i in range(6)]
        assert all(a >= b for a, b in zip(lens, lens[1:]))
        n = len(calls)
        step = model.model(inputs_embeds=torch.randn(1, 1, 256, device="cuda", dtype=torch.bfloat16),
                           past_key_values=res.past_key_values, use_cache=True)
        assert step.last_hidden_state.shape == (1, 1, 256) and torch.isfinite(step.last_hidden_state.float()).all()
        assert len(calls) == n and [step.past_key_values.get_seq_length(i) for i in range(6)] == [l + 1 for l in lens]
