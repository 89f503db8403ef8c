"""GPU: the CUDA path (through the C ABI) against the golden fixtures of the unmodified reference and against
the numpy oracle.  Bit-exact for indices, masks, patch types, position tensors and merged rows; similarities
bit-exact except where the oracle proves a float32 summation order can move the value (then bracketed)."""
import json
import os

import numpy as np
import pytest
import torch

from _harness import MODES, set_mode, GOLDEN_DIR, DT, case_names, run_and_compare, t2f
from oracle import ff_oracle as orc

pytestmark = pytest.mark.gpu


class CudaAdapter:
    """``framefusion_b200.main.FrameFusion`` behind the interface the harness drives."""

    def __init__(self, cost, slb, rlb, dtype, mode):
        from framefusion_b200.main import FrameFusion
        self.ff = FrameFusion(cost, slb, rlb)
        self.ff.debug_trace = True
        set_mode(self.ff, mode)

    def prepare(self, *args):
        self.ff.prepare(*args)

    def __call__(self, hidden, pos, mask, attn=None):
        return self.ff(hidden, pos, mask, attn)

    finish_merging = property(lambda s: s.ff.finish_merging)
    finish_pruning = property(lambda s: s.ff.finish_pruning)
    sparsity_list = property(lambda s: s.ff.sparsity_list)
    patch_type = property(lambda s: s.ff.patch_type)

    def last_trace(self):
        return self.ff.last_trace


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", case_names())
def test_cuda_matches_reference_sequence(name, mode):
    rep = run_and_compare(name, lambda c, s, r, dt: CudaAdapter(c, s, r, dt, mode), device="cuda")
    assert rep["n_sim"] > 0


@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
def test_static_similarity_and_merge_vs_oracle(dtype):
    from framefusion_b200 import synth
    from framefusion_b200.main import FrameFusion
    wl = synth.make_workload(frames=9, patch_num=37, hidden=1024, dtype=DT[dtype], seed=21)
    hidden = wl.hidden.cuda()
    sim, order = FrameFusion.compute_similarity_and_token_index_by_patch(hidden, wl.patch_type.cuda(), wl.patch_num)
    sr = orc.similarity_by_patch(t2f(wl.hidden[0]), wl.patch_type.numpy().reshape(-1), wl.patch_num, dtype)
    assert sim.dtype == hidden.dtype and order.dtype == torch.int64 and sim.shape == order.shape == (1, sr.order.shape[0])
    assert np.array_equal(order[0].cpu().numpy(), sr.order)
    got = t2f(sim[0])
    neq = got != sr.sim
    assert not (neq & ~sr.fragile).any()
    assert ((got >= sr.lo) & (got <= sr.hi))[neq].all()
    # merge a hand-made index set (runs of length 1..4, never a chain head) in place
    thr = orc.threshold_in_dtype(0.5, dtype)
    idx = np.nonzero(got >= thr)[0]
    want_h, want_keep = orc.merge_tokens_and_get_mask(t2f(wl.hidden[0]), sr.order, idx, dtype)
    h2 = hidden.clone()
    out, keep = FrameFusion.merge_tokens_and_get_mask(h2, sim, order, torch.from_numpy(idx).cuda())
    assert out.data_ptr() == h2.data_ptr()                    # in place, like the reference
    assert keep.dtype == torch.bool and np.array_equal(keep[0].cpu().numpy(), want_keep)
    if dtype == "f32":
        assert np.allclose(t2f(out[0]), want_h, rtol=1e-6, atol=1e-7)
    else:
        assert np.array_equal(t2f(out[0]), want_h)
    # empty index: untouched, all-True mask (main.py:264-266)
    out, keep = FrameFusion.merge_tokens_and_get_mask(h2, sim, order, torch.zeros(0, dtype=torch.int64).cuda())
    assert bool(keep.all())


def test_importance_vs_reference():
    from framefusion_b200 import synth
    from framefusion_b200.utils import scaled_dot_product_attention
    z = np.load(os.path.join(GOLDEN_DIR, "importance.npz"))
    for tag in "abcdef":
        spec = json.loads(str(z[f"imp_{tag}_spec"]))
        dt = spec["dtype"]
        q, k = synth.make_attention_inputs(spec["s_len"], 28, 4, 128, DT[dt], seed=spec["s_len"])
        # K is handed over before repeat_kv: the kernel is GQA aware
        got = scaled_dot_product_attention(q.cuda(), k.cuda(), None, num=spec["num"], is_causal=spec["causal"], enable_gqa=True)
        got = t2f(got[0])
        want = orc.bits_to_f32(z[f"imp_{tag}_out"], dt).reshape(got.shape)
        if dt == "f32":
            assert np.allclose(got, want, rtol=3e-5, atol=1e-9)
        else:
            ulp = np.abs(want) * (2.0 ** -7 if dt == "bf16" else 2.0 ** -10) + 1e-30
            assert (np.abs(got - want) <= ulp * 1.01).all()
            assert (got != want).mean() < 0.02, (tag, (got != want).mean())


def test_importance_strided_inputs():
    """q / k as the attention hook has them: [1, S, heads, D] storage viewed as [1, heads, S, D]."""
    from framefusion_b200 import synth
    from framefusion_b200.utils import scaled_dot_product_attention
    q, k = synth.make_attention_inputs(300, 28, 4, 128, torch.bfloat16, seed=5)
    qs = q.transpose(1, 2).contiguous().transpose(1, 2).cuda()
    ks = k.transpose(1, 2).contiguous().transpose(1, 2).cuda()
    assert not qs.is_contiguous()
    a = scaled_dot_product_attention(qs, ks, None, num=1, is_causal=True, enable_gqa=True)
    b = scaled_dot_product_attention(q.cuda(), k.cuda(), None, num=1, is_causal=True, enable_gqa=True)
    assert torch.equal(a, b)
    kk = k.repeat_interleave(7, dim=1).cuda()
    c = scaled_dot_product_attention(q.cuda(), kk, None, num=1, is_causal=True)
    assert torch.equal(a, c)


def test_errors_and_passthrough():
    from framefusion_b200 import synth
    from framefusion_b200.main import FrameFusion
    wl = synth.make_workload(frames=4, patch_num=8, hidden=256, dtype=torch.bfloat16, seed=1)
    ff = FrameFusion(0.01, 0.0, 0.1)
    ff.prepare(*wl.prepare_args())
    h = wl.hidden.cuda()
    pos = [wl.cos.cuda(), wl.sin.cuda()]
    # decode step: untouched (main.py:61,104 guards)
    h1, p1, m1 = ff(h[:, :1], pos, None)
    assert h1.shape[1] == 1 and p1 is pos and m1 is None
    ff.sparsity_list = [0.0] * 10
    with pytest.raises(ValueError, match="The cost is too small"):
        ff(h, pos, None)
    ff = FrameFusion()
    ff.prepare(*wl.prepare_args())
    with pytest.raises(NotImplementedError):
        ff(h, (wl.cos.cuda(), wl.sin.cuda()), None)           # tuple, not list (main.py:176-177)
    with pytest.raises(NotImplementedError):
        ff(h, wl.cos.cuda(), None)                            # 3-D tensor (main.py:174-175)
    with pytest.raises(AssertionError, match="Only support batch size 1"):
        ff(torch.cat([h, h]), pos, None)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ff(wl.hidden, [wl.cos, wl.sin], None)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_full_size_against_oracle(cfg, mode):
    """BASELINE configs at full size: first merge call, CUDA vs the numpy oracle on identical bits."""
    from framefusion_b200 import synth
    from framefusion_b200.main import FrameFusion
    c = synth.CONFIGS[cfg]
    wl = synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0)
    ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
    set_mode(ff, mode)
    ff.debug_trace = True
    ff.prepare(*wl.prepare_args())
    pos = [wl.cos.cuda(), wl.sin.cuda()]
    out, pos, _ = ff(wl.hidden.cuda(), pos, None)
    tr = ff.last_trace
    o = orc.OracleFrameFusion(c["cost"], c["slb"], c["rlb"], "bf16")
    o.prepare(wl.patch_type.numpy(), wl.patch_num, *wl.prepare_args()[2:])
    hid = t2f(wl.hidden[0])
    want_h, want_pos, _ = o.forward(hid, [t2f(wl.cos[0]), t2f(wl.sin[0])], None)
    sr = o.last["sim"]
    got_sim = tr["sim_values"]
    neq = got_sim != sr.sim
    assert not (neq & ~sr.fragile).any()
    assert ((got_sim >= sr.lo) & (got_sim <= sr.hi))[neq].all()
    thr = orc.threshold_in_dtype(c["slb"], "bf16")
    flips = int(((got_sim >= thr) != (sr.sim >= thr)).sum())
    print(f"{cfg}: N={got_sim.shape[0]} sims differing from oracle (all inside the fragile bracket): {int(neq.sum())}, threshold flips: {flips}")
    assert np.array_equal(tr["order"], sr.order)
    if flips == 0:
        assert np.array_equal(tr["keep_mask"], o.last["keep_mask"])
        assert out.shape[1] == want_h.shape[0]
        assert np.array_equal(t2f(out[0]), want_h)
        assert np.array_equal(t2f(pos[0][0]), want_pos[0]) and np.array_equal(t2f(pos[1][0]), want_pos[1])
        assert np.array_equal(ff.patch_type[0].cpu().numpy(), o.patch_type)
        assert ff.sparsity_list == o.sparsity_list
    else:
        # a similarity sitting on a rounding boundary of T moved across the threshold: selection differs in
        # exactly those positions
        diff = np.nonzero(tr["keep_mask"] != o.last["keep_mask"])[0]
        assert len(diff) == flips


def test_kernel_events_hook_times_the_library_side_launches():
    """``kernel_events`` hands two events to ff_ctx_timing: the library records them around its own launches."""
    from framefusion_b200 import synth
    from framefusion_b200.main import FrameFusion
    wl = synth.to_device(synth.make_workload(24, 128, 512, torch.bfloat16, seed=2), "cuda")
    ff = FrameFusion(0.3, 0.6, 0.1)
    ff.reserve_kernel_events(2)
    ff.prepare(*wl.prepare_args())
    ff.kernel_events = []
    h, _pos, _m = ff(wl.hidden, [wl.cos, wl.sin], None)
    torch.cuda.synchronize()
    assert len(ff.kernel_events) == 1
    name, q_len, e0, e1 = ff.kernel_events[0]
    assert name == "ff_merge_layer" and q_len == wl.seq_len
    ms = e0.elapsed_time(e1)
    assert 0.001 < ms < 50.0
    ff.kernel_events = None
    ff.prepare(*wl.prepare_args())
    h2, _pos, _m = ff(wl.hidden, [wl.cos, wl.sin], None)               # the hook off again: same result, no events touched
    assert torch.equal(h, h2)


# ---- importance -> prune, end to end at full size ------------------------------------------------------------------------
def test_importance_to_prune_selection_at_full_size():
    """q / K -> ff_importance -> head mean -> top-k at C2's post-merge length (S = 22 290, k = 30 % of the vision rows):
    the retained index set against the reference's arithmetic (utils.py:27-57 -> main.py:69-92, restated in the oracle and
    pinned by tests/golden/importance.npz).  The selection kernel must reproduce the oracle's choice exactly on the GPU's
    own importance; against the reference's importance every differing index must be explained by one of two causes:
    a tie at the k-th value (torch.topk leaves the choice among equal values open) or a probability that differs by one
    ulp of T and sits next to the k-th value."""
    from framefusion_b200 import synth
    from framefusion_b200.main import FrameFusion
    from framefusion_b200.utils import scaled_dot_product_attention
    S, n_pre, n_post, H = 22290, 14, 20, 256
    length = S - n_pre - n_post
    q, k = synth.make_attention_inputs(S, 28, 4, 128, torch.bfloat16, seed=3)
    q_last = q[:, :, -1:, :].contiguous()
    gpu_attn = scaled_dot_product_attention(q_last.cuda(), k.cuda(), None, num=1, is_causal=True, enable_gqa=True)
    assert gpu_attn.shape == (1, 28, 1, S)
    ref_attn = orc.last_query_attention(t2f(q[0]), t2f(k[0]), 1, "bf16", is_causal=True)            # [28, 1, S]
    got_attn = t2f(gpu_attn[0])
    ulp = np.abs(got_attn.view(np.int32) - ref_attn.view(np.int32)) >> 16                           # bf16 steps (same sign: probabilities)
    # one step of T nearly everywhere; a handful of probabilities move further: their LOGIT (|q.k| up to ~40, one bf16 step
    # there is 0.25) sits on a rounding boundary of T and the float32 accumulation order of the dot product decides the side
    # — the same fragility the similarity has, one stage earlier (a 0.25 step of a logit is 2 % of the probability)
    n_ulp, n_far = int((ulp != 0).sum()), int((ulp > 1).sum())
    assert n_far <= 1e-3 * ulp.size and n_ulp <= 0.03 * ulp.size
    # the operator on the GPU's importance
    g = torch.Generator().manual_seed(5)
    hidden = torch.randn(1, S, H, generator=g).to(torch.bfloat16).cuda()
    pos = [torch.randn(1, S, 128, generator=g).to(torch.bfloat16).cuda() for _ in range(2)]
    pt = torch.tensor([[-1] * n_pre + [i % 576 for i in range(length)] + [-1] * n_post])
    ff = FrameFusion(0.3, 0.6, 0.1)
    ff.debug_trace = True
    ff.prepare(pt.cuda(), 576, n_pre, n_pre + length - 1, length, S, finish_merging=True)
    out, _pos, _ = ff(hidden, pos, None, gpu_attn)
    keep_gpu = ff.last_trace["keep"]
    ratio = orc.compute_pruning_ratio([], 0.3)
    kk = round(length * (1 - ratio))
    imp_gpu = orc.mean_heads(got_attn, "bf16")
    keep_same_input = orc.prune_keep_indices(imp_gpu, n_pre, length, S, ratio)
    assert np.array_equal(keep_gpu, keep_same_input), "selection differs from the oracle on identical importance"
    assert out.shape[1] == n_pre + kk + n_post
    # against the reference's importance
    imp_ref = orc.mean_heads(ref_attn, "bf16")
    keep_ref = orc.prune_keep_indices(imp_ref, n_pre, length, S, ratio)
    diff = np.setxor1d(keep_gpu, keep_ref)
    vis = imp_ref[n_pre:n_pre + length]
    kth = np.sort(vis)[::-1][kk - 1]
    tie = imp_ref[diff] == kth
    moved = imp_gpu[diff] != imp_ref[diff]
    # a row whose own value did not move and is not tied can only change sides because a neighbour in the ranking did
    pushed = ~tie & ~moved
    n_tied_total = int((vis == kth).sum())
    print(f"S={S} k={kk}: probabilities differing {n_ulp} of {ulp.size} ({n_far} by more than one bf16 step, max {int(ulp.max())}); mean-importance values differing "
          f"{int((imp_gpu != imp_ref).sum())} of {S}; retained indices differing {diff.size} "
          f"(tie at the k-th value: {int(tie.sum())} of {n_tied_total} tied rows, own value moved by one ulp: {int((~tie & moved).sum())}, "
          f"displaced by those: {int(pushed.sum())})")
    steps = np.abs(imp_gpu.view(np.int32) - imp_ref.view(np.int32)) >> 16
    assert steps.max() <= 2 and int((steps > 1).sum()) <= 1e-3 * S
    dist = np.abs(imp_ref[diff].view(np.int32) - kth.view(np.int32)) >> 16
    assert np.all(dist <= 2), "a differing index is not next to the k-th value"
    assert diff.size <= 2 * (n_tied_total + int((imp_gpu != imp_ref).sum()))


# ---- the whole bench step at full size --------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["C2", "C3", "C4"])
def test_full_step_against_oracle(cfg):
    """Every call of the bench step — merge, merge (closes merging), importance, prune — at the BASELINE sizes, CUDA against
    the numpy oracle call by call on identical bits (the oracle gets the GPU's importance: the step above pins that)."""
    from framefusion_b200 import synth
    from framefusion_b200.main import FrameFusion
    from framefusion_b200.utils import scaled_dot_product_attention
    c = synth.CONFIGS[cfg]
    wl = synth.make_workload(c["frames"], c["patch_num"], c["hidden"], c["dtype"], seed=0)
    q, k = synth.make_attention_inputs(wl.seq_len, 28, 4, 128, c["dtype"], seed=0)
    q_last, k = q[:, :, -1:, :].contiguous().cuda(), k.cuda()
    ff = FrameFusion(c["cost"], c["slb"], c["rlb"])
    ff.debug_trace = True
    ff.prepare(*synth.to_device(wl, "cuda").prepare_args())
    o = orc.OracleFrameFusion(c["cost"], c["slb"], c["rlb"], "bf16")
    o.prepare(wl.patch_type.numpy(), wl.patch_num, *wl.prepare_args()[2:])
    h, pos = wl.hidden.cuda(), [wl.cos.cuda(), wl.sin.cuda()]
    stages = []
    for call in range(8):
        if ff.finish_merging and ff.finish_pruning:
            break
        attn = None
        if ff.finish_merging:
            attn = scaled_dot_product_attention(q_last, k[:, :, :h.shape[1]], None, num=1, is_causal=True, enable_gqa=True)
        h_in, p_in = t2f(h[0]), [t2f(pos[0][0]), t2f(pos[1][0])]
        h, pos, _ = ff(h, pos, None, attn)
        want_h, want_p, _ = o.forward(h_in, p_in, None, None if attn is None else t2f(attn[0]))
        stage = o.last["stage"]
        stages.append(stage)
        if stage == "merge":
            sr, got_sim = o.last["sim"], ff.last_trace["sim_values"]
            neq = got_sim != sr.sim
            assert not (neq & ~sr.fragile).any() and ((got_sim >= sr.lo) & (got_sim <= sr.hi))[neq].all()
            thr = orc.threshold_in_dtype(c["slb"], "bf16")
            assert int(((got_sim >= thr) != (sr.sim >= thr)).sum()) == 0, "a fragile similarity crossed the threshold: pick another seed"
        assert h.shape[1] == want_h.shape[0], f"{cfg} call {call} ({stage}): kept {h.shape[1]}, oracle {want_h.shape[0]}"
        assert np.array_equal(t2f(h[0]), want_h), f"{cfg} call {call} ({stage}): hidden_states differ"
        assert np.array_equal(t2f(pos[0][0]), want_p[0]) and np.array_equal(t2f(pos[1][0]), want_p[1])
        assert (ff.finish_merging, ff.finish_pruning) == (o.finish_merging, o.finish_pruning)
        assert ff.sparsity_list == o.sparsity_list
    assert stages[0] == "merge" and stages[-1] == "prune", stages
    print(f"{cfg}: S {wl.seq_len} -> {h.shape[1]} through {stages}")
