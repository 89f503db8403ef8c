"""GPU: a tiny random-init Qwen2 decoder patched by ``apply_framefusion`` runs its prefill through the CUDA operator;
every operator call is checked IN SITU against the numpy oracle fed with the very same inputs (copied to the host),
so the two state machines stay in lock step through merge, merge-closing and prune calls."""
import numpy as np
import pytest
import torch

from _harness import MODES, set_mode, t2f
from oracle import ff_oracle as orc
from framefusion_b200 import synth

pytestmark = pytest.mark.gpu


def tiny_model():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(0)
    cfg = Qwen2Config(vocab_size=128, hidden_size=256, intermediate_size=512, num_hidden_layers=6, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=8192, rope_theta=1e6)
    cfg._attn_implementation = "sdpa"
    return Qwen2ForCausalLM(cfg).eval().to(torch.bfloat16).cuda()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("lo,hi", [(0.0, 1.0), (0.0, 0.5)], ids=["mixed", "lowsim_prune"])
def test_patched_prefill_matches_oracle_call_by_call(lo, hi, mode):
    from framefusion_b200.interface import apply_framefusion
    model = tiny_model()
    apply_framefusion(model, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
    ff = model.framefusion
    set_mode(ff, mode)
    wl = synth.make_workload(10, 24, 256, torch.bfloat16, seed=11, r_lo=lo, r_hi=hi, n_pre=5, n_post=7, rot_dim=64)
    o = orc.OracleFrameFusion(0.3, 0.6, 0.1, "bf16")
    o.prepare(wl.patch_type.numpy(), wl.patch_num, *wl.prepare_args()[2:])
    calls = []
    inner = ff.forward

    def checked(hidden, pos, mask, attn=None):
        h_in = t2f(hidden[0]); p_in = [t2f(pos[0][0]), t2f(pos[1][0])]
        a_in = None if attn is None else t2f(attn[0])
        out = inner(hidden, pos, mask, attn)
        want_h, want_p, _ = o.forward(h_in, p_in, None, a_in)
        got = t2f(out[0][0])
        fragile = o.last is not None and o.last.get("stage") == "merge" and bool(o.last["sim"].fragile.any())
        if got.shape == want_h.shape or not fragile:
            assert got.shape == want_h.shape, f"call {len(calls)}: {got.shape} vs oracle {want_h.shape}"
            assert np.array_equal(got, want_h), f"call {len(calls)}: hidden_states differ from the oracle"
            assert np.array_equal(t2f(out[1][0][0]), want_p[0]) and np.array_equal(t2f(out[1][1][0]), want_p[1])
        assert (ff.finish_merging, ff.finish_pruning) == (o.finish_merging, o.finish_pruning)
        calls.append((hidden.shape[1], out[0].shape[1], attn is not None))
        return out

    ff.forward = checked
    with torch.no_grad():
        ff.prepare(*synth.to_device(wl, "cuda").prepare_args())
        res = model.model(inputs_embeds=wl.hidden.cuda(), use_cache=True)
    assert len(calls) == 1 + 6
    assert res.last_hidden_state.shape[1] == calls[-1][1] < wl.seq_len
    assert torch.isfinite(res.last_hidden_state.float()).all()
    assert ff.sparsity_list == o.sparsity_list
    if hi <= 0.5:
        assert ff.finish_pruning and any(c[2] for c in calls)      # importance was produced and consumed


@pytest.mark.parametrize("mode", MODES)
def test_merge_call_is_repeatable_bit_for_bit(mode):
    """The kernels contain spin-waits and atomics: the result must not depend on their timing."""
    from framefusion_b200.main import FrameFusion
    wl = synth.to_device(synth.make_workload(16, 96, 3584, torch.bfloat16, seed=5, per_patch_r=True), "cuda")
    ref = None
    for rep in range(25):
        ff = FrameFusion(0.3, 0.6, 0.1)
        set_mode(ff, mode)
        ff.prepare(*wl.prepare_args())
        h, pos, _ = ff(wl.hidden, [wl.cos, wl.sin], None)
        cur = (h.clone(), pos[0].clone(), ff.patch_type.clone())
        if ref is None:
            ref = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(ref, cur)), f"repetition {rep} differs"


def test_decode_after_reduced_prefill_on_the_gpu():
    """Prefill through the CUDA operator, then decode: every layer's KV cache keeps the length that layer saw
    (modeling_qwen2.py:143-145: ragged, non-increasing), decode steps pass through the operator untouched and grow
    every layer by one; the result matches a dense decode of the same model only in shape, not in value."""
    from framefusion_b200.interface import apply_framefusion
    model = tiny_model()
    apply_framefusion(model, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
    ff = model.framefusion
    wl = synth.make_workload(10, 24, 256, torch.bfloat16, seed=11, r_lo=0.0, r_hi=1.0, n_pre=5, n_post=7, rot_dim=64)
    n_layers = len(model.model.layers)
    with torch.no_grad():
        ff.prepare(*synth.to_device(wl, "cuda").prepare_args())
        out = model.model(inputs_embeds=wl.hidden.cuda(), use_cache=True)
        lens = [out.past_key_values.get_seq_length(i) for i in range(n_layers)]
        assert lens[0] < wl.seq_len and all(a >= b for a, b in zip(lens, lens[1:]))      # layer 0 already merges
        assert out.last_hidden_state.shape[1] <= lens[-1]
        cache = out.past_key_values
        state = (ff.finish_merging, ff.finish_pruning, list(ff.sparsity_list))
        for step in range(2):
            tok = torch.randn(1, 1, 256, device="cuda", dtype=torch.bfloat16)
            mask = torch.ones(1, wl.seq_len + 1 + step, dtype=torch.long, device="cuda")      # generate() style mask
            res = model.model(inputs_embeds=tok, past_key_values=cache, use_cache=True, attention_mask=mask)
            assert res.last_hidden_state.shape == (1, 1, 256) and torch.isfinite(res.last_hidden_state.float()).all()
            cache = res.past_key_values
            assert [cache.get_seq_length(i) for i in range(n_layers)] == [l + 1 + step for l in lens]
        assert (ff.finish_merging, ff.finish_pruning, list(ff.sparsity_list)) == state      # q_len == 1: a no-op


def tiny_qwen2vl():
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    torch.manual_seed(0)
    cfg = Qwen2VLConfig(
        text_config=dict(vocab_size=128, hidden_size=256, intermediate_size=512, num_hidden_layers=5, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=4096,
                         rope_parameters={"rope_type": "default", "mrope_section": [8, 12, 12], "rope_theta": 1e6}),
        vision_config=dict(depth=1, embed_dim=32, hidden_size=64, num_heads=2, in_channels=3, patch_size=14,
                           spatial_merge_size=2, temporal_patch_size=2))
    cfg.text_config._attn_implementation = "sdpa"
    return Qwen2VLForConditionalGeneration(cfg).eval().to(torch.bfloat16).cuda()


def test_qwen2vl_prefill_matches_oracle_call_by_call():
    """The Qwen2-VL trio on the GPU: position embeddings are [3, 1, S, D] (M-RoPE, compacted along dim 2 as three
    planes of one aux tensor) and the importance comes from the last FOUR queries (modeling_qwen2_vl.py:289-301)."""
    from framefusion_b200.interface import apply_framefusion
    model = tiny_qwen2vl()
    apply_framefusion(model, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
    ff, llm = model.framefusion, model.model.language_model
    wl = synth.make_workload(10, 24, 256, torch.bfloat16, seed=13, r_lo=0.0, r_hi=0.5, n_pre=5, n_post=7, rot_dim=64)
    o = orc.OracleFrameFusion(0.3, 0.6, 0.1, "bf16")
    o.prepare(wl.patch_type.numpy(), wl.patch_num, *wl.prepare_args()[2:])
    calls, inner = [], ff.forward

    def checked(hidden, pos, mask, attn=None):
        assert pos[0].dim() == 4 and pos[0].shape[0] == 3
        h_in = t2f(hidden[0]); p_in = [t2f(pos[0][:, 0]), t2f(pos[1][:, 0])]
        a_in = None if attn is None else t2f(attn[0])
        if attn is not None:
            assert attn.shape[2] == 4
        out = inner(hidden, pos, mask, attn)
        want_h, want_p, _ = o.forward(h_in, [p[:, None] for p in p_in], None, a_in)
        assert np.array_equal(t2f(out[0][0]), want_h), f"call {len(calls)}: hidden_states differ from the oracle"
        for k in range(2):
            assert np.array_equal(t2f(out[1][k][:, 0]), np.asarray(want_p[k])[:, 0]), f"call {len(calls)}: position plane {k}"
        assert (ff.finish_merging, ff.finish_pruning) == (o.finish_merging, o.finish_pruning)
        calls.append((hidden.shape[1], out[0].shape[1], attn is not None))
        return out

    ff.forward = checked
    with torch.no_grad():
        ff.prepare(*synth.to_device(wl, "cuda").prepare_args())
        res = llm(inputs_embeds=wl.hidden.cuda(), use_cache=True)
    assert res.last_hidden_state.shape[1] == calls[-1][1] < wl.seq_len
    assert ff.finish_pruning and any(c[2] for c in calls)
    assert ff.sparsity_list == o.sparsity_list


# ---- f2: decode after a reduced prefill, by value ---------------------------------------------------------------------
class OracleOperator(torch.nn.Module):
    """The reference's algorithm (oracle/ff_oracle.py, numpy on the host — pinned against the unmodified reference) behind
    the same hooks: the checker for everything downstream of the operator — caches, masks, decode steps.  The tensors hop to
    the host and back around every call.  (The torch-op port on torch-CUDA is no checker here: it adds the members of a
    run with atomics in no fixed order, and torch-CPU's topk breaks ties at the k-th value its own way, which on a
    170-token sequence in bf16 decides several rows.)"""

    def __init__(self, cost, slb, rlb):
        super().__init__()
        from _harness import OracleAdapter
        object.__setattr__(self, "ad", OracleAdapter(cost, slb, rlb, "bf16"))

    def prepare(self, patch_type, *a):
        self.ad.prepare(patch_type.cpu(), *[int(x) if isinstance(x, torch.Tensor) else x for x in a])

    def forward(self, hidden, pos, mask, attn=None):
        dev = hidden.device
        cpu = lambda x: None if x is None else x.cpu()
        h, p, m = self.ad(hidden.cpu(), [cpu(pos[0]), cpu(pos[1])], cpu(mask), cpu(attn))
        pos[0], pos[1] = p[0].to(dev), p[1].to(dev)              # the hooks hand the same list from layer to layer
        return h.to(dev), pos, None if m is None else m.to(dev)

    finish_merging = property(lambda s: s.ad.finish_merging)
    finish_pruning = property(lambda s: s.ad.finish_pruning)
    sparsity_list = property(lambda s: s.ad.sparsity_list)


@pytest.mark.parametrize("lo,hi", [(0.0, 1.0), (0.0, 0.5)], ids=["mixed", "lowsim_prune"])
def test_prefill_then_decode_logits_match_the_reference_operator_behind_the_same_hooks(lo, hi):
    """The reference never compacts the KV cache: layer l keeps the keys of the tokens IT saw (modeling_qwen2.py:143-145),
    so after a reduced prefill the caches are ragged and a decode step attends to a different key set per layer.  Here the
    hooks leave the cache alone in the same way.  Proof by value: the same weights, once with the CUDA operator and once
    with the reference's algorithm (the oracle) behind the same hooks, give the same logits for the prefill and for four greedy
    decode steps, and the same cache contents layer by layer."""
    # ("port" below: the oracle operator)
    from framefusion_b200.interface import apply_framefusion
    wl = synth.make_workload(10, 24, 256, torch.bfloat16, seed=11, r_lo=lo, r_hi=hi, n_pre=5, n_post=7, rot_dim=64)
    args = synth.to_device(wl, "cuda").prepare_args()
    runs = []
    for kind in ("cuda", "port"):
        model = tiny_model()                                    # same seed: same weights
        apply_framefusion(model, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
        if kind == "port":
            op = OracleOperator(0.3, 0.6, 0.1)
            for m in [model, model.model] + list(model.model.layers) + [l.self_attn for l in model.model.layers]:
                m.framefusion = op
        ff = model.framefusion
        logits, tokens = [], []
        with torch.no_grad():
            ff.prepare(*[a.clone() if isinstance(a, torch.Tensor) else a for a in args])
            out = model(inputs_embeds=wl.hidden.cuda(), use_cache=True)
            cache = out.past_key_values
            lens = [cache.get_seq_length(i) for i in range(len(model.model.layers))]
            keys = [cache.layers[i].keys.clone() for i in range(len(model.model.layers))]
            logits.append(out.logits[:, -1].float().clone())
            for step in range(4):
                tok = logits[-1].argmax(-1, keepdim=True)
                tokens.append(int(tok))
                out = model(input_ids=tok, past_key_values=cache, use_cache=True)
                cache = out.past_key_values
                logits.append(out.logits[:, -1].float().clone())
        runs.append(dict(lens=lens, keys=keys, logits=logits, tokens=tokens, state=(ff.finish_merging, ff.finish_pruning, list(ff.sparsity_list))))
    a, b = runs
    assert a["lens"] == b["lens"] and wl.seq_len >= a["lens"][0] > a["lens"][-1]         # ragged, and identical
    assert all(x >= y for x, y in zip(a["lens"], a["lens"][1:]))
    assert a["state"] == b["state"]
    for i, (ka, kb) in enumerate(zip(a["keys"], b["keys"])):
        assert torch.equal(ka, kb), f"layer {i}: cached keys differ"
    assert a["tokens"] == b["tokens"]
    for i, (la, lb) in enumerate(zip(a["logits"], b["logits"])):
        assert torch.equal(la, lb), f"logits of step {i} differ (max {float((la - lb).abs().max())})"


# ---- f1: the Qwen2-VL embed-stage patch on the GPU ----------------------------------------------------------------------
def test_qwen2vl_video_prefill_through_the_embed_patch_matches_the_oracle():
    """``apply_framefusion`` on Qwen2VLForConditionalGeneration installs the top-level forward patch (reference
    models/qwenvl/modeling_qwen2_vl.py:117-138): the layout comes from ``input_ids`` / ``video_grid_thw``, nobody calls
    ``prepare`` by hand.  Every operator call of the prefill is checked against the numpy oracle prepared with the layout the
    reference's formula gives, and the decode step after it passes through."""
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    from framefusion_b200.interface import apply_framefusion
    torch.manual_seed(0)
    cfg = Qwen2VLConfig(
        text_config=dict(vocab_size=128, hidden_size=256, intermediate_size=512, num_hidden_layers=5, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=4096,
                         rope_parameters={"rope_type": "default", "mrope_section": [8, 12, 12], "rope_theta": 1e6}),
        vision_config=dict(depth=1, embed_dim=32, hidden_size=256, num_heads=2, in_channels=3, patch_size=14,
                           spatial_merge_size=2, temporal_patch_size=2),
        video_token_id=100, image_token_id=101, vision_start_token_id=102, vision_end_token_id=103)
    cfg.text_config._attn_implementation = "sdpa"
    model = Qwen2VLForConditionalGeneration(cfg).eval().to(torch.bfloat16).cuda()
    apply_framefusion(model, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
    ff = model.framefusion
    t, h, w = 8, 8, 12                                          # 8 frames of 24 tokens
    n_tok, patch_num = t * h * w // 4, h * w // 4
    ids = torch.tensor([[1, 2, 3, 4, 102] + [100] * n_tok + [103, 5, 6, 7, 8, 9]], device="cuda")
    # frames that resemble each other, so that the first layers have something to merge
    base = torch.randn(h * w, 3 * 2 * 14 * 14)
    pv = torch.cat([base + 0.05 * f * torch.randn_like(base) for f in range(t)]).to(torch.bfloat16).cuda()
    seq = ids.shape[1]
    o = orc.OracleFrameFusion(0.3, 0.6, 0.1, "bf16")
    want_pt = np.array([[-1] * 5 + list(range(patch_num)) * t + [-1] * 6])
    o.prepare(want_pt, patch_num, 5, 5 + n_tok - 1, n_tok, seq)
    calls, inner = [], ff.forward

    def checked(hidden, pos, mask, attn=None):
        h_in = t2f(hidden[0]); p_in = [t2f(pos[0][:, 0]), t2f(pos[1][:, 0])]
        a_in = None if attn is None else t2f(attn[0])
        out = inner(hidden, pos, mask, attn)
        want_h, want_p, _ = o.forward(h_in, [p[:, None] for p in p_in], None, a_in)
        fragile = o.last is not None and o.last.get("stage") == "merge" and bool(o.last["sim"].fragile.any())
        got = t2f(out[0][0])
        if got.shape == want_h.shape or not fragile:
            assert np.array_equal(got, want_h), f"call {len(calls)}: hidden_states differ from the oracle"
        assert (ff.finish_merging, ff.finish_pruning) == (o.finish_merging, o.finish_pruning)
        calls.append((hidden.shape[1], out[0].shape[1]))
        return out

    ff.forward = checked
    with torch.no_grad():
        out = model(input_ids=ids, pixel_values_videos=pv, video_grid_thw=torch.tensor([[t, h, w]], device="cuda"),
                    mm_token_type_ids=(ids == 100).int() * 2, use_cache=True)
    assert torch.equal(ff.patch_type.cpu() if ff.patch_type.shape[1] == seq else torch.tensor(want_pt), torch.tensor(want_pt)) or calls[0][1] < seq
    assert (model.patch_num, int(model.image_token_start_index), model.image_token_length, model.original_length) == (patch_num, 5, n_tok, seq)
    assert calls and calls[0][0] == seq and out.logits.shape[1] == calls[-1][1] < seq
    assert torch.isfinite(out.logits.float()).all()
    with torch.no_grad():
        step = model(input_ids=out.logits[:, -1].argmax(-1, keepdim=True), past_key_values=out.past_key_values, use_cache=True)
    assert step.logits.shape[1] == 1 and torch.isfinite(step.logits.float()).all()


# ---- config 5: the decoder split by layers over two GPUs -------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (the driver's 1-GPU tier skips it)")
def test_layer_split_over_two_gpus_matches_one_gpu():
    """``dispatch.split_layers``: the same weights on one GPU and split over two give the same reduced prefill — the
    operator keeps a context and a workspace per device, rebuilds its chain links where the sequence changes device, and the
    importance / prune stages run wherever their layer lives."""
    from framefusion_b200.interface import apply_framefusion
    from framefusion_b200.dispatch import split_layers
    wl = synth.make_workload(10, 24, 256, torch.bfloat16, seed=11, r_lo=0.0, r_hi=0.5, n_pre=5, n_post=7, rot_dim=64)
    outs = []
    for devices in (["cuda:0"], ["cuda:0", "cuda:1"]):
        model = tiny_model()
        apply_framefusion(model, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
        if len(devices) > 1:
            placement = split_layers(model.model, devices)
            assert {str(d) for d in placement} == {"cuda:0", "cuda:1"}
        ff = model.framefusion
        with torch.no_grad():
            ff.prepare(*synth.to_device(wl, "cuda:0").prepare_args())
            res = model.model(inputs_embeds=wl.hidden.to("cuda:0"), use_cache=True)
        outs.append((res.last_hidden_state.cpu(), list(ff.sparsity_list), ff.finish_merging, ff.finish_pruning,
                     [res.past_key_values.get_seq_length(i) for i in range(len(model.model.layers))]))
        assert torch.cuda.current_device() == 0                 # the entry points restore the caller's device
    a, b = outs
    assert a[1:] == b[1:]
    assert torch.equal(a[0], b[0])


@pytest.mark.parametrize("mode", ["default", "frame"])
def test_two_operators_on_two_threads_and_streams(mode):
    """Two models served from two host threads, each on its own stream (include/framefusion_b200.h: a context per operator and
    device, the stream passed per call, ``ff_last_error`` thread-local): the interleaved runs give bit for bit what each
    operator gives alone.  "frame": the first merge call of both runs as the frame-pipelined kernel (96 + 60 persistent CTAs
    that wait for each other through global flags: more than the 148 SMs hold at once)."""
    import threading
    from framefusion_b200.main import FrameFusion
    from framefusion_b200.utils import scaled_dot_product_attention

    def make(seed, frames, patches, lo, hi):
        wl = synth.to_device(synth.make_workload(frames, patches, 1024, torch.bfloat16, seed=seed, r_lo=lo, r_hi=hi), "cuda")
        q, k = synth.make_attention_inputs(wl.seq_len, 28, 4, 128, torch.bfloat16, seed=seed)
        return wl, q[:, :, -1:, :].contiguous().cuda(), k.cuda()

    def run(ff, wl, q_last, keys):
        ff.prepare(*wl.prepare_args())
        h, pos = wl.hidden, [wl.cos, wl.sin]
        guard = 0
        while not ff.finish_merging and guard < 8:
            h, pos, _ = ff(h, pos, None)
            guard += 1
        if not ff.finish_pruning:
            attn = scaled_dot_product_attention(q_last, keys[:, :, :h.shape[1]], None, num=1, is_causal=True, enable_gqa=True)
            h, pos, _ = ff(h, pos, None, attn)
        return h.clone(), pos[0].clone(), list(ff.sparsity_list)

    def operator():
        ff = FrameFusion(0.3, 0.6, 0.1)
        if mode == "frame":
            ff.use_frame = "force"
        return ff

    jobs = [make(21, 24, 96, 0.0, 1.0), make(22, 32, 60, 0.0, 0.5)]
    alone = [run(operator(), *j) for j in jobs]
    torch.cuda.synchronize()
    results, errors = [None, None], []

    def worker(i):
        try:
            stream = torch.cuda.Stream()
            ff = operator()
            with torch.cuda.stream(stream):
                for _ in range(12):
                    out = run(ff, *jobs[i])
                stream.synchronize()
            results[i] = out
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    for got, want in zip(results, alone):
        assert got is not None and got[2] == want[2]
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
