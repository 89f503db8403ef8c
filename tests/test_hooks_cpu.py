"""CPU: the Qwen2 hook trio and ``apply_framefusion`` on a tiny random-init decoder.

The operator itself is CUDA-only, so the patched model is driven here with the numpy oracle standing in for it
(tests may use the oracle as a checker): what is verified is the PLUMBING the reference's hooks define
(models/qwen2/modeling_qwen2.py:44-47, 66-68, 84-86, 166-178, 262-266, 303-306) — where the operator is called,
that the compacted position embeddings / mask travel from layer to layer, that importance is produced only while
pruning is armed, that the KV cache keeps per-layer ragged lengths and a decode step passes through."""
import numpy as np
import pytest
import torch
from transformers import Qwen2Config, Qwen2ForCausalLM

from _harness import OracleAdapter
from oracle import ff_torch_port as port
from framefusion_b200 import synth


def tiny_model(attn="sdpa"):
    torch.manual_seed(0)
    cfg = Qwen2Config(vocab_size=128, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=4096, rope_theta=1e6)
    cfg._attn_implementation = attn
    return Qwen2ForCausalLM(cfg).eval().float()


class OracleOperator(torch.nn.Module):
    """OracleAdapter with the attribute surface the hooks read."""

    def __init__(self, cost, slb, rlb):
        super().__init__()
        object.__setattr__(self, "ad", OracleAdapter(cost, slb, rlb, "f32"))
        self.calls = []

    def prepare(self, *a):
        self.ad.prepare(*a)

    def forward(self, hidden, pos, mask, attn=None):
        self.calls.append((hidden.shape[1], attn is not None))
        return self.ad(hidden, pos, mask, attn)

    finish_merging = property(lambda s: s.ad.finish_merging)
    finish_pruning = property(lambda s: s.ad.finish_pruning)
    sparsity_list = property(lambda s: s.ad.sparsity_list)


def install(model, op):
    """apply_framefusion, then swap the CUDA operator for the stand-in on every module that holds it."""
    from framefusion_b200.interface import apply_framefusion
    apply_framefusion(model, cost=0.3, similarity_lower_bound=0.6, ratio_lower_bound=0.1)
    shared = model.framefusion
    assert type(shared).__name__ == "FrameFusion"
    for m in [model, model.model] + list(model.model.layers) + [l.self_attn for l in model.model.layers]:
        assert m.framefusion is shared                     # ONE shared operator (interface.py:187-212)
        m.framefusion = op


@pytest.fixture()
def patched_importance(monkeypatch):
    import framefusion_b200.hooks.qwen2 as hk

    def cpu_importance(q, k, v, num=1, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False):
        return port.last_query_attention(q, k, num=num, is_causal=is_causal, scale=scale)
    monkeypatch.setattr(hk, "scaled_dot_product_attention", cpu_importance)


def workload(frames=6, patches=12, hidden=64, lo=0.0, hi=1.0, seed=4):
    return synth.make_workload(frames, patches, hidden, torch.float32, seed=seed, r_lo=lo, r_hi=hi, n_pre=3, n_post=5,
                               rot_dim=16)


def manual_reference(model, wl, op):
    """The schedule of the reference hooks, spelled out with stock modules (no patched forwards)."""
    m = model.model
    h = wl.hidden.clone()
    S = h.shape[1]
    pos_ids = torch.arange(S)[None]
    pe = list(m.rotary_emb(h, pos_ids))
    op.prepare(*wl.prepare_args())
    from transformers.models.qwen2.modeling_qwen2 import apply_rotary_pos_emb
    for li, layer in enumerate(m.layers):
        if li == 0:
            h, pe, _ = op(h, pe, None)
        res = h
        x = layer.input_layernorm(h)
        att = layer.self_attn
        shp = (*x.shape[:-1], -1, att.head_dim)
        q = att.q_proj(x).view(shp).transpose(1, 2)
        k = att.k_proj(x).view(shp).transpose(1, 2)
        v = att.v_proj(x).view(shp).transpose(1, 2)
        q, k = apply_rotary_pos_emb(q, k, pe[0], pe[1])
        w = None
        if op.finish_merging and not op.finish_pruning:
            w = port.last_query_attention(q, k, num=1, is_causal=True, scale=att.scaling)
        kk = k.repeat_interleave(att.num_key_value_groups, dim=1)
        vv = v.repeat_interleave(att.num_key_value_groups, dim=1)
        o = torch.nn.functional.scaled_dot_product_attention(q, kk, vv, is_causal=True, scale=att.scaling)
        o = att.o_proj(o.transpose(1, 2).reshape(*x.shape[:-1], -1))
        h = res + o
        h, pe, _ = op(h, pe, None, w)
        h = h + layer.mlp(layer.post_attention_layernorm(h))
    return m.norm(h)


@pytest.mark.parametrize("lo,hi", [(0.0, 1.0), (0.0, 0.5), (0.8, 1.0)], ids=["mixed", "lowsim_prune", "topk"])
def test_hook_schedule_matches_manual_reference(patched_importance, lo, hi):
    model = tiny_model()
    wl = workload(lo=lo, hi=hi)
    with torch.no_grad():
        want = manual_reference(model, wl, OracleOperator(0.3, 0.6, 0.1))
        op = OracleOperator(0.3, 0.6, 0.1)
        install(model, op)
        model.framefusion.prepare(*wl.prepare_args())
        out = model.model(inputs_embeds=wl.hidden.clone(), use_cache=True)
    got = out.last_hidden_state
    assert got.shape == want.shape and got.shape[1] < wl.seq_len
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)
    # call #0 before layer-0 attention without importance, then one call per layer (modeling_qwen2.py:44-47, 66-68)
    assert len(op.calls) == 1 + len(model.model.layers)
    assert op.calls[0] == (wl.seq_len, False)
    # importance only while pruning is armed (modeling_qwen2.py:168)
    armed = [w for (_s, w) in op.calls]
    assert sum(armed) <= 1
    # the KV cache keeps, per layer, the length that layer SAW (modeling_qwen2.py:143-145): non-increasing, ragged
    lens = [out.past_key_values.get_seq_length(i) for i in range(len(model.model.layers))]
    assert lens[0] < wl.seq_len and all(a >= b for a, b in zip(lens, lens[1:])) and lens[-1] >= got.shape[1]


def test_decode_step_passes_through(patched_importance):
    model = tiny_model()
    wl = workload()
    op = OracleOperator(0.3, 0.6, 0.1)
    install(model, op)
    with torch.no_grad():
        model.framefusion.prepare(*wl.prepare_args())
        out = model.model(inputs_embeds=wl.hidden.clone(), use_cache=True)
        n_calls = len(op.calls)
        lens = [out.past_key_values.get_seq_length(i) for i in range(4)]
        step = model.model(inputs_embeds=torch.randn(1, 1, 64), past_key_values=out.past_key_values, use_cache=True,
                           attention_mask=torch.ones(1, wl.seq_len + 1, dtype=torch.long))   # generate() style mask
    assert step.last_hidden_state.shape == (1, 1, 64)
    assert [c[0] for c in op.calls[n_calls:]] == [1] * 5                       # q_len == 1: the operator is a no-op
    assert [step.past_key_values.get_seq_length(i) for i in range(4)] == [l + 1 for l in lens]


def test_eager_attention_compacts_the_4d_mask(patched_importance):
    model = tiny_model("eager")
    wl = workload()
    op = OracleOperator(0.3, 0.6, 0.1)
    install(model, op)
    with torch.no_grad():
        model.framefusion.prepare(*wl.prepare_args())
        out = model.model(inputs_embeds=wl.hidden.clone(), use_cache=False)
    assert out.last_hidden_state.shape[1] < wl.seq_len


def test_eager_decode_after_a_reduced_prefill(patched_importance):
    """Eager attention adds the 4-D mask without slicing it: after a reduced prefill the caches of later layers are shorter
    than layer 0's, which the mask is built for — the attention hook slices it like modeling_qwen2.py:150-152.  The decode
    step must also equal the sdpa path's (same weights, same caches)."""
    outs = {}
    for attn in ("eager", "sdpa"):
        model = tiny_model(attn)
        wl = workload()
        op = OracleOperator(0.3, 0.6, 0.1)
        install(model, op)
        with torch.no_grad():
            model.framefusion.prepare(*wl.prepare_args())
            out = model.model(inputs_embeds=wl.hidden.clone(), use_cache=True)
            lens = [out.past_key_values.get_seq_length(i) for i in range(4)]
            assert lens[0] > lens[-1]                          # ragged: the case the slicing is for
            torch.manual_seed(1)
            step = model.model(inputs_embeds=torch.randn(1, 1, 64), past_key_values=out.past_key_values, use_cache=True)
        outs[attn] = step.last_hidden_state
        assert [step.past_key_values.get_seq_length(i) for i in range(4)] == [l + 1 for l in lens]
    assert torch.allclose(outs["eager"], outs["sdpa"], rtol=1e-4, atol=1e-5)


def test_unsupported_model_raises_like_the_reference(capsys):
    from framefusion_b200.interface import apply_framefusion
    with pytest.raises(NotImplementedError):
        apply_framefusion(torch.nn.Linear(2, 2), 0.3, 0.6, 0.1)
    assert "Model not supported" in capsys.readouterr().out


def test_get_attr_by_name():
    from framefusion_b200.utils import get_attr_by_name
    model = tiny_model()
    assert get_attr_by_name(model, "model.layers.1.self_attn.q_proj") is model.model.layers[1].self_attn.q_proj


# ---- Qwen2-VL: M-RoPE (4-D cos / sin), importance from the last four queries -------------------------------
def tiny_qwen2vl():
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    torch.manual_seed(0)
    cfg = Qwen2VLConfig(
        text_config=dict(vocab_size=128, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=4096,
                         rope_parameters={"rope_type": "default", "mrope_section": [2, 3, 3], "rope_theta": 1e6}),
        vision_config=dict(depth=1, embed_dim=32, hidden_size=64, num_heads=2, in_channels=3, patch_size=14,
                           spatial_merge_size=2, temporal_patch_size=2))
    return Qwen2VLForConditionalGeneration(cfg).eval().float()


def test_qwen2vl_trio_runs_with_mrope_and_four_query_importance(monkeypatch):
    import framefusion_b200.hooks.qwen2_vl as hk
    seen = []

    def cpu_importance(q, k, v, num=1, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False):
        seen.append(num)
        return port.last_query_attention(q, k, num=num, is_causal=is_causal, scale=scale)
    monkeypatch.setattr(hk, "scaled_dot_product_attention", cpu_importance)
    from framefusion_b200.interface import apply_framefusion
    model = tiny_qwen2vl()
    apply_framefusion(model, 0.3, 0.6, 0.1)
    llm = model.model.language_model
    shared = model.framefusion
    op = OracleOperator(0.3, 0.6, 0.1)
    for m in [model, llm] + list(llm.layers) + [l.self_attn for l in llm.layers]:
        assert m.framefusion is shared
        m.framefusion = op
    wl = workload(lo=0.0, hi=0.5)                           # little to merge: pruning gets armed
    pos_cache = {}
    inner = op.forward

    def spy(hidden, pos, mask, attn=None):
        pos_cache["in"] = [p.shape for p in pos]
        out = inner(hidden, pos, mask, attn)
        pos_cache["out"] = [p.shape for p in out[1]]
        return out
    op.forward = spy
    with torch.no_grad():
        op.prepare(*wl.prepare_args())
        out = llm(inputs_embeds=wl.hidden.clone(), use_cache=True)
    assert out.last_hidden_state.shape[1] < wl.seq_len
    assert len(pos_cache["in"][0]) == 4 and pos_cache["in"][0][0] == 3            # [3, B, S, D]
    assert pos_cache["out"][0][2] == out.last_hidden_state.shape[1]               # compacted along dim 2
    assert seen and all(n == 4 for n in seen)                                     # reference :292-300


def test_layout_builders_match_the_reference_formulas():
    from framefusion_b200.layout import qwen2vl_prepare_args, llava_video_prepare_args
    # Qwen2-VL (models/qwenvl/modeling_qwen2_vl.py:118-127): 3 frames of a 4x6 grid merged 2x2 -> 6 tokens per frame
    ids = torch.tensor([[11, 12] + [99] * 18 + [13, 14, 15]])
    pt, patch_num, start, end, length, orig = qwen2vl_prepare_args(ids, 99, torch.tensor([[3, 4, 6]]), 2)
    want = [-1] * 2 + list(range(6)) * 3 + [-1] * (23 - 19 - 1)
    assert (patch_num, start, end, length, orig) == (6, 2, 19, 18, 23) and pt.tolist() == [want]
    # LLaVA-Video (models/llava_video/modeling_llava_video.py:322-336): 27 patches per side, bilinear pool -> 14 * 15
    ids = torch.tensor([[1, 2, 3, -200, 4, 5]])
    pt, patch_num, start, end, length, orig = llava_video_prepare_args(ids, -200, 210 * 4, 27)
    assert patch_num == 210 and (start, end, length, orig) == (3, 3 + 840 - 1, 840, 6 + 840 - 1)
    assert pt.tolist() == [[-1] * 3 + list(range(210)) * 4 + [-1] * (orig - end - 1)]


def test_layer_split_moves_inputs_between_blocks(patched_importance):
    """dispatch.split_layers: contiguous layer blocks and input-moving pre-hooks (what accelerate's device_map does
    for the reference).  On a CPU-only machine both "devices" are the CPU: the hooks must be transparent."""
    from framefusion_b200.dispatch import layer_devices, split_layers
    assert [str(d) for d in layer_devices(7, ["cuda:0", "cuda:1", "cuda:2"])] == ["cuda:0"] * 3 + ["cuda:1"] * 2 + ["cuda:2"] * 2
    model = tiny_model()
    wl = workload()
    with torch.no_grad():
        want = manual_reference(model, wl, OracleOperator(0.3, 0.6, 0.1))
        op = OracleOperator(0.3, 0.6, 0.1)
        install(model, op)
        placement = split_layers(model.model, ["cpu", "cpu"])
        assert len(placement) == len(model.model.layers)
        model.framefusion.prepare(*wl.prepare_args())
        got = model.model(inputs_embeds=wl.hidden.clone(), use_cache=True).last_hidden_state
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5)


# ---- Qwen2-VL embed-stage patch (reference models/qwenvl/modeling_qwen2_vl.py:117-138) and the drop-in import path ----
def tiny_qwen2vl_video():
    """The tiny Qwen2-VL with in-vocabulary placeholder ids, and one video: grid (T, H, W) = (4, 4, 6) merged 2 x 2 ->
    4 frames of 6 tokens between 3 leading and 4 trailing text tokens."""
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    torch.manual_seed(0)
    cfg = Qwen2VLConfig(
        text_config=dict(vocab_size=128, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=4096,
                         rope_parameters={"rope_type": "default", "mrope_section": [2, 3, 3], "rope_theta": 1e6}),
        vision_config=dict(depth=1, embed_dim=32, hidden_size=64, num_heads=2, in_channels=3, patch_size=14,
                           spatial_merge_size=2, temporal_patch_size=2),
        video_token_id=100, image_token_id=101, vision_start_token_id=102, vision_end_token_id=103)
    model = Qwen2VLForConditionalGeneration(cfg).eval().float()
    t, h, w = 4, 4, 6
    n_tok = t * h * w // 4
    ids = torch.tensor([[1, 2, 102] + [100] * n_tok + [103, 5, 6, 7]])
    inputs = dict(input_ids=ids, pixel_values_videos=torch.randn(t * h * w, 3 * 2 * 14 * 14),
                  video_grid_thw=torch.tensor([[t, h, w]]), mm_token_type_ids=(ids == 100).int() * 2)
    return model, inputs, n_tok


class RecordingOperator(OracleOperator):
    def prepare(self, *a):
        self.prepared = a
        super().prepare(*a)


def _swap_operator(model, op):
    llm = model.model.language_model
    shared = model.framefusion
    for m in [model, llm] + list(llm.layers) + [l.self_attn for l in llm.layers]:
        assert m.framefusion is shared
        m.framefusion = op


def test_qwen2vl_embed_patch_prepares_the_layout_from_the_video_grid(monkeypatch):
    """apply_framefusion installs the top-level forward patch: a prefill with video_grid_thw reaches ``prepare`` with the
    reference's arguments (:118-137), leaves them on the model (:129-133), reduces the sequence, and a decode step passes
    through without another ``prepare``."""
    import framefusion_b200.hooks.qwen2_vl as hk
    monkeypatch.setattr(hk, "scaled_dot_product_attention",
                        lambda q, k, v, num=1, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False:
                        port.last_query_attention(q, k, num=num, is_causal=is_causal, scale=scale))
    from framefusion_b200.interface import apply_framefusion
    model, inputs, n_tok = tiny_qwen2vl_video()
    apply_framefusion(model, 0.3, 0.6, 0.1)
    assert model.forward.__func__ is hk.forward
    op = RecordingOperator(0.3, 0.6, 0.1)
    _swap_operator(model, op)
    with torch.no_grad():
        out = model(**inputs, use_cache=True)
    pt, patch_num, start, end, length, orig = op.prepared
    seq = inputs["input_ids"].shape[1]
    assert (patch_num, int(start), int(end), length, orig) == (6, 3, 3 + n_tok - 1, n_tok, seq)
    assert pt.tolist() == [[-1] * 3 + list(range(6)) * 4 + [-1] * 4]
    assert (model.patch_num, int(model.image_token_start_index), model.image_token_length, model.original_length) == (6, 3, n_tok, seq)
    assert op.calls and out.logits.shape[1] < seq                                 # the prefill was reduced
    cache = out.past_key_values
    assert cache.get_seq_length(0) == seq and cache.get_seq_length(3) == out.logits.shape[1]
    del op.prepared
    n_calls = len(op.calls)
    with torch.no_grad():
        step = model(input_ids=torch.tensor([[9]]), past_key_values=cache, use_cache=True)
    assert step.logits.shape[1] == 1 and not hasattr(op, "prepared")
    assert all(c[0] == 1 for c in op.calls[n_calls:])                             # q_len == 1: the operator is a no-op


def test_get_token_type_installs_only_the_embed_patch():
    """reference interface.py:140-166: the layout is derived and left on the model, the decoder keeps its own forwards; a
    model that carries a ``mode`` attribute is not prepared (:136)."""
    import framefusion_b200.hooks.qwen2_vl as hk
    from framefusion_b200.interface import get_token_type
    model, inputs, n_tok = tiny_qwen2vl_video()
    llm_forward = model.model.language_model.forward
    get_token_type(model)
    assert model.forward.__func__ is hk.forward and model.model.language_model.forward == llm_forward
    model.mode = "token_type_only"
    with torch.no_grad():
        out = model(**inputs)
    assert out.logits.shape[1] == inputs["input_ids"].shape[1]                    # dense: nothing was reduced
    assert (model.patch_num, model.image_token_length) == (6, n_tok)
    with pytest.raises(NotImplementedError):
        get_token_type(torch.nn.Linear(2, 2))
    with pytest.raises(NotImplementedError, match="layout"):
        get_token_type(tiny_model())                                              # Qwen2: its embed patch is third-party code


def test_drop_in_import_path_resolves_to_this_implementation():
    """reference README.md:123 / example_llava.py:136: ``from framefusion.interface import apply_framefusion``."""
    import framefusion
    import framefusion.interface as fi
    import framefusion.main as fm
    import framefusion.utils as fu
    import framefusion.models.qwen2.modeling_qwen2 as q2
    import framefusion.models.qwen2.modeling_qwen2_vl as q2vl
    import framefusion.models.qwenvl.modeling_qwen2_vl as qvl
    import framefusion_b200.interface as bi
    import framefusion_b200.main as bm
    assert fi.apply_framefusion is bi.apply_framefusion and fi.replace_framefusion_forward is bi.replace_framefusion_forward
    assert fi.get_token_type is bi.get_token_type and fm.FrameFusion is bm.FrameFusion and framefusion.FrameFusion is bm.FrameFusion
    assert fu.TEXT_TOKEN == -1 and fu.IGNORE_TOKEN == -2 and callable(fu.scaled_dot_product_attention) and callable(fu.get_attr_by_name)
    assert callable(fm.find_contigious_latter_index) and callable(fm.cosine_similarity)
    for mod, names in ((q2, ("Qwen2Model_merge_then_fastv_cost_given_forward", "Qwen2DecoderLayer_merge_then_prune_by_cost_forward",
                             "Qwen2SdpaAttention_merge_then_prune_by_cost_forward")),
                       (q2vl, ("Qwen2VLModel_merge_then_fastv_cost_given_forward", "Qwen2VLDecoderLayer_merge_then_fastv_cost_given_forward",
                               "Qwen2VLSdpaAttention_merge_then_fastv_cost_given_forward")),
                       (qvl, ("forward",))):
        for n in names:
            assert callable(getattr(mod, n))
    model = tiny_model()
    fi.apply_framefusion(model, 0.3, 0.6, 0.1)                                    # the reference's call, unchanged
    assert type(model.framefusion) is bm.FrameFusion
