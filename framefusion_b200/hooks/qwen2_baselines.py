"""Hooks of the fixed-amount comparison methods on a plain Qwen2 decoder stack (transformers 5.x module API).

Reference: ``/root/reference/framefusion/models/qwen2/modeling_qwen2_baseline.py`` — meta interface
``replace_Qwen2_forward(model, mode, **kwargs)`` (:45-109) and the per-method installers ``replace_Qwen2_fastv``
(:175-188), ``replace_Qwen2_merging`` (:860-874), ``replace_Qwen2_merge_then_fastv`` (:1339-1355),
``replace_Qwen2_fastv_then_merge`` (:2055-2069), ``replace_Qwen2_streamingllm`` (:579-590).  Same names, same keyword arguments and defaults; one
``TokenReductionBaseline`` operator (``framefusion_b200/baselines.py``) hangs on the model as ``model.baseline`` and
does every tensor operation through the C ABI.

What the three patched forwards add to the stock ones:

* model: in front of every decoder layer the operator may merge (fixed sparsity of that layer, reference :916-920) and,
  in front of layer ``fastv_k``, prune with the attention of layer ``fastv_k - 1`` (reference :308-342); ``position_ids``
  become the kept indices (:337), position embeddings and mask travel compacted;
* decoder layer: returns ``(hidden_states, last_query_attention_or_None)`` (reference :566-571);
* attention: produces the last query's probabilities at layer ``fastv_k - 1`` only (reference :474-492).

The caller fills the token layout before the prefill, as the reference's embed-stage patches do on ``model``
(``model.image_token_start_index`` etc., reference :303-304): ``model.baseline.prepare(...)``.
"""
from __future__ import annotations

from types import MethodType
from typing import Optional

import torch
from transformers.cache_utils import Cache

from ..baselines import TokenReductionBaseline, compute_density_overhead
from ..utils import get_attr_by_name
from .qwen2 import attention_with_importance, model_inputs, model_outputs


def Qwen2DecoderLayer_fastv_forward(
    self,
    hidden_states: torch.Tensor,
    attention_mask: Optional[torch.Tensor] = None,
    position_ids: Optional[torch.LongTensor] = None,
    past_key_values: Optional[Cache] = None,
    use_cache: Optional[bool] = False,
    position_embeddings=None,
    **kwargs,
):
    """-> ``(hidden_states, self_attn_weights)``; the second is None except at layer ``fastv_k - 1``."""
    if "past_key_value" in kwargs:                       # 4.45 spelling used by the reference's callers
        past_key_values = kwargs.pop("past_key_value")
    for stale in ("output_attentions", "cache_position"):
        kwargs.pop(stale, None)
    residual = hidden_states
    hidden_states = self.input_layernorm(hidden_states)
    hidden_states, self_attn_weights = self.self_attn(
        hidden_states=hidden_states,
        attention_mask=attention_mask,
        position_ids=position_ids,
        past_key_values=past_key_values,
        use_cache=use_cache,
        position_embeddings=position_embeddings,
        **kwargs,
    )
    hidden_states = residual + hidden_states
    residual = hidden_states
    hidden_states = self.post_attention_layernorm(hidden_states)
    hidden_states = self.mlp(hidden_states)
    hidden_states = residual + hidden_states
    return (hidden_states, self_attn_weights)


def Qwen2SdpaAttention_fastv_forward(
    self,
    hidden_states: torch.Tensor,
    position_embeddings=None,
    attention_mask: Optional[torch.Tensor] = None,
    past_key_values: Optional[Cache] = None,
    **kwargs,
):
    """-> ``(attn_output, last_query_attention_or_None)`` (reference :474-492: only layer ``fastv_k - 1`` of a prefill)."""
    want = hidden_states.shape[1] > 1 and self.baseline.wants_attention(self.layer_idx)
    return attention_with_importance(self, hidden_states, position_embeddings, attention_mask, past_key_values, want, **kwargs)


def Qwen2Model_fastv_forward(
    self,
    input_ids: Optional[torch.LongTensor] = None,
    attention_mask: Optional[torch.Tensor] = None,
    position_ids: Optional[torch.LongTensor] = None,
    past_key_values: Optional[Cache] = None,
    inputs_embeds: Optional[torch.FloatTensor] = None,
    use_cache: Optional[bool] = None,
    output_attentions: Optional[bool] = None,
    output_hidden_states: Optional[bool] = None,
    return_dict: Optional[bool] = None,
    cache_position: Optional[torch.LongTensor] = None,
    **kwargs,
):
    op = self.baseline
    hidden_states, position_ids, past_key_values, masks, position_embeddings, use_cache = model_inputs(
        self, input_ids, attention_mask, position_ids, past_key_values, inputs_embeds, use_cache)
    prefill = hidden_states.shape[1] > 1
    if prefill and op.fastv_k is not None and not use_cache:
        raise NotImplementedError("fastv only support use_cache=True")          # reference :345-346

    all_hidden_states = () if output_hidden_states else None
    last_attention = None
    for i, decoder_layer in enumerate(self.layers[: self.config.num_hidden_layers]):
        if output_hidden_states:
            all_hidden_states += (hidden_states,)
        kind = self.config.layer_types[i]
        if prefill:
            q_before = hidden_states.shape[1]
            if op.fastv_k is not None and i == op.fastv_k and i > 0:
                hidden_states, position_embeddings, masks[kind] = op.fastv_at(
                    i, hidden_states, position_embeddings, masks[kind], last_attention)
                position_ids = op.keep_indexs().unsqueeze(0)                     # reference :337
            hidden_states, position_embeddings, masks[kind] = op.merge_at(i, hidden_states, position_embeddings, masks[kind])
            if hidden_states.shape[1] != q_before:
                if position_ids.shape[-1] != hidden_states.shape[1]:
                    position_ids = position_ids[..., : hidden_states.shape[1]]   # (merges: only the length matters downstream)
                for other in masks:
                    if other != kind and masks[other] is not None:
                        raise NotImplementedError("token reduction with a 4-D mask per layer kind is not supported")
        layer_outputs = decoder_layer(
            hidden_states,
            attention_mask=masks[kind],
            position_embeddings=position_embeddings,
            position_ids=position_ids,
            past_key_values=past_key_values,
            use_cache=use_cache,
            **kwargs,
        )
        hidden_states = layer_outputs[0]
        last_attention = layer_outputs[1]

    return model_outputs(self, hidden_states, past_key_values, use_cache, all_hidden_states, return_dict)


# fixed-sparsity merging and the combination run through the same three forwards: what happens is the operator's setting
Qwen2Model_merging_forward = Qwen2Model_fastv_forward
Qwen2DecoderLayer_merging_forward = Qwen2DecoderLayer_fastv_forward
Qwen2SdpaAttention_merging_forward = Qwen2SdpaAttention_fastv_forward
Qwen2Model_merge_then_fastv_forward = Qwen2Model_fastv_forward
Qwen2DecoderLayer_merge_then_fastv_forward = Qwen2DecoderLayer_fastv_forward
Qwen2SdpaAttention_merge_then_fastv_forward = Qwen2SdpaAttention_fastv_forward


def _install(model, operator: TokenReductionBaseline, llm_key: str = "model"):
    """``llm_key``: dotted path of the Qwen2 decoder stack — ``model`` for LLaVA-Video / Qwen2 (reference :179), ``llm.model``
    for the MiniCPM-V and NVILA wrappers (reference :194, :209)."""
    llm = get_attr_by_name(model, llm_key)
    names = {c.__name__ for c in type(llm).__mro__}
    if "Qwen2Model" not in names:
        raise TypeError("language model is not Qwen2.")                           # reference :187-188
    model.baseline = operator
    llm.baseline = operator
    llm.forward = MethodType(Qwen2Model_fastv_forward, llm)
    for layer in llm.layers:
        layer.baseline = operator
        layer.forward = MethodType(Qwen2DecoderLayer_fastv_forward, layer)
        layer.self_attn.baseline = operator
        layer.self_attn.forward = MethodType(Qwen2SdpaAttention_fastv_forward, layer.self_attn)
    return operator


def replace_Qwen2_fastv(model, fastv_k=3, fastv_r=0.5):
    """reference :175-188"""
    model.fastv_k = fastv_k
    model.fastv_r = fastv_r
    return _install(model, TokenReductionBaseline(None, fastv_k, fastv_r))


def replace_Qwen2_merging(model, sparsity=[0.1] * 28):
    """reference :860-874"""
    model.sparsity = sparsity
    return _install(model, TokenReductionBaseline(sparsity, None))


def replace_Qwen2_merge_then_fastv(model, sparsity=[0.1] * 28, fastv_k=3, fastv_r=0.5):
    """reference :1339-1355"""
    model.sparsity = sparsity
    model.fastv_k = fastv_k
    model.fastv_r = fastv_r
    return _install(model, TokenReductionBaseline(sparsity, fastv_k, fastv_r))


def replace_Qwen2_fastv_then_merge(model, fastv_k=2, fastv_r=0.75, merging_sparsity=0.3):
    """reference :2055-2069: FastV in front of layer ``fastv_k``, ONE merge of ``merging_sparsity`` at layer ``fastv_k + 1``
    (:2283-2285)."""
    model.fastv_k = fastv_k
    model.fastv_r = fastv_r
    model.merging_sparsity = merging_sparsity
    return _install(model, TokenReductionBaseline([0.0] * (fastv_k + 1) + [merging_sparsity], fastv_k, fastv_r))


def replace_Qwen2_streamingllm(model, init_num=4, length_rate=0.3):
    """reference :579-590 — an attention kernel of the un-vendored ``minference`` package (optional import in the
    reference, :13-17), not a token reduction: nothing of it is on this path."""
    raise NotImplementedError(
        "StreamingLLM replaces the attention kernel (minference.streaming_forward, not installed) and reduces no tokens; "
        "it is outside the merge / prune path this package implements")


def replace_minicpmv_fastv(model, fastv_k=3, fastv_r=0.5):
    """reference :190-203 (the decoder stack sits under ``model.llm.model``)"""
    model.fastv_k = fastv_k
    model.fastv_r = fastv_r
    return _install(model, TokenReductionBaseline(None, fastv_k, fastv_r), llm_key="llm.model")


def replace_nvila_fastv(model, fastv_k=3, fastv_r=0.5):
    """reference :205-218"""
    return replace_minicpmv_fastv(model, fastv_k, fastv_r)


def replace_minicpmv_streamingllm(model, init_num=4, length_rate=0.3):
    """reference :592-603"""
    return replace_Qwen2_streamingllm(model, init_num, length_rate)


def replace_nvila_streamingllm(model, init_num=4, length_rate=0.3):
    """reference :605-616"""
    return replace_Qwen2_streamingllm(model, init_num, length_rate)


def _wrapper_forward(name, installers, model, mode, kwargs):
    print(f"{name} mode: {mode} and kwargs: {kwargs}")
    if mode == "fastv":
        cfg = {"fastv_k": kwargs.get("fastv_k", 3), "fastv_r": kwargs.get("fastv_r", 0.5)}
        print(f"Config\n{cfg}")
        return installers[0](model, **cfg)
    if mode == "streamingllm":
        cfg = {"init_num": kwargs.get("init_num", 8), "length_rate": kwargs.get("length_rate", 0.3)}
        print(f"Config\n{cfg}")
        return installers[1](model, **cfg)
    raise NotImplementedError(f"Mode {mode} is not implemented yet.")


def replace_minicpmv_forward(model, mode="fastv", **kwargs):
    """Meta interface of the MiniCPM-V wrapper (reference :111-135)."""
    return _wrapper_forward("replace_minicpmv_forward", (replace_minicpmv_fastv, replace_minicpmv_streamingllm), model, mode, kwargs)


def replace_nvila_forward(model, mode="merge_then_fastv_cost_given", **kwargs):
    """Meta interface of the NVILA wrapper (reference :138-164); like the reference's, its default mode raises."""
    model.mode = mode
    return _wrapper_forward("replace_nvila_forward", (replace_nvila_fastv, replace_nvila_streamingllm), model, mode, kwargs)


def replace_Qwen2_forward(model, mode="merge_then_fastv_cost_given", **kwargs):
    """Meta interface (reference :45-109): same modes, same keyword arguments and defaults."""
    print(f"replace_Qwen2_forward mode: {mode} and kwargs: {kwargs}")
    if mode == "prefill_merge":
        cfg = {"sparsity": kwargs.get("sparsity", [0.0] * 28)}
        print(f"Config\n{cfg}")
        cost, remaining_density = compute_density_overhead(cfg["sparsity"])
        print(f"Computational cost: {cost:.3f}, Remaining density: {remaining_density:.3f}")
        return replace_Qwen2_merging(model, **cfg)
    if mode == "fastv":
        cfg = {"fastv_k": kwargs.get("fastv_k", 3), "fastv_r": kwargs.get("fastv_r", 0.5)}
        print(f"Config\n{cfg}")
        return replace_Qwen2_fastv(model, **cfg)
    if mode == "merge_then_fastv":
        cfg = {"sparsity": kwargs.get("sparsity", [0.1] * 28), "fastv_k": kwargs.get("fastv_k", 3),
               "fastv_r": kwargs.get("fastv_r", 0.5)}
        print(f"Config\n{cfg}")
        return replace_Qwen2_merge_then_fastv(model, **cfg)
    if mode == "streamingllm":
        cfg = {"init_num": kwargs.get("init_num", 8), "length_rate": kwargs.get("length_rate", 0.3)}
        print(f"Config\n{cfg}")
        return replace_Qwen2_streamingllm(model, **cfg)
    if mode == "fastv_then_merge":
        cfg = {"fastv_k": kwargs.get("fastv_k", 2), "fastv_r": kwargs.get("fastv_r", 0.75),
               "merging_sparsity": kwargs.get("merging_sparsity", 0.3)}
        print(f"Config\n{cfg}")
        return replace_Qwen2_fastv_then_merge(model, **cfg)
    raise NotImplementedError(f"Mode {mode} is not implemented yet.")
