"""Hook trio for a plain Qwen2 decoder stack (LLaVA-Video / MiniCPM-V / NVILA language models).

Same roles and names as ``/root/reference/framefusion/models/qwen2/modeling_qwen2.py`` (decoder hook :11-86,
attention hook :89-195, model hook :198-333), re-targeted from transformers 4.45.2 to the 5.x module API this
image ships (``Qwen2DecoderLayer.forward(hidden_states, attention_mask, position_ids, past_key_values,
use_cache, position_embeddings, **kw)``, ``Qwen2Attention.forward(hidden_states, position_embeddings,
attention_mask, past_key_values, **kw) -> (out, weights)``).  What the reference adds to the stock forwards,
and what is kept here:

* the decoder layer calls ``self.framefusion`` before attention at layer 0 and after attention (before the MLP)
  at every layer, and returns the possibly-compacted ``position_embeddings`` and ``attention_mask`` as the LAST
  TWO elements of its output tuple (reference :44-47, :66-68, :84-86);
* the attention computes the last-query attention probabilities only while pruning is armed
  (``finish_merging and not finish_pruning``, reference :166-178) and returns them as its second output;
* the model keeps ``position_embeddings`` in a *list* so the operator can swap its entries, and threads the two
  returned values into the next layer (reference :262-266, :303-306).

The importance signal goes through ``framefusion_b200.utils.scaled_dot_product_attention`` with K *before*
``repeat_kv`` (the kernel is GQA aware), everything else is the stock module code path.
"""
from __future__ import annotations

from typing import Optional

import torch
from transformers.cache_utils import Cache, DynamicCache
from transformers.masking_utils import create_causal_mask, create_sliding_window_causal_mask
from transformers.modeling_outputs import BaseModelOutputWithPast
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
from transformers.models.qwen2.modeling_qwen2 import apply_rotary_pos_emb, eager_attention_forward

from ..utils import scaled_dot_product_attention


def Qwen2DecoderLayer_merge_then_prune_by_cost_forward(
    self,
    hidden_states: torch.Tensor,
    attention_mask: Optional[torch.Tensor] = None,
    position_ids: Optional[torch.LongTensor] = None,
    past_key_values: Optional[Cache] = None,
    use_cache: Optional[bool] = False,
    position_embeddings=None,
    **kwargs,
):
    """-> ``(hidden_states, position_embeddings, attention_mask)``; the last two as updated by FrameFusion."""
    if "past_key_value" in kwargs:                       # 4.45 spelling used by the reference's callers
        past_key_values = kwargs.pop("past_key_value")
    for stale in ("output_attentions", "cache_position"):
        kwargs.pop(stale, None)

    # token merging at layer 0 before attention (reference :44-47)
    if self.self_attn.layer_idx == 0:
        hidden_states, position_embeddings, attention_mask = self.framefusion(hidden_states, position_embeddings, attention_mask)

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

    # token merging or importance pruning after attention (reference :66-68)
    hidden_states, position_embeddings, attention_mask = self.framefusion(
        hidden_states, position_embeddings, attention_mask, self_attn_weights)

    residual = hidden_states
    hidden_states = self.post_attention_layernorm(hidden_states)
    hidden_states = self.mlp(hidden_states)
    hidden_states = residual + hidden_states
    return (hidden_states, position_embeddings, attention_mask)


def Qwen2SdpaAttention_merge_then_prune_by_cost_forward(
    self,
    hidden_states: torch.Tensor,
    position_embeddings=None,
    attention_mask: Optional[torch.Tensor] = None,
    past_key_values: Optional[Cache] = None,
    **kwargs,
):
    """-> ``(attn_output, importance_or_None)``.  ``importance`` is ``[1, heads, 1, S]`` (reference :166-178)."""
    ff = self.framefusion
    want = hidden_states.shape[1] > 1 and ff.finish_merging and not ff.finish_pruning
    return attention_with_importance(self, hidden_states, position_embeddings, attention_mask, past_key_values, want, **kwargs)


def attention_with_importance(self, hidden_states, position_embeddings, attention_mask, past_key_values, want_importance,
                              **kwargs):
    """The stock attention forward plus, when ``want_importance``, the last query's attention probabilities
    ``[1, heads, 1, S]`` as second output (shared by the FrameFusion hook above and the FastV baseline hook,
    ``hooks/qwen2_baselines.py``)."""
    for stale in ("position_ids", "use_cache", "output_attentions", "cache_position"):
        kwargs.pop(stale, None)
    input_shape = hidden_states.shape[:-1]
    q_len = input_shape[1]
    hidden_shape = (*input_shape, -1, self.head_dim)

    query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
    key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
    value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)

    cos, sin = position_embeddings
    query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)

    if past_key_values is not None:
        # the cache keeps this layer's keys at the length the layer SAW (the reduction that follows the attention
        # does not touch it, reference :143-145): per-layer ragged lengths are fine for DynamicCache
        key_states, value_states = past_key_values.update(key_states, value_states, self.layer_idx)

    if attention_mask is not None and attention_mask.ndim == 4:
        # a decode step after a reduced prefill: the 4-D mask was built for layer 0's cache, later layers hold fewer
        # keys (reference models/qwen2/modeling_qwen2.py:150-152 slices the same way)
        attention_mask = attention_mask[..., : key_states.shape[-2]]

    is_causal = attention_mask is None and q_len > 1

    attn_weights = None
    if want_importance:
        attn_weights = scaled_dot_product_attention(
            query_states, key_states, value_states, num=1, attn_mask=None,
            dropout_p=self.attention_dropout if self.training else 0.0,
            is_causal=is_causal, scale=self.scaling, enable_gqa=True)

    attention_interface = ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation, eager_attention_forward)
    attn_output, _ = attention_interface(
        self, query_states, key_states, value_states, attention_mask,
        dropout=0.0 if not self.training else self.attention_dropout,
        scaling=self.scaling, sliding_window=self.sliding_window, **kwargs)
    attn_output = attn_output.reshape(*input_shape, -1).contiguous()
    attn_output = self.o_proj(attn_output)
    return attn_output, attn_weights


def model_inputs(self, input_ids, attention_mask, position_ids, past_key_values, inputs_embeds, use_cache):
    """What every patched model forward does before its layer loop: embeddings, cache, positions, the causal masks per
    layer kind and the position embeddings as a LIST (so that an operator can swap its entries, reference :262-266).
    -> ``(hidden_states, position_ids, past_key_values, masks, position_embeddings, use_cache)``."""
    use_cache = use_cache if use_cache is not None else self.config.use_cache
    if (input_ids is None) ^ (inputs_embeds is not None):
        raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
    if inputs_embeds is None:
        inputs_embeds = self.embed_tokens(input_ids)
    if use_cache and past_key_values is None:
        past_key_values = DynamicCache(config=self.config)

    q_len = inputs_embeds.shape[1]
    # layer 0 holds the shortest cache (it is filled after the pre-attention merge): positions of a decode step
    # continue from there, exactly as the reference derives cache_position (reference :248-254)
    past_seen_tokens = past_key_values.get_seq_length() if past_key_values is not None else 0
    if position_ids is None:
        position_ids = (torch.arange(q_len, device=inputs_embeds.device) + past_seen_tokens).unsqueeze(0)

    if attention_mask is not None and not isinstance(attention_mask, dict) and attention_mask.ndim == 2 \
            and attention_mask.shape[-1] != past_seen_tokens + q_len:
        # generate() keeps extending the 2-D padding mask at the ORIGINAL length; after a reduction the caches are
        # shorter.  Batch size is 1 (no padding), so the mask carries no information: drop it.
        attention_mask = None

    if not isinstance(causal_mask_mapping := attention_mask, dict):
        mask_kwargs = dict(config=self.config, inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                           past_key_values=past_key_values, position_ids=position_ids)
        causal_mask_mapping = {"full_attention": create_causal_mask(**mask_kwargs)}
        if getattr(self, "has_sliding_layers", False):
            causal_mask_mapping["sliding_attention"] = create_sliding_window_causal_mask(**mask_kwargs)

    hidden_states = inputs_embeds
    # a list, so that FrameFusion can replace its entries (reference :262-266)
    position_embeddings = list(self.rotary_emb(hidden_states, position_ids))
    return hidden_states, position_ids, past_key_values, dict(causal_mask_mapping), position_embeddings, use_cache


def model_outputs(self, hidden_states, past_key_values, use_cache, all_hidden_states, return_dict):
    hidden_states = self.norm(hidden_states)
    if all_hidden_states is not None:
        all_hidden_states += (hidden_states,)
    out = BaseModelOutputWithPast(
        last_hidden_state=hidden_states,
        past_key_values=past_key_values if use_cache else None,
        hidden_states=all_hidden_states,
    )
    if return_dict is False:
        return out.to_tuple()
    return out


def Qwen2Model_merge_then_fastv_cost_given_forward(
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
    hidden_states, position_ids, past_key_values, masks, position_embeddings, use_cache = model_inputs(
        self, input_ids, attention_mask, position_ids, past_key_values, inputs_embeds, use_cache)

    all_hidden_states = () if output_hidden_states else None
    for i, decoder_layer in enumerate(self.layers[: self.config.num_hidden_layers]):
        if output_hidden_states:
            all_hidden_states += (hidden_states,)
        kind = self.config.layer_types[i]
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
        # the position embeddings and mask as FrameFusion left them (reference :303-306)
        position_embeddings = layer_outputs[-2]
        masks[kind] = layer_outputs[-1]

    return model_outputs(self, hidden_states, past_key_values, use_cache, all_hidden_states, return_dict)
