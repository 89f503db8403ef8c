"""Hook trio for the Qwen2-VL text decoder (``/root/reference/framefusion/models/qwen2/modeling_qwen2_vl.py``),
re-targeted to the transformers 5.x layout (``Qwen2VLForConditionalGeneration.model.language_model`` is the
``Qwen2VLTextModel``).  Differences from the plain Qwen2 trio, all from the reference:

* rotary embeddings are M-RoPE: cos / sin are ``[3, B, S, D]``, so FrameFusion compacts them along dim 2
  (main.py:145-147, 165-167) — the decoder hook is shared with Qwen2, the operator handles the rank;
* importance uses the last FOUR queries (``num=4``, reference :289-301).

The top-level ``forward`` patch of the reference only builds ``patch_type`` and calls ``prepare``
(models/qwenvl/modeling_qwen2_vl.py:117-138).  The reference does it by pasting a copy of the whole
``Qwen2VLForConditionalGeneration.forward`` around those lines; here ``forward`` below does the same layout step and
then hands every argument to the model class's own ``forward`` — nothing of transformers is duplicated, so the patch
follows the installed version.
"""
from __future__ import annotations

from typing import Optional

import torch
from transformers.cache_utils import Cache, DynamicCache
from transformers.masking_utils import create_causal_mask, create_sliding_window_causal_mask
from transformers.modeling_outputs import BaseModelOutputWithPast
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
from transformers.models.qwen2_vl.modeling_qwen2_vl import apply_multimodal_rotary_pos_emb, eager_attention_forward

from ..utils import scaled_dot_product_attention
from .qwen2 import Qwen2DecoderLayer_merge_then_prune_by_cost_forward as Qwen2VLDecoderLayer_merge_then_fastv_cost_given_forward


def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
            labels=None, use_cache=None, pixel_values=None, pixel_values_videos=None, image_grid_thw=None,
            video_grid_thw=None, **kwargs):
    """Embed-stage patch of ``Qwen2VLForConditionalGeneration`` (reference models/qwenvl/modeling_qwen2_vl.py:11-138,
    installed by ``apply_framefusion`` / ``get_token_type``): on a prefill with one video, derive the token layout from
    ``input_ids`` and ``video_grid_thw`` — ``patch_num = H * W / merge^2`` tokens per frame between the first and the last
    video placeholder (reference :118-127) —, remember it on the model like the reference (:129-133) and call
    ``self.framefusion.prepare`` (:136-137, skipped when the model carries a ``mode`` attribute); then run the model's own
    forward.  Decode steps (one token) and calls without a video pass straight through."""
    from .. import layout
    seq = input_ids.shape[1] if input_ids is not None else (inputs_embeds.shape[1] if inputs_embeds is not None else 1)
    if seq != 1 and video_grid_thw is not None and input_ids is not None:
        vision_cfg = getattr(self.config, "vision_config", None)
        merge = getattr(vision_cfg, "spatial_merge_size", None)
        if merge is None:                                    # transformers 4.x keeps it on the tower
            merge = self.visual.config.spatial_merge_size
        device = self.get_input_embeddings().weight.device
        args = layout.qwen2vl_prepare_args(input_ids, self.config.video_token_id, video_grid_thw, merge, device=device)
        (_patch_type, self.patch_num, self.image_token_start_index, self.image_token_end_index,
         self.image_token_length, self.original_length) = args
        if not hasattr(self, "mode"):
            self.framefusion.prepare(*args)
    return type(self).forward(self, input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                              past_key_values=past_key_values, inputs_embeds=inputs_embeds, labels=labels,
                              use_cache=use_cache, pixel_values=pixel_values, pixel_values_videos=pixel_values_videos,
                              image_grid_thw=image_grid_thw, video_grid_thw=video_grid_thw, **kwargs)


def llm_key(model) -> str:
    return "model.language_model" if hasattr(model.model, "language_model") else "model"


def trio():
    return (Qwen2VLModel_merge_then_fastv_cost_given_forward, Qwen2VLDecoderLayer_merge_then_fastv_cost_given_forward,
            Qwen2VLSdpaAttention_merge_then_fastv_cost_given_forward)


def Qwen2VLSdpaAttention_merge_then_fastv_cost_given_forward(
    self,
    hidden_states: torch.Tensor,
    attention_mask: Optional[torch.Tensor] = None,
    position_ids: Optional[torch.LongTensor] = None,
    past_key_values: Optional[Cache] = None,
    output_attentions: bool = False,
    use_cache: bool = False,
    position_embeddings=None,
    **kwargs,
):
    """-> ``(attn_output, importance_or_None)``; importance is ``[1, heads, 4, S]`` (reference :289-301)."""
    kwargs.pop("cache_position", None)
    bsz, q_len, _ = hidden_states.size()
    query_states = self.q_proj(hidden_states).view(bsz, q_len, -1, self.head_dim).transpose(1, 2)
    key_states = self.k_proj(hidden_states).view(bsz, q_len, -1, self.head_dim).transpose(1, 2)
    value_states = self.v_proj(hidden_states).view(bsz, q_len, -1, self.head_dim).transpose(1, 2)

    cos, sin = position_embeddings
    query_states, key_states = apply_multimodal_rotary_pos_emb(
        query_states, key_states, cos, sin, self.config.rope_parameters["mrope_section"])

    if past_key_values is not None:
        key_states, value_states = past_key_values.update(key_states, value_states, self.layer_idx)

    if attention_mask is not None and attention_mask.ndim == 4:
        # a decode step after a reduced prefill: the 4-D mask was built for layer 0's cache, later layers hold fewer
        # keys (reference models/qwen2/modeling_qwen2.py:150-152 slices the same way)
        attention_mask = attention_mask[..., : key_states.shape[-2]]

    is_causal = attention_mask is None and q_len > 1
    attn_weights = None
    ff = self.framefusion
    if q_len > 1 and ff.finish_merging and not ff.finish_pruning:
        attn_weights = scaled_dot_product_attention(
            query_states, key_states, value_states, num=4, attn_mask=None,
            dropout_p=self.attention_dropout if self.training else 0.0,
            is_causal=is_causal, scale=self.scaling, enable_gqa=True)

    attention_interface = ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation, eager_attention_forward)
    attn_output, _ = attention_interface(
        self, query_states, key_states, value_states, attention_mask,
        dropout=0.0 if not self.training else self.attention_dropout,
        scaling=self.scaling, sliding_window=self.sliding_window, position_ids=position_ids, **kwargs)
    attn_output = attn_output.reshape(bsz, q_len, -1).contiguous()
    attn_output = self.o_proj(attn_output)
    return attn_output, attn_weights


def Qwen2VLModel_merge_then_fastv_cost_given_forward(
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
    use_cache = use_cache if use_cache is not None else self.config.use_cache
    if (input_ids is None) ^ (inputs_embeds is not None):
        raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
    if use_cache and past_key_values is None:
        past_key_values = DynamicCache(config=self.config)
    if inputs_embeds is None:
        inputs_embeds = self.embed_tokens(input_ids)

    q_len = inputs_embeds.shape[1]
    past_seen_tokens = past_key_values.get_seq_length() if past_key_values is not None else 0
    # the hard coded 3 is for temporal, height and width
    if position_ids is None:
        position_ids = torch.arange(q_len, device=inputs_embeds.device) + past_seen_tokens
        position_ids = position_ids.view(1, 1, -1).expand(3, inputs_embeds.shape[0], -1)
    elif position_ids.ndim == 2:
        position_ids = position_ids[None, ...].expand(3, position_ids.shape[0], -1)
    if position_ids.ndim == 3 and position_ids.shape[0] == 4:
        text_position_ids = position_ids[0]
        position_ids = position_ids[1:]
    else:
        text_position_ids = None

    if attention_mask is not None and not isinstance(attention_mask, dict) and attention_mask.ndim == 2 \
            and attention_mask.shape[-1] != past_seen_tokens + q_len:
        attention_mask = None          # batch 1, no padding: see hooks/qwen2.py

    if not isinstance(causal_mask_mapping := attention_mask, dict):
        mask_kwargs = dict(config=self.config, inputs_embeds=inputs_embeds, attention_mask=attention_mask,
                           past_key_values=past_key_values, position_ids=text_position_ids)
        causal_mask_mapping = {"full_attention": create_causal_mask(**mask_kwargs)}
        if getattr(self, "has_sliding_layers", False):
            causal_mask_mapping["sliding_attention"] = create_sliding_window_causal_mask(**mask_kwargs)

    hidden_states = inputs_embeds
    # a list of two [3, B, S, D] tensors, so that FrameFusion can replace its entries
    position_embeddings = list(self.rotary_emb(hidden_states, position_ids))

    all_hidden_states = () if output_hidden_states else None
    masks = dict(causal_mask_mapping)
    for i, decoder_layer in enumerate(self.layers):
        if output_hidden_states:
            all_hidden_states += (hidden_states,)
        kind = self.config.layer_types[i]
        layer_outputs = decoder_layer(
            hidden_states,
            attention_mask=masks[kind],
            position_embeddings=position_embeddings,
            position_ids=text_position_ids,
            past_key_values=past_key_values,
            use_cache=use_cache,
            **kwargs,
        )
        hidden_states = layer_outputs[0]
        position_embeddings = layer_outputs[-2]
        masks[kind] = layer_outputs[-1]

    hidden_states = self.norm(hidden_states)
    if output_hidden_states:
        all_hidden_states += (hidden_states,)
    out = BaseModelOutputWithPast(last_hidden_state=hidden_states,
                                  past_key_values=past_key_values if use_cache else None,
                                  hidden_states=all_hidden_states)
    if return_dict is False:
        return out.to_tuple()
    return out
