"""``apply_framefusion`` and the forward patcher — mirror of ``/root/reference/framefusion/interface.py``.

``apply_framefusion(model, cost, similarity_lower_bound, ratio_lower_bound)`` (reference :47-137) picks the hook
trio for the model family and calls ``replace_framefusion_forward`` (reference :169-214), which creates ONE
``FrameFusion`` operator and hangs it on the model, the language model, every decoder layer and every attention
module, rebinding their ``forward`` with ``MethodType``.  The three callables stay user-replaceable — the
reference documents them as its extension point (README.md:171-175).

Families handled here are the ones on the north-star path: a plain Qwen2 decoder stack (LLaVA-Video's
``LlavaQwenForCausalLM``, ``Qwen2ForCausalLM``; keys ``model / layers / self_attn``, reference :69-77) and the
wrappers that hold one under ``llm.model`` (MiniCPM-V, NVILA; reference :80-99).  The vision-side embed patches of
those families live in un-vendored third-party packages (``llava``, remote-code models) and are out of scope:
callers hand the token layout to ``model.framefusion.prepare`` themselves, or use ``framefusion_b200.layout`` to
build it.  Anything else raises ``NotImplementedError`` after printing the model, like the reference (:120-124).
"""
from __future__ import annotations

from types import MethodType
from typing import Callable

import torch
import torch.nn as nn
from transformers import PreTrainedModel

from .main import FrameFusion
from .utils import TEXT_TOKEN, IGNORE_TOKEN, get_attr_by_name  # noqa: F401  (re-exported like the reference)


def _qwen2_trio():
    from .hooks.qwen2 import (Qwen2Model_merge_then_fastv_cost_given_forward,
                              Qwen2DecoderLayer_merge_then_prune_by_cost_forward,
                              Qwen2SdpaAttention_merge_then_prune_by_cost_forward)
    return (Qwen2Model_merge_then_fastv_cost_given_forward, Qwen2DecoderLayer_merge_then_prune_by_cost_forward,
            Qwen2SdpaAttention_merge_then_prune_by_cost_forward)


def _family(model):
    """-> (llm_key, trio) or None.  Dispatch order follows the reference (:58-124)."""
    names = {c.__name__ for c in type(model).__mro__}
    arch = (getattr(getattr(model, "config", None), "architectures", None) or [None])[0]
    if "LlavaQwenForCausalLM" in names or "Qwen2ForCausalLM" in names:
        return "model", _qwen2_trio()
    if arch == "MiniCPMV" or "LlavaLlamaModel" in names:
        return "llm.model", _qwen2_trio()
    if "Qwen2VLForConditionalGeneration" in names:
        from .hooks.qwen2_vl import trio, llm_key
        return llm_key(model), trio()
    return None


def apply_framefusion(model, cost, similarity_lower_bound, ratio_lower_bound):
    """
    Apply FrameFusion to the model

    Args:
        model: the model to apply FrameFusion to
        cost: the cost of the FrameFusion
        similarity_lower_bound: the similarity lower bound of the FrameFusion
        ratio_lower_bound: the ratio lower bound of the FrameFusion
    """
    fam = _family(model)
    if fam is None:
        print(f"Model not supported")
        print(f"Model type: {type(model)}")
        print(model)
        raise NotImplementedError
    llm_key, (llm_forward, decoder_forward, attention_forward) = fam
    replace_framefusion_forward(
        model,
        cost=cost,
        similarity_lower_bound=similarity_lower_bound,
        ratio_lower_bound=ratio_lower_bound,
        llm_forward=llm_forward,
        decoder_forward=decoder_forward,
        attention_forward=attention_forward,
        llm_key=llm_key,
        decoder_key="layers",
        attention_key="self_attn",
    )


def get_token_type(model):
    """The reference installs only the embed-stage patch here (:140-166).  Those patches belong to third-party
    model code that is not in this image; the token layout builders are in ``framefusion_b200.layout``."""
    if _family(model) is None:
        raise NotImplementedError
    return None


def replace_framefusion_forward(
    module: torch.nn.Module,
    cost: float,
    similarity_lower_bound: float,
    ratio_lower_bound: float,
    llm_forward: Callable,
    decoder_forward: Callable,
    attention_forward: Callable,
    llm_key: str = "model",
    decoder_key: str = "layers",
    attention_key: str = "self_attn",
):
    """
    Replace the forward method of the model with the framefusion forward method.
    Make framefusion a property of the model.

    The keys are accessed in an hierarchical manner: llm_key -> decoder_key -> attention_key. Each key can have
    multiple hierarchies, e.g. "llm.model", which will be accessed by module.llm.model
    """
    framefusion = FrameFusion(cost, similarity_lower_bound, ratio_lower_bound)
    module.framefusion = framefusion

    llm = get_attr_by_name(module, llm_key)
    assert isinstance(llm, PreTrainedModel), f"{llm_key} is not a PreTrainedModel"
    llm.framefusion = framefusion
    llm.forward = MethodType(llm_forward, llm)

    decoder_layers = get_attr_by_name(llm, decoder_key)
    for i, decoder_layer in enumerate(decoder_layers):
        assert isinstance(decoder_layer, nn.Module), f"{decoder_key}[{i}] is not a nn.Module"
        decoder_layer.framefusion = framefusion
        decoder_layer.forward = MethodType(decoder_forward, decoder_layer)

        # keep accelerate's device-alignment hook alive across the rebinding (reference :204-207)
        if hasattr(decoder_layer, "_hf_hook"):
            try:
                from accelerate.hooks import add_hook_to_module
            except ModuleNotFoundError:
                add_hook_to_module = None
            if add_hook_to_module is not None:
                decoder_layer._old_forward = MethodType(decoder_forward, decoder_layer)
                add_hook_to_module(decoder_layer, decoder_layer._hf_hook)

        attention = get_attr_by_name(decoder_layer, attention_key)
        assert isinstance(attention, nn.Module), f"{decoder_key}[{i}].{attention_key} is not a nn.Module"
        attention.framefusion = framefusion
        attention.forward = MethodType(attention_forward, attention)
