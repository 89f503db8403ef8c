"""``apply_framefusion`` and the forward patcher — mirror of ``/root/reference/framefusion/interface.py``.

``apply_framefusion(model, cost, similarity_lower_bound, ratio_lower_bound)`` (reference :47-137) picks the hook
trio for the model family and calls ``replace_framefusion_forward`` (reference :169-214), which creates ONE
``FrameFusion`` operator and hangs it on the model, the language model, every decoder layer and every attention
module, rebinding their ``forward`` with ``MethodType``.  The three callables stay user-replaceable — the
reference documents them as its extension point (README.md:171-175).

Families handled here are the ones on the north-star path: a plain Qwen2 decoder stack (LLaVA-Video's
``LlavaQwenForCausalLM``, ``Qwen2ForCausalLM``; keys ``model / layers / self_attn``, reference :69-77), the
wrappers that hold one under ``llm.model`` (MiniCPM-V, NVILA; reference :80-99) and Qwen2-VL (reference :101-109),
whose embed-stage patch — the top-level ``forward`` that derives the token layout and calls ``prepare`` — is
installed as well (``hooks/qwen2_vl.forward``).  The embed patches of the other families rewrite functions of
un-vendored third-party packages (``llava``, remote-code models) that are not in this image: for those, callers hand
the token layout to ``model.framefusion.prepare`` themselves, or use ``framefusion_b200.layout`` to build it.
Anything else raises ``NotImplementedError`` after printing the model, like the reference (:120-124).
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


def _embed_patch(model):
    """The embed-stage patch of the model's family (what the reference's ``get_token_type`` installs, :140-166), or None
    when it belongs to third-party code that is not in this image."""
    names = {c.__name__ for c in type(model).__mro__}
    if "Qwen2VLForConditionalGeneration" in names:
        from .hooks.qwen2_vl import forward
        return "forward", forward
    return None


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
    """Patch ``model`` in place so that its prefill runs FrameFusion (reference interface.py:47-137).

    ``cost`` is the compute budget relative to the dense model, ``similarity_lower_bound`` the cosine threshold
    above which adjacent-frame tokens merge, ``ratio_lower_bound`` the merged fraction below which merging stops
    and pruning takes over.  Unsupported model types print the model and raise ``NotImplementedError``."""
    family = _family(model)
    if family is None:
        print(f"Model not supported")
        print(f"Model type: {type(model)}")
        print(model)
        raise NotImplementedError
    llm_key, trio = family
    patch = _embed_patch(model)
    if patch is not None:
        setattr(model, patch[0], MethodType(patch[1], model))
    replace_framefusion_forward(model, cost, similarity_lower_bound, ratio_lower_bound, *trio,
                                llm_key=llm_key, decoder_key="layers", attention_key="self_attn")


def get_token_type(model):
    """Installs ONLY the embed-stage patch of the model's family (reference :140-166) — the function that derives the
    token layout (``patch_type``, ``patch_num``, the vision span) during the prefill and leaves it on the model.  Qwen2-VL:
    the top-level ``forward`` (reference :157-158).  The patches of the other families rewrite third-party functions that
    are not in this image (``llava``'s ``prepare_inputs_labels_for_multimodal``, NVILA's ``_embed``, MiniCPM-V's
    ``get_vllm_embedding``): for them this raises ``NotImplementedError`` naming ``framefusion_b200.layout``, whose
    builders produce the same arguments.  Unknown models raise ``NotImplementedError`` like the reference (:165-166)."""
    if _family(model) is None:
        raise NotImplementedError
    patch = _embed_patch(model)
    if patch is None:
        raise NotImplementedError(
            f"the embed-stage patch of {type(model).__name__} rewrites third-party code that is not installed here; "
            "build the layout with framefusion_b200.layout and call model.framefusion.prepare(...)")
    setattr(model, patch[0], MethodType(patch[1], model))


def _rebind(target: nn.Module, fn: Callable, operator: FrameFusion):
    target.framefusion = operator
    target.forward = MethodType(fn, target)


def _keep_accelerate_hook(layer: nn.Module, fn: Callable):
    """accelerate's device-alignment hook wraps ``forward``; after the rebinding it has to wrap the new one
    (reference :204-207).  A no-op when the layer was not dispatched or accelerate is not installed."""
    hook = getattr(layer, "_hf_hook", None)
    if hook is None:
        return
    try:
        from accelerate.hooks import add_hook_to_module
    except ModuleNotFoundError:
        return
    layer._old_forward = MethodType(fn, layer)
    add_hook_to_module(layer, hook)


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
    """Creates ONE ``FrameFusion`` operator, hangs it on ``module``, on the language model found under ``llm_key``,
    on every decoder layer under ``decoder_key`` and on every attention module under ``attention_key`` (dotted
    paths, e.g. ``"llm.model"``), and rebinds their ``forward`` to the three callables (reference :169-214)."""
    operator = FrameFusion(cost, similarity_lower_bound, ratio_lower_bound)
    module.framefusion = operator

    llm = get_attr_by_name(module, llm_key)
    assert isinstance(llm, PreTrainedModel), f"{llm_key} is not a PreTrainedModel"
    _rebind(llm, llm_forward, operator)

    for index, layer in enumerate(get_attr_by_name(llm, decoder_key)):
        assert isinstance(layer, nn.Module), f"{decoder_key}[{index}] is not a nn.Module"
        _rebind(layer, decoder_forward, operator)
        _keep_accelerate_hook(layer, decoder_forward)
        attention = get_attr_by_name(layer, attention_key)
        assert isinstance(attention, nn.Module), f"{decoder_key}[{index}].{attention_key} is not a nn.Module"
        _rebind(attention, attention_forward, operator)
