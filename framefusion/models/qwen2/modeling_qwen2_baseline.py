"""Comparison methods of the Qwen2 decoder (reference models/qwen2/modeling_qwen2_baseline.py) -> ``framefusion_b200.hooks.qwen2_baselines``."""
from framefusion_b200.baselines import compute_density_overhead  # noqa: F401
from framefusion_b200.hooks.qwen2_baselines import (  # noqa: F401
    Qwen2DecoderLayer_fastv_forward,
    Qwen2DecoderLayer_merge_then_fastv_forward,
    Qwen2DecoderLayer_merging_forward,
    Qwen2Model_fastv_forward,
    Qwen2Model_merge_then_fastv_forward,
    Qwen2Model_merging_forward,
    Qwen2SdpaAttention_fastv_forward,
    Qwen2SdpaAttention_merge_then_fastv_forward,
    Qwen2SdpaAttention_merging_forward,
    replace_Qwen2_fastv,
    replace_Qwen2_fastv_then_merge,
    replace_Qwen2_forward,
    replace_Qwen2_merge_then_fastv,
    replace_Qwen2_merging,
    replace_Qwen2_streamingllm,
)
