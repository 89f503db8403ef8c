"""Hook trio of the Qwen2 decoder (reference models/qwen2/modeling_qwen2.py:11-333) -> ``framefusion_b200.hooks.qwen2``."""
from framefusion_b200.hooks.qwen2 import (  # noqa: F401
    Qwen2DecoderLayer_merge_then_prune_by_cost_forward,
    Qwen2Model_merge_then_fastv_cost_given_forward,
    Qwen2SdpaAttention_merge_then_prune_by_cost_forward,
)
