"""Hook trio of the Qwen2-VL text decoder (reference models/qwen2/modeling_qwen2_vl.py) -> ``framefusion_b200.hooks.qwen2_vl``."""
from framefusion_b200.hooks.qwen2_vl import (  # noqa: F401
    Qwen2VLDecoderLayer_merge_then_fastv_cost_given_forward,
    Qwen2VLModel_merge_then_fastv_cost_given_forward,
    Qwen2VLSdpaAttention_merge_then_fastv_cost_given_forward,
)
