"""Embed-stage patch of Qwen2-VL (reference models/qwenvl/modeling_qwen2_vl.py:11-138) -> ``framefusion_b200.hooks.qwen2_vl.forward``."""
from framefusion_b200.hooks.qwen2_vl import forward  # noqa: F401
from framefusion_b200.utils import TEXT_TOKEN  # noqa: F401
