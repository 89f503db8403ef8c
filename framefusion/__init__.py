"""Drop-in import path: ``from framefusion.interface import apply_framefusion`` (reference README.md:123,
script/playground/example_llava.py:136) resolves to the B200-native implementation in ``framefusion_b200``.

Module names mirror the reference package (``framefusion.main``, ``.interface``, ``.utils``,
``.models.qwen2.modeling_qwen2``, ``.models.qwen2.modeling_qwen2_vl``, ``.models.qwenvl.modeling_qwen2_vl``); every
one of them only re-exports — the code lives in ``framefusion_b200``.  The unmodified reference used as the CPU baseline
is loaded by file path from ``baseline/_ref`` (``oracle/ref_locate.py``), never through this name.
"""
from framefusion_b200 import *  # noqa: F401,F403
from framefusion_b200 import __all__  # noqa: F401
