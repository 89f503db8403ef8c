"""framefusion_b200 — B200-native FrameFusion token-reduction path.

Public surface mirrors ``/root/reference/framefusion``: ``main.FrameFusion``, ``interface.apply_framefusion``,
``utils.scaled_dot_product_attention``.  Submodules are imported lazily so that ``framefusion_b200.synth``
(used by the test fixtures) does not pull the CUDA binding in.
"""
__all__ = ["FrameFusion", "apply_framefusion", "replace_framefusion_forward", "get_token_type"]


def __getattr__(name):
    if name == "FrameFusion":
        from .main import FrameFusion
        return FrameFusion
    if name in ("apply_framefusion", "replace_framefusion_forward", "get_token_type"):
        from . import interface
        return getattr(interface, name)
    raise AttributeError(name)
