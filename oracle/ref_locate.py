"""Where the UNMODIFIED reference lives, if anywhere.  TEST / BASELINE INFRASTRUCTURE (bench.py's CPU legs, tests).

Resolution order: ``$FF_REFERENCE_DIR`` -> ``/root/reference`` (the build container) ->
``baseline/_ref`` (the ``pip install --no-deps --target baseline/_ref`` copy that travels to the GPU box with a repo
snapshot; git-ignored, never committed).  A directory counts if it holds ``framefusion/main.py``.
"""
from __future__ import annotations

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def candidates():
    env = os.environ.get("FF_REFERENCE_DIR")
    if env:
        yield env
    yield "/root/reference"
    yield os.path.join(ROOT, "baseline", "_ref")


def find_reference_dir():
    for d in candidates():
        if os.path.isfile(os.path.join(d, "framefusion", "main.py")):
            return d
    return None


def load_reference_operator():
    """``(FrameFusion class, scaled_dot_product_attention, directory)`` of the unmodified reference, or ``None``.
    ``main.py`` is loaded by file path; ``utils.py`` imports matplotlib / torchvision for its debug dumps, so the one
    function the hot path uses is taken from its source text as is (``oracle/gen_golden.py``)."""
    d = find_reference_dir()
    if d is None:
        return None
    from oracle import gen_golden
    try:
        mod = gen_golden.load_reference(d)
        sdpa = gen_golden.load_reference_sdpa(d)
    except Exception:  # noqa: BLE001 - a broken copy is the same as none
        return None
    return mod.FrameFusion, sdpa, d
