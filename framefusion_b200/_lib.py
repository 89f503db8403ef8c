"""ctypes binding of the C ABI declared in ``include/framefusion_b200.h``.

This is the only way the Python host reaches the kernels: raw device pointers (``tensor.data_ptr()``), sizes
and the current CUDA stream handle.  The library must exist — there is no CPU or eager-torch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_i64 = C.c_int64
_vp = C.c_void_p

ABI_VERSION = 6            # FF_ABI_VERSION of include/framefusion_b200.h
FF_BF16, FF_F16, FF_F32 = 0, 1, 2
FF_MAX_AUX = 6

# status slots (enum ff_status_slot)
ST_SEQ_KEEP, ST_COUNT, ST_NVIS, ST_NCHAIN, ST_BRANCH, ST_TOPK, ST_ERROR, ST_NMERGED, ST_FUSED, ST_INTERNAL, ST_SEQ = range(11)
ST_SLOTS = 16

EXPORTS = [
    "ff_abi_version", "ff_last_error", "ff_launch_count", "ff_ctx_create", "ff_ctx_destroy", "ff_ctx_status", "ff_ctx_timing", "ff_stream_sync", "ff_status_wait", "ff_workspace_bytes",
    "ff_build_links", "ff_build_links_for", "ff_similarity", "ff_merge_apply", "ff_merge_layer", "ff_importance", "ff_prune_layer",
    "ff_compact_mask", "ff_debug_read", "ff_debug_frame_trace",
]


class FFAux(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("planes", _i64), ("src_plane_stride", _i64),
                ("dst_plane_stride", _i64), ("row_bytes", _i64)]


class FFError(RuntimeError):
    pass


_lib = None


def load():
    """Load (building first if the sources are newer) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FF_LIB_PATH") or _build.LIB_PATH     # FF_LIB_PATH: a prebuilt library elsewhere (no rebuild)
    if path == _build.LIB_PATH and _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on a runtime box: use the shipped binary if there is one
            if not os.path.exists(path):
                raise FFError(f"libframefusion_b200.so is missing and could not be built: {e}") from e
    lib = C.CDLL(path)
    lib.ff_abi_version.restype = C.c_int
    lib.ff_last_error.restype = C.c_char_p
    lib.ff_launch_count.restype = _i64
    lib.ff_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
    lib.ff_ctx_destroy.argtypes = [_vp]
    lib.ff_ctx_status.argtypes = [_vp]
    lib.ff_ctx_status.restype = C.POINTER(_i64)
    lib.ff_stream_sync.argtypes = [_vp, _vp]
    lib.ff_status_wait.argtypes = [_vp, _vp]
    lib.ff_ctx_timing.argtypes = [_vp, _vp, _vp]
    lib.ff_workspace_bytes.argtypes = [_i64, _i64]
    lib.ff_workspace_bytes.restype = _i64
    lib.ff_build_links.argtypes = [_vp, _vp, _i64, _vp, _i64, _i64, _vp]
    lib.ff_build_links_for.argtypes = [_vp, _vp, _i64, _vp, _i64, _i64, _i64, C.c_int, _vp]
    lib.ff_similarity.argtypes = [_vp, _vp, _i64, _vp, C.c_int, _i64, _i64, C.c_double, _vp, _vp, _vp]
    lib.ff_merge_apply.argtypes = [_vp, _vp, _i64, _vp, C.c_int, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _vp]
    lib.ff_merge_layer.argtypes = [_vp, _vp, _i64, _vp, _vp, C.c_int, _i64, _i64, C.c_double, C.c_double,
                                   C.POINTER(FFAux), C.c_int, C.c_int, _vp]
    lib.ff_importance.argtypes = [_vp, _vp, _vp, C.c_int, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                  C.c_int, C.c_double, _vp, _vp, _i64, _vp]
    lib.ff_prune_layer.argtypes = [_vp, _vp, _i64, _vp, _i64, _vp, _vp, C.c_int, _i64, _i64, _i64, _i64, _i64,
                                   C.POINTER(FFAux), C.c_int, _vp, _vp]
    lib.ff_compact_mask.argtypes = [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp]
    lib.ff_debug_read.argtypes = [_vp, _vp, _i64, C.c_int, _vp, _i64, C.c_int, _vp]
    lib.ff_debug_frame_trace.argtypes = [_vp, _vp, _i64]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("ff_last_error", "ff_launch_count", "ff_ctx_status", "ff_workspace_bytes"):
            fn.restype = C.c_int
    if lib.ff_abi_version() != ABI_VERSION:
        raise FFError(f"ABI version {lib.ff_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().ff_last_error().decode()
        if rc == -1:
            raise ValueError(f"framefusion_b200: {msg}")
        raise FFError(f"framefusion_b200 (code {rc}): {msg}")
