"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/framefusion_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "framefusion_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ff_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from framefusion_b200 import _build, _lib
    _build.build()
    lib = ctypes.CDLL(_build.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(_lib.EXPORTS) <= set(names)
    l = _lib.load()
    assert l.ff_abi_version() == _lib.ABI_VERSION == 6
    assert l.ff_workspace_bytes(36898, 576) > 36898 * 4 * 8
    assert l.ff_workspace_bytes(-1, 0) == -1


def test_bad_arguments_are_rejected_without_a_gpu():
    from framefusion_b200 import _lib
    l = _lib.load()
    rc = l.ff_build_links(None, None, 0, None, 10, 4, None)
    assert rc == -1 and b"null" in l.ff_last_error()
    rc = l.ff_merge_layer(None, None, 0, None, None, 0, 10, 16, 0.6, 0.7, None, 0, 0, None)
    assert rc == -1


def test_host_budget_formula_matches_oracle():
    from framefusion_b200.main import FrameFusion
    from oracle import ff_oracle as orc
    import pytest
    for cost in (0.2, 0.3, 0.5, 0.9):
        for sl in ([], [0.39], [0.39, 0.2], [0.5, 0.4, 0.3, 0.05]):
            try:
                want = orc.compute_pruning_ratio(sl, cost)
            except ValueError:
                with pytest.raises(ValueError, match="The cost is too small"):
                    FrameFusion._compute_pruning_ratio(sl, cost)
                continue
            assert FrameFusion._compute_pruning_ratio(sl, cost) == want


def test_cpu_tensors_fail_loudly():
    import pytest
    import torch
    from framefusion_b200 import synth
    from framefusion_b200.main import FrameFusion, find_contigious_latter_index
    wl = synth.make_workload(frames=3, patch_num=4, hidden=64, dtype=torch.bfloat16, seed=0)
    ff = FrameFusion()
    ff.prepare(*wl.prepare_args())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ff(wl.hidden, [wl.cos, wl.sin], None)
    # exported helper keeps the reference's known-answer vector (main.py:361-363)
    assert find_contigious_latter_index(torch.tensor([[0, 1, 1, 1, 0, 0, 1, 1]])).tolist() == [[0, 0, 0, 3, 0, 0, 0, 2]]
