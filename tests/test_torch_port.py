"""CPU: pins the torch-CPU port (``oracle/ff_torch_port.py`` — the CPU baseline ``bench.py`` times) against the
fixtures produced by the unmodified reference, through the same harness the oracle and the CUDA path go through."""
import json
import os

import numpy as np
import pytest
import torch

from _harness import GOLDEN_DIR, DT, case_names, run_and_compare
from oracle import ff_oracle as orc
from oracle import ff_torch_port as port


class PortAdapter:
    def __init__(self, cost, slb, rlb, dtype):
        self.ff = port.TorchPortFrameFusion(cost, slb, rlb)

    def prepare(self, *args):
        self.ff.prepare(*args)

    def __call__(self, hidden, pos, mask, attn=None):
        return self.ff(hidden, pos, mask, attn)

    finish_merging = property(lambda s: s.ff.finish_merging)
    finish_pruning = property(lambda s: s.ff.finish_pruning)
    sparsity_list = property(lambda s: s.ff.sparsity_list)
    patch_type = property(lambda s: s.ff.patch_type)

    def last_trace(self):
        return self.ff.last


@pytest.mark.parametrize("name", case_names())
def test_port_matches_reference_sequence(name):
    rep = run_and_compare(name, PortAdapter)
    assert rep["n_sim"] > 0


def test_port_budget_and_importance():
    z = np.load(os.path.join(GOLDEN_DIR, "statics.npz"))
    lists = ([], [0.39], [0.39, 0.2], [0.5, 0.4, 0.3, 0.05], [0.0, 0.0])
    i = 0
    for cost in (0.2, 0.3, 0.5, 0.7, 0.9):
        for sl in lists:
            want = z["budget"][i][2]
            i += 1
            if np.isnan(want):
                with pytest.raises(ValueError, match="The cost is too small"):
                    port.pruning_ratio(sl, cost)
            else:
                assert port.pruning_ratio(sl, cost) == want
    from framefusion_b200 import synth
    zi = np.load(os.path.join(GOLDEN_DIR, "importance.npz"))
    for tag in "abcdef":
        spec = json.loads(str(zi[f"imp_{tag}_spec"]))
        dt = spec["dtype"]
        q, k = synth.make_attention_inputs(spec["s_len"], 28, 4, 128, DT[dt], seed=spec["s_len"])
        got = port.last_query_attention(q, k, spec["num"], spec["causal"])[0].float().numpy()
        want = orc.bits_to_f32(zi[f"imp_{tag}_out"], dt).reshape(got.shape)
        if dt == "f32":
            assert np.allclose(got, want, rtol=2e-5, atol=1e-9)
        else:
            assert np.array_equal(got, want)          # same ATen kernels as the reference: bit equal
