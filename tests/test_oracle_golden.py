"""CPU: pins the numpy oracle against fixtures produced by the unmodified reference (oracle/gen_golden.py)
and against the one known-answer vector the reference carries (main.py:361-363)."""
import json
import os

import numpy as np
import pytest

from _harness import GOLDEN_DIR, OracleAdapter, case_names, run_and_compare
from oracle import ff_oracle as orc


def test_reference_docstring_kat():
    # /root/reference/framefusion/main.py:361-363
    got = orc.find_contiguous_latter_index(np.array([0, 1, 1, 1, 0, 0, 1, 1]))
    assert got.tolist() == [0, 0, 0, 3, 0, 0, 0, 2]


def test_statics_vs_reference():
    z = np.load(os.path.join(GOLDEN_DIR, "statics.npz"))
    assert np.array_equal(orc.find_contiguous_latter_index(z["runs_kat_in"][0]), z["runs_kat_out"][0])
    for row_in, row_out in zip(z["runs_rand_in"], z["runs_rand_out"]):
        assert np.array_equal(orc.find_contiguous_latter_index(row_in), row_out)
    lists = ([], [0.39], [0.39, 0.2], [0.5, 0.4, 0.3, 0.05], [0.0, 0.0])
    i = 0
    for cost in (0.2, 0.3, 0.5, 0.7, 0.9):
        for sl in lists:
            want = z["budget"][i][2]
            i += 1
            if np.isnan(want):
                with pytest.raises(ValueError, match="The cost is too small"):
                    orc.compute_pruning_ratio(sl, cost)
            else:
                assert orc.compute_pruning_ratio(sl, cost) == want      # exact double arithmetic


@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
def test_cosine_chain_vs_reference(dtype):
    z = np.load(os.path.join(GOLDEN_DIR, "statics.npz"))
    a = orc.bits_to_f32(z[f"cos_{dtype}_a"], dtype)
    b = orc.bits_to_f32(z[f"cos_{dtype}_b"], dtype)
    want = orc.bits_to_f32(z[f"cos_{dtype}_out"], dtype)
    hidden = np.concatenate([a, b])                     # pair (i, i + n) as a two-token chain per patch id
    n = a.shape[0]
    pt = np.concatenate([np.arange(n), np.arange(n)])
    sr = orc.similarity_by_patch(hidden, pt, n, dtype)
    got, lo, hi, frag = sr.sim[1::2], sr.lo[1::2], sr.hi[1::2], sr.fragile[1::2]
    if dtype == "f32":
        assert np.allclose(got, want, rtol=0, atol=1e-6)
    else:
        assert np.array_equal(got[~frag], want[~frag])
        assert ((want >= lo) & (want <= hi)).all()
        assert frag.mean() < 0.05


def test_threshold_is_cast_to_tensor_dtype():
    # bf16(0.6) = 0.6015625, bf16(0.7) = 0.69921875
    assert float(orc.threshold_in_dtype(0.6, "bf16")) == 0.6015625
    assert float(orc.threshold_in_dtype(0.7, "bf16")) == 0.69921875
    assert float(orc.threshold_in_dtype(0.6, "f16")) == 0.60009765625


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference_sequence(name):
    rep = run_and_compare(name, lambda c, s, r, dt: OracleAdapter(c, s, r, dt))
    assert rep["n_sim"] > 0


def test_importance_vs_reference():
    from framefusion_b200 import synth
    from oracle.gen_golden import DT
    z = np.load(os.path.join(GOLDEN_DIR, "importance.npz"))
    for tag in "abcdef":
        spec = json.loads(str(z[f"imp_{tag}_spec"]))
        dt = spec["dtype"]
        q, k = synth.make_attention_inputs(spec["s_len"], 28, 4, 128, DT[dt], seed=spec["s_len"])
        got = orc.last_query_attention(q[0].float().numpy(), k[0].float().numpy(), spec["num"], dt, spec["causal"])
        want = orc.bits_to_f32(z[f"imp_{tag}_out"], dt).reshape(got.shape)
        if dt == "f32":
            assert np.allclose(got, want, rtol=2e-5, atol=1e-9)
        else:
            ulp = np.abs(want) * (2.0 ** -7 if dt == "bf16" else 2.0 ** -10) + 1e-30
            assert (np.abs(got - want) <= ulp * 1.01).all()
            assert (got != want).mean() < 0.02, (tag, (got != want).mean())
        # mean over (heads, num) as torch-CPU does it
        gm = orc.mean_heads(want, dt)
        wm = orc.bits_to_f32(z[f"imp_{tag}_mean"], dt)
        if dt == "f32":
            assert np.allclose(gm, wm, rtol=1e-5, atol=1e-12)
        else:
            assert (gm != wm).mean() < 0.01, (tag, (gm != wm).mean())
