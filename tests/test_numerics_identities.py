"""CPU checks of the two arithmetic identities the gather kernel relies on (framefusion_b200/csrc/ff_common.cuh):

* Divider: for bf16, T(x * RN(1/n)) == T(x / n) for EVERY finite x and every run length n <= 256, so the kernel may
  multiply by the reciprocal instead of dividing (main.py:314-317 divides); f16 does have exceptions, which is why the
  kernel keeps the IEEE division there.
* add_vec: rounding the exact sum of two T values once (what add.bf16x2 / add.f16x2 do) equals the float32 add followed
  by the rounding to T (what the reference's index_add_ does), checked on random pairs and on a sweep of exponent gaps.
"""
import torch


def all_finite(T):
    bits = torch.arange(65536, dtype=torch.int32).to(torch.int16)
    x = bits.view(T).to(torch.float32)
    return x[torch.isfinite(x)]


def test_bf16_reciprocal_multiply_is_exact_for_all_inputs():
    x = all_finite(torch.bfloat16)
    for n in range(1, 257):
        d = torch.tensor(float(n), dtype=torch.bfloat16).to(torch.float32)
        want = (x / d).to(torch.bfloat16)
        got = (x * (torch.tensor(1.0) / d)).to(torch.bfloat16)
        assert torch.equal(want.view(torch.int16), got.view(torch.int16)), f"n={n}"


def test_f16_reciprocal_multiply_is_not_exact():
    x = all_finite(torch.float16)
    d = torch.tensor(14.0)
    want, got = (x / d).to(torch.float16), (x * (torch.tensor(1.0) / d)).to(torch.float16)
    assert not torch.equal(want.view(torch.int16), got.view(torch.int16))      # the kernel divides for f16
    for n in (2, 4, 8, 64, 256):                                               # power-of-two runs are exact scalings
        d = torch.tensor(float(n))
        assert torch.equal((x / d).to(torch.float16).view(torch.int16), (x * (1.0 / d)).to(torch.float16).view(torch.int16))


def test_single_rounding_of_the_exact_sum_equals_float32_add_then_round():
    g = torch.Generator().manual_seed(0)
    for T in (torch.bfloat16, torch.float16):
        a = (torch.randn(1 << 20, generator=g) * 4).to(T)
        scale = torch.exp2(torch.randint(-20, 8, (1 << 20,), generator=g).float())
        b = (torch.randn(1 << 20, generator=g) * scale).to(T)
        via_f32 = (a.float() + b.float()).to(T)
        exact_once = (a.double() + b.double()).to(T)                           # float64 holds the sum of two T values exactly
        assert torch.equal(via_f32.view(torch.int16), exact_once.view(torch.int16))
