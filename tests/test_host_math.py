"""CPU checks of the product's per-element arithmetic (csrc/smaq_math.cuh compiled for the host by
tests/host_math_harness.cpp) against the oracle and the reference-generated golden vectors.

This is how the literal operation sequence — and the three-instruction correctly rounded division
the kernels use in place of a per-element IEEE divide — is validated without a GPU.  The harness
is test-only; the product has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle.smaq import SmaqConfig, smaq_roundtrip
from tests.golden_util import assert_bit_equal, load_golden, uses_bn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = load_golden()


@pytest.fixture(scope="session")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("harness") / "libharness.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "host_math_harness.cpp"), "-o", str(out)])
    lib = C.CDLL(str(out))
    lib.harness_div_check.restype = C.c_int64
    lib.harness_div_check.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_int]
    lib.harness_uniform24.restype = C.c_float
    lib.harness_uniform24.argtypes = [C.c_uint32]
    return lib


def run_roundtrip(lib, x, probs, cfg, mean, std, *, all_positive=False, saturate=False, variant=1):
    x = np.ascontiguousarray(x.numpy().reshape(-1), dtype=np.float32)
    y = np.empty_like(x)
    codes = np.empty_like(x)
    p = None if probs is None else np.ascontiguousarray(probs.numpy().reshape(-1), dtype=np.float32)
    used_fast = C.c_int(0)
    f = C.c_float
    lib.harness_roundtrip(
        x.ctypes.data_as(C.c_void_p), None if p is None else p.ctypes.data_as(C.c_void_p),
        y.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p), C.c_int64(x.size),
        f(float(mean)), f(float(std)), f(cfg.main_std_dev_threshold), f(cfg.range_normal), f(cfg.range_outlier),
        f(cfg.clamped_range[0]), f(cfg.clamped_range[1]), cfg.num_bits_main, cfg.num_bits_outlier,
        int(cfg.stochastic_rounding), int(all_positive), int(saturate), variant, C.byref(used_fast))
    return torch.from_numpy(y), torch.from_numpy(codes), used_fast.value == 1


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if not c["same_object"] and not uses_bn(c)))
def test_kernel_math_matches_reference_golden(harness, name, variant):
    c = CASES[name]
    ref = smaq_roundtrip(c["x"].clone(), c["cfg"], probs=c["probs"], idx=c["idx"], **c["kwargs"])
    y, codes, _ = run_roundtrip(harness, c["x"], c["probs"], c["cfg"], ref.mean, ref.std, variant=variant,
                                all_positive=c["kwargs"].get("all_positive", False))  # (the affine wrap of --use_batch_norm is kernel-side only)
    assert_bit_equal(y.view(c["y"].shape), c["y"], name)
    assert_bit_equal(codes.view(c["y"].shape), ref.code, name + " codes")


@pytest.mark.parametrize("stochastic", [True, False])
def test_kernel_math_large_random(harness, stochastic):
    g = torch.Generator().manual_seed(42)
    n = 1 << 22
    x = torch.randn(n, generator=g)
    x[torch.randperm(n, generator=g)[: n // 100]] *= 10
    probs = torch.rand(n, generator=g)
    cfg = SmaqConfig(stochastic_rounding=stochastic)
    for saturate in (False, True):
        ref = smaq_roundtrip(x, cfg, probs=probs, saturate=saturate)
        y, codes, fast = run_roundtrip(harness, x, probs, cfg, ref.mean, ref.std, saturate=saturate)
        assert fast
        assert_bit_equal(y, ref.y, "y")
        assert_bit_equal(codes, ref.code, "codes")


def test_fast_division_is_correctly_rounded(harness):
    """div_rn (multiply, FMA remainder, FMA correction) against the IEEE divide on 2^26 random
    numerators per divisor, adversarial divisors included (all-ones significands, powers of two,
    the default ranges 15 and 42, and the [2^-60, 2^60] validity edges)."""
    rng = np.random.default_rng(0)
    n = 1 << 24
    a = rng.standard_normal(n).astype(np.float32) * np.float32(10.0) ** rng.integers(-20, 20, n).astype(np.float32)
    a[:8] = [0.0, -0.0, np.inf, -np.inf, np.nan, 3.4e38, -3.4e38, 1e-45]
    # numerators of the second division are rounded codes: every integer up to 2^17 either way, then
    # random integer-valued floats of any magnitude (incl. inf/NaN from non-finite inputs)
    ints = np.concatenate([np.arange(-(1 << 17), 1 << 17, dtype=np.float32),
                           np.trunc(a[np.abs(a) >= 1.0]), a[:7]]).astype(np.float32)
    divisors = [15.0, 42.0, 1.0, 1.4142135, 0.70710677, 3.0, 1.9999999, 1.0000001, 0.99999994, 7.0 / 1.5,
                63.0 / 1.6, 2.0 ** -60, 2.0 ** 60, 1e-12, 1e12, 3.3665016, 1.41]
    divisors += list(rng.uniform(0.5, 2.0, 24)) + list(np.float32(10.0) ** rng.uniform(-15, 15, 24))
    for b in divisors:
        b = float(np.float32(b))
        assert harness.harness_div_check(a.ctypes.data_as(C.c_void_p), n, b, 1) == 0, b
        if 2.0 ** -20 <= b <= 2.0 ** 20:
            assert harness.harness_div_check(ints.ctypes.data_as(C.c_void_p), ints.size, b, 0) == 0, b


def test_uniform24_grid(harness):
    assert harness.harness_uniform24(0) == 0.0
    assert harness.harness_uniform24(0xFFFFFFFF) == 1.0 - 2.0 ** -24
    assert harness.harness_uniform24(0x100) == 2.0 ** -24
