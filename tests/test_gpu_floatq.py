"""GPU parity tests for FP8 / generic float_quantize / S2FP8.

The oracle here is a restatement of qtorch 0.2.0 (oracle/floatq.py: PARITY UNPINNED — qtorch is
absent from the image), so "bit-exact" means bit-exact against that restatement given the same
random integers.  S2FP8 adds transcendental functions: bit-exactness is asserted against the
oracle evaluated with torch CUDA ops on the same GPU and the same (mu, m); against the CPU
evaluation a tolerance is stated.
"""
import numpy as np
import pytest
import torch

from oracle import floatq as ofq
from oracle import s2fp8 as os2
from tests import cabi
from tests.golden_util import assert_bit_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def interesting_values(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g) * torch.pow(10.0, torch.randint(-8, 7, (n,), generator=g).float())
    special = torch.tensor([0.0, -0.0, 1.0, -1.0, 114688.0, -114688.0, 57344.0, 100000.0, 6.1035e-05, 3e-6, -3e-6,
                            1e-40, 3.4e38, -3.4e38, float("inf"), -float("inf"), float("nan"), 98304.0, 1.75, 1.25])
    x[: special.numel()] = special
    r = torch.randint(0, 2**31 - 1, (n,), generator=g, dtype=torch.int32)
    return x, r


@pytest.mark.parametrize("exp,man", [(5, 2), (4, 3), (5, 10), (8, 7), (2, 1)])
@pytest.mark.parametrize("check_inf", [True, False])
def test_float_quantize_bit_exact_vs_oracle(exp, man, check_inf):
    x, r = interesting_values(200003, seed=exp * 100 + man)
    want = ofq.float_quantize(x, exp, man, r, check_inf=check_inf)
    got = cabi.float_quantize(x.to(DEV), cabi.floatq_params(exp, man, check_inf=check_inf), rand_bits=r.to(DEV))
    assert_bit_equal(got.cpu(), want, f"e{exp}m{man}")


def test_float_quantize_nearest_and_new_max_rule():
    x, r = interesting_values(50000, seed=1)
    want = torch.from_numpy(ofq.qtorch_float_quantize(x.numpy(), 5, 2, "nearest", max_exp_bias=-1))
    got = cabi.float_quantize(x.to(DEV), cabi.floatq_params(5, 2, rounding=0, check_inf=False, max_exp_bias=-1))
    assert_bit_equal(got.cpu(), want, "nearest")
    assert float(ofq.max_value(5, 2)) == 114688.0 and float(ofq.max_value(5, 2, -1)) == 57344.0


def test_float_quantize_in_place_unaligned_and_philox():
    x, r = interesting_values(100001, seed=2)
    p = cabi.floatq_params(5, 2)
    want = ofq.float_quantize(x, 5, 2, r)
    xd = x.to(DEV)
    cabi.float_quantize(xd, p, rand_bits=r.to(DEV), out=xd)
    assert_bit_equal(xd.cpu(), want, "in place")
    bx = torch.empty(x.numel() + 1, device=DEV)
    br = torch.empty(x.numel() + 3, device=DEV, dtype=torch.int32)
    bx[1:].copy_(x)
    br[3:].copy_(r)
    got = cabi.float_quantize(bx[1:], p, rand_bits=br[3:])
    assert_bit_equal(got.cpu(), want, "unaligned")
    # in-kernel Philox: deterministic per (seed, offset); result always one of the two neighbours
    xs = torch.randn(1 << 20, generator=torch.Generator().manual_seed(5))
    a = cabi.float_quantize(xs.to(DEV), cabi.floatq_params(5, 2, seed=9, offset=1))
    b = cabi.float_quantize(xs.to(DEV), cabi.floatq_params(5, 2, seed=9, offset=1))
    c = cabi.float_quantize(xs.to(DEV), cabi.floatq_params(5, 2, seed=9, offset=2))
    assert torch.equal(a, b) and not torch.equal(a, c)
    down = ofq.float_quantize(xs, 5, 2, torch.zeros(xs.numel(), dtype=torch.int32))
    up = ofq.float_quantize(xs, 5, 2, torch.full((xs.numel(),), (1 << 21) - 1, dtype=torch.int32))
    a = a.cpu()
    assert bool(((a == down) | (a == up)).all())
    # unbiased in expectation
    assert abs((a.double() - xs.double()).mean().item()) < 5 * 0.25 / (1 << 10)


def test_fp8_plugin_matches_oracle():
    from argparse import ArgumentParser

    from smart_compress.compress import FP8, FP16, BF16

    x, r = interesting_values(70001, seed=3)
    for cls, (e, m) in ((FP8, (5, 2)), (FP16, (5, 10)), (BF16, (8, 7))):
        args = cls.add_argparse_args(ArgumentParser()).parse_args([])
        args.precision = 32
        y = cls(args)(x.to(DEV), tag="t", _rand_bits=r)
        assert_bit_equal(y.cpu(), ofq.float_quantize(x, e, m, r), cls.__name__)
    args = FP8.add_argparse_args(ArgumentParser()).parse_args(["--no_float_quantize_check_inf"])
    args.precision = 32
    y = FP8(args)(x.to(DEV), _rand_bits=r)
    assert_bit_equal(y.cpu(), ofq.float_quantize(x, 5, 2, r, check_inf=False), "no check_inf")


# ---- S2FP8 -----------------------------------------------------------------------------------------
def s2_input(n, seed, scale=0.01):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g) * scale
    x[::97] = 0.0
    r = torch.randint(0, 2**31 - 1, (n,), generator=g, dtype=torch.int32)
    return x, r


@pytest.mark.parametrize("n", [1000, (1 << 20) + 1])
def test_s2fp8_statistics(n):
    x, _ = s2_input(n, seed=n)
    mu, m = os2.s2fp8_statistics(x.double())
    got = cabi.s2fp8_stats(x.to(DEV)).cpu()
    assert abs(got[0].item() - mu.item()) <= 1e-6 * abs(mu.item())
    assert abs(got[1].item() - m.item()) <= 2e-7 * abs(m.item())  # max of log2: 1 ulp of libdevice vs fp64


@pytest.mark.parametrize("n", [1000, (1 << 18) + 5])
def test_s2fp8_apply_bit_exact_vs_torch_cuda_oracle(n):
    """Same GPU, same (mu, m), same random integers: the kernel and the reference's op chain
    evaluated by torch's CUDA operators must agree bit for bit."""
    x, r = s2_input(n, seed=n + 1)
    xd = x.to(DEV)
    mu_max = cabi.s2fp8_stats(xd)
    want, _, _, _ = os2.s2fp8(xd, r, mu=mu_max[0].clone(), m=mu_max[1].clone())
    got = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2), rand_bits=r.to(DEV))
    diff = (got.view(torch.int32) != want.view(torch.int32)) & ~(torch.isnan(got) & torch.isnan(want))
    assert int(diff.sum()) == 0, f"{int(diff.sum())} of {n} differ"


# ---- the screened path of the apply kernel (approximate pow decides, near-carry elements are recomputed) ---------
# the constants of float_quantize.cu (kLg2AbsErr, kLg2RelErr, kEx2RelErr): keep in step
S2_SCREEN_CONSTANTS = (4.8e-7, 3.0e-7, 4.8e-7)


def test_s2fp8_screen_error_constants_hold_for_every_input_of_the_special_function_unit():
    """The margin the screened path trusts is built from three error constants of lg2.approx / ex2.approx.  They
    are measured here over EVERY normal input of lg2 and every input of ex2 in [-126, 128) against fp64, and must
    leave 20 % of headroom; the end-to-end bound |bits(v~) - bits(v)| <= margin is measured over 2^26 random
    (alpha, beta, magnitude) triples and must leave headroom too."""
    lg2_abs, lg2_rel, ex2_rel, ratio, margin_max, used = cabi.selftest_s2_screen(1 << 26)
    print(f"lg2 abs {lg2_abs:.3e} rel {lg2_rel:.3e}  ex2 rel {ex2_rel:.3e}  bound ratio {ratio:.3f}  "
          f"margin max {margin_max:.0f}  samples {used:.0f}")
    assert 0 < lg2_abs <= S2_SCREEN_CONSTANTS[0] / 1.2
    assert 0 < lg2_rel <= S2_SCREEN_CONSTANTS[1] / 1.2
    assert 0 < ex2_rel <= S2_SCREEN_CONSTANTS[2] / 1.2
    assert used > (1 << 24)
    assert 0 < ratio <= 0.8
    assert margin_max <= 4096


def _s2_scale_case(case, n, seed):
    g = torch.Generator().manual_seed(seed)
    if case == "unit":
        x = torch.randn(n, generator=g)
    elif case == "gradient":       # beta ~ +60: the margin grows with |beta|
        x = torch.randn(n, generator=g) * 3e-6
    elif case == "large":          # beta negative
        x = torch.randn(n, generator=g) * 2e5
    elif case == "relu":
        x = torch.randn(n, generator=g).clamp_(min=0)
    elif case == "heavy_tail":     # small alpha
        x = torch.randn(n, generator=g) * torch.exp2(4 * torch.randn(n, generator=g))
    else:
        raise KeyError(case)
    return x


@pytest.mark.parametrize("rounding", ["stochastic", "nearest"])
@pytest.mark.parametrize("case", ["unit", "gradient", "large", "relu", "heavy_tail"])
def test_s2fp8_apply_screened_path_bit_exact_at_2_pow_24(case, rounding):
    """2^24 elements per case: a few thousand of them fall within the margin of a carry and are recomputed; all
    of them must carry the bits torch's CUDA operators give."""
    n = 1 << 24
    xd = _s2_scale_case(case, n, seed=len(case)).to(DEV)
    mu_max = cabi.s2fp8_stats(xd)
    if rounding == "stochastic":
        r = torch.randint(0, 2**31 - 1, (n,), generator=torch.Generator(device=DEV).manual_seed(5), dtype=torch.int32,
                          device=DEV)
        got = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2), rand_bits=r)
    else:
        r = torch.full((n,), 1 << 20, dtype=torch.int32, device=DEV)
        got = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2, rounding=0))
    want, _, _, _ = os2.s2fp8(xd, r, mu=mu_max[0].clone(), m=mu_max[1].clone())
    diff = (got.view(torch.int32) != want.view(torch.int32)) & ~(torch.isnan(got) & torch.isnan(want))
    assert int(diff.sum()) == 0, f"{case}: {int(diff.sum())} of {n} differ, first at {int(diff.nonzero()[0])}"


@pytest.mark.parametrize("case", ["unit", "gradient", "relu"])
def test_s2fp8_apply_adversarial_random_numbers_at_the_carry(case):
    """Random integers chosen so that bits(v) + r lands within +-3 of a multiple of 2^21 for EVERY element: the
    rounding direction then hangs on the last bits of powf.  An approximate pow that was trusted there would be
    wrong on about half of them; the screen must send all of them to the exact path."""
    n = 1 << 20
    xd = _s2_scale_case(case, n, seed=9).to(DEV)
    mu_max = cabi.s2fp8_stats(xd)
    mu, m = mu_max[0].clone(), mu_max[1].clone()
    alpha = 15.0 / (m - mu)
    beta_pow2 = 2.0 ** (-alpha * mu)
    pre = xd.abs().pow_(alpha).mul_(beta_pow2)            # v, as the reference's chain computes it on this GPU
    d = torch.randint(-3, 4, (n,), device=DEV, dtype=torch.int32)
    r = (-(pre.view(torch.int32)) + d) & ((1 << 21) - 1)  # bits(v) + r == d (mod 2^21)
    got = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2), rand_bits=r)
    want, _, _, _ = os2.s2fp8(xd, r, mu=mu, m=m)
    diff = (got.view(torch.int32) != want.view(torch.int32)) & ~(torch.isnan(got) & torch.isnan(want))
    assert int(diff.sum()) == 0, f"{case}: {int(diff.sum())} of {n} differ, first at {int(diff.nonzero()[0])}"


def test_s2fp8_plugin_vs_cpu_oracle_tolerance():
    """Against the CPU evaluation (different libm): the quantised intermediate may flip on a
    vanishing fraction of elements; everywhere else the result agrees to 4 ulp."""
    from argparse import ArgumentParser

    from smart_compress.compress import S2FP8

    x, r = s2_input(200000, seed=8)
    args = S2FP8.add_argparse_args(ArgumentParser()).parse_args([])
    args.precision = 32
    y = S2FP8(args)(x.to(DEV), tag="t", _rand_bits=r).cpu()
    want, mu, m, _ = os2.s2fp8(x, r)
    close = torch.isclose(y, want, rtol=4 * 1.2e-7 * 8, atol=0)  # pow amplifies by 1/alpha ~ a few
    assert close.float().mean().item() > 0.999
    assert bool(torch.isfinite(y).all())
    assert ((y == 0) == (x == 0)).all()


def _s2_cases(case):
    g = torch.Generator().manual_seed(len(case) * 17)
    n = 40003
    if case == "wide":  # |alpha * log2 a| far beyond 125: pow over- and underflows, quads decline
        x = torch.exp2(torch.rand(n, generator=g) * 120 - 60) * torch.sign(torch.randn(n, generator=g))
        return x, torch.tensor([-2.0, 1.0])
    if case == "specials":  # inf / NaN / denormals / +-0 / 1.0 inside otherwise ordinary quads
        x, _ = interesting_values(n, seed=11)
        x = x * 1e-3
        x[:20] = torch.tensor([0.0, -0.0, 1.0, -1.0, 1e-40, -1e-40, 1e-45, 1.1754944e-38, -1.1754942e-38, 3.4e38,
                               float("inf"), -float("inf"), float("nan"), 0.5, 2.0, 1.0000001, 0.99999994, 3e-39,
                               -0.0, 7.0])
        return x, torch.tensor([-9.0, 3.0])
    if case == "relu":  # half the elements are zeros of either sign
        x = torch.randn(n, generator=g)
        x[x < 0] = 0.0
        x[::5] = -0.0
        return x, None
    if case == "tiny":  # magnitudes around FLT_MIN, denormals among them
        x = torch.randn(n, generator=g) * 1e-37
        x[::11] *= 1e-3
        return x, None
    if case == "steep":  # narrow distribution: alpha ~ 50, exponents saturate both ways
        x = (1.0 + 0.05 * torch.randn(n, generator=g)) * 3.0
        return x, None
    if case == "constant":  # m == mu: alpha = inf, every scalar degenerate
        return torch.full((n,), 0.37), None
    raise KeyError(case)


@pytest.mark.parametrize("rounding", ["stochastic", "nearest"])
@pytest.mark.parametrize("case", ["wide", "specials", "relu", "tiny", "steep", "constant"])
def test_s2fp8_apply_bit_exact_where_the_fast_path_declines(case, rounding):
    """The packed pow fast path covers ordinary quads only; zeros, denormals, non-finite inputs, results outside
    the normal range and degenerate scalars go to the direct formula.  Either way: the bits torch's CUDA ops give."""
    x, mu_max = _s2_cases(case)
    xd = x.to(DEV)
    mu_max = cabi.s2fp8_stats(xd) if mu_max is None else mu_max.to(DEV)
    g = torch.Generator().manual_seed(3)
    if rounding == "stochastic":
        r = torch.randint(0, 2**31 - 1, (x.numel(),), generator=g, dtype=torch.int32)
        got = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2), rand_bits=r.to(DEV))
    else:
        r = torch.full((x.numel(),), 1 << 20, dtype=torch.int32)  # adding half a step, then truncating == nearest
        got = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2, rounding=0))
    want, _, _, _ = os2.s2fp8(xd, r, mu=mu_max[0].clone(), m=mu_max[1].clone())
    diff = (got.view(torch.int32) != want.view(torch.int32)) & ~(torch.isnan(got) & torch.isnan(want))
    assert int(diff.sum()) == 0, f"{case}: {int(diff.sum())} of {x.numel()} differ, first at {int(diff.nonzero()[0])}"


def test_s2fp8_apply_in_kernel_philox_rounds_to_a_neighbour():
    x, _ = s2_input((1 << 18) + 3, seed=21)
    xd = x.to(DEV)
    mu_max = cabi.s2fp8_stats(xd)
    a = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2, seed=4, offset=1))
    b = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2, seed=4, offset=1))
    c = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2, seed=4, offset=2))
    assert torch.equal(a, b) and not torch.equal(a, c)
    lo = torch.zeros(x.numel(), dtype=torch.int32, device=DEV)
    down = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2), rand_bits=lo)
    up = cabi.s2fp8_apply(xd, mu_max, cabi.floatq_params(5, 2), rand_bits=lo + ((1 << 21) - 1))
    assert bool(((a == down) | (a == up)).all())
    assert 0.2 < (a == up).float().mean().item() < 0.8


@pytest.mark.parametrize("case", ["relu", "tiny", "steep", "nan", "inf", "all_zero", "one_zero"])
def test_s2fp8_statistics_chunks_with_zeros_subnormals_and_non_finite(case):
    """The statistics pass takes a lean route for chunks of normal numbers and a per-element one for chunks that
    hold a zero (L = 0), a subnormal or a NaN: both against the fp64 evaluation of s2fp8.py:33-37."""
    if case in ("relu", "tiny", "steep"):
        x, _ = _s2_cases(case)
    else:
        x = torch.randn(100003, generator=torch.Generator().manual_seed(5)) * 0.3
        if case == "nan":
            x[77777] = float("nan")
        elif case == "inf":
            x[5] = -float("inf")
        elif case == "all_zero":
            x.zero_()
        elif case == "one_zero":
            x = x.abs() * 0.01 + 1e-4  # every log2 negative: the single zero's L = 0 is the maximum
            x[31337] = 0.0
    mu, m = os2.s2fp8_statistics(x.double())
    got = cabi.s2fp8_stats(x.to(DEV)).cpu()
    for g, w, tol in ((got[0].item(), mu.item(), 1e-6), (got[1].item(), m.item(), 2e-7)):
        if w != w:
            assert g != g
        elif abs(w) == float("inf") or w == 0.0:
            assert g == w
        else:
            assert abs(g - w) <= tol * abs(w), (case, g, w)


@pytest.mark.parametrize("y", [0.37, 1.0, 1.5, 2.0, 3.75, 7.3, 0.015625, 21.4, 1e-3, 63.0])
def test_s2fp8_fast_pow_is_torch_pow_bit_for_bit(y):
    """The packed pow restates libdevice's powf main line; here its RAW result (no quantisation to hide a last-bit
    difference) against torch.pow with a tensor exponent — the op s2fp8.py:44 runs — over every binade."""
    g = torch.Generator().manual_seed(int(y * 1000) + 1)
    n = 1 << 22
    bits = torch.randint(0x00800000, 0x7F800000, (n,), generator=g, dtype=torch.int32)  # every normal float
    a = bits.view(torch.float32).clone()
    near_one = 1.0 + (torch.rand(n // 4, generator=g) - 0.5) * 2.0 ** torch.randint(-23, 0, (n // 4,), generator=g).float()
    a[: n // 4] = near_one.abs()
    a[n // 4: n // 4 + 8] = torch.tensor([1.0, 2.0, 0.5, 1.17549435e-38, 3.4028234663852886e38, 0.99999994, 1.0000001, 4.0])
    ad = a.to(DEV)
    want = torch.pow(ad, torch.tensor(y, dtype=torch.float32, device=DEV))
    got, acc = cabi.selftest_pow(ad, y)
    diff = got.view(torch.int32) != want.view(torch.int32)
    assert int(diff.sum()) == 0, f"y={y}: {int(diff.sum())} differ, e.g. a={a[diff.cpu()][:4].tolist()}"
    frac = acc.float().mean().item()
    # accepted iff |y * log2 a| <= 125 for the whole quad: most quads for small y, a shrinking share as y grows
    assert frac > 0.0
    if y <= 0.37:
        assert frac > 0.99


def test_s2fp8_fast_pow_declines_what_it_must():
    a = torch.tensor([0.0, 1.0, 1.0, 1.0,   1e-40, 1.0, 1.0, 1.0,   float("inf"), 1.0, 1.0, 1.0,
                      float("nan"), 1.0, 1.0, 1.0,   2.0 ** 100, 1.0, 1.0, 1.0,   2.0 ** -100, 1.0, 1.0, 1.0,
                      3.0, 1.0, 0.5, 7.0], device=DEV)
    got, acc = cabi.selftest_pow(a, 1.5)
    assert acc.tolist() == [0, 0, 0, 0, 0, 0, 1]
    want = torch.pow(a, torch.tensor(1.5, device=DEV))
    same = (got.view(torch.int32) == want.view(torch.int32)) | (torch.isnan(got) & torch.isnan(want))
    assert bool(same.all())


# ---- many tensors in two launches (the optimizer side with --compress fp8 | fp16 | bf16) ----------------------
MANY_SIZES = [1, 7, 8, 513, 16384, 16385, 40003, 5, 100000, 16384 * 3]


def _many_tensors(seed=0):
    g = torch.Generator().manual_seed(seed)
    ts = [(torch.randn(n, generator=g) * (10.0 ** ((i % 5) - 2))).to(DEV) for i, n in enumerate(MANY_SIZES)]
    buf = (torch.randn(20001, generator=g)).to(DEV)
    ts.append(buf[1:])            # contiguous, pointer not 32-byte aligned: element path
    return ts


@pytest.mark.parametrize("rounding", [1, 0])
@pytest.mark.parametrize("exp,man", [(5, 2), (8, 7)])
def test_float_quantize_multi_equals_the_per_tensor_calls(exp, man, rounding):
    import ctypes as C
    from smart_compress import _native as N
    lib = N.load()
    ts = _many_tensors(seed=exp * 10 + man)
    base = 1000
    want = []
    for j, t in enumerate(ts):
        want.append(cabi.float_quantize(t, cabi.floatq_params(exp, man, rounding=rounding, seed=11, offset=base + j)))
    outs = [torch.full_like(t, float("nan")) for t in ts]
    host = (N.TensorDesc * len(ts))()
    for j, (t, o) in enumerate(zip(ts, outs)):
        host[j].x, host[j].y, host[j].n, host[j].all_positive, host[j].stream = t.data_ptr(), o.data_ptr(), t.numel(), 0, j
    descs = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(DEV)
    need = lib.smaq_floatq_multi_workspace_bytes(len(ts))
    ws = torch.full((need,), 0xA5, dtype=torch.uint8, device=DEV)
    p = cabi.floatq_params(exp, man, rounding=rounding, seed=11, offset=base)
    for _ in range(2):   # reused scratch
        N.check(lib.smaq_float_quantize_multi(descs.data_ptr(), len(ts), sum(t.numel() for t in ts), C.byref(p),
                                              ws.data_ptr(), ws.numel(), N.stream_ptr(DEV)), "float_quantize_multi")
        for j, (o, w) in enumerate(zip(outs, want)):
            assert_bit_equal(o, w, f"tensor {j} ({ts[j].numel()} elements)")
    assert lib.smaq_float_quantize_multi(descs.data_ptr(), len(ts), 1, C.byref(p), ws.data_ptr(), 4, N.stream_ptr(DEV)) == 3
    assert lib.smaq_float_quantize_multi(None, 0, 0, C.byref(p), None, 0, N.stream_ptr(DEV)) == 0


def test_fp8_plugin_compress_many_equals_the_loop():
    """FP8.compress_many (what OptimLP calls per phase) numbers its Philox streams like the per-tensor loop."""
    import itertools
    from argparse import ArgumentParser
    from smart_compress.compress.fp8 import FP8
    from smart_compress.util.pytorch import quantization as Q
    hp = FP8.add_argparse_args(ArgumentParser()).parse_args([])
    hp.precision = 32
    fp = FP8(hp)
    ts = _many_tensors(seed=3)
    torch.manual_seed(5)
    Q._calls = itertools.count(700)
    want = [fp(t.clone(), tag="optimizer_grad") for t in ts]
    Q._calls = itertools.count(700)
    mine = [t.clone() for t in ts]
    got = fp.compress_many(mine, None, tag="optimizer_grad")
    for j, (g, w, m) in enumerate(zip(got, want, mine)):
        assert g is m                                   # in place, same objects
        assert_bit_equal(g, w, f"tensor {j}")


def test_s2fp8_compress_many_matches_the_reference_chain_per_tensor():
    """S2FP8.compress_many (what OptimLP calls per phase: smaq_s2fp8_multi, three launches) against the reference's
    op chain evaluated by torch's CUDA operators on this GPU, per tensor: with the (mu, m) the kernel reports — which
    must be within 1e-6 of the reference's — and the random fields the kernel drew (oracle/rng.py), bit for bit."""
    from argparse import ArgumentParser

    from oracle import rng as orng
    from smart_compress.compress.s2fp8 import S2FP8

    hp = S2FP8.add_argparse_args(ArgumentParser()).parse_args([])
    hp.precision = 32
    g = torch.Generator().manual_seed(3)
    sizes = [10, 512, 4096, 5000, 16384, 16385, 70000, (1 << 18) + 3]
    xs = [torch.randn(n, generator=g) * (0.01 * (i + 1)) for i, n in enumerate(sizes)]
    xs[2][::7] = 0.0
    torch.manual_seed(77)
    codec = S2FP8(hp)
    from smart_compress.util.pytorch import quantization as Q
    import itertools
    Q._calls = itertools.count()            # stream numbers from 0 for this test
    mine = [x.to(DEV) for x in xs]
    stats = {}
    got = codec.compress_many(mine, None, tag="optimizer_grad", stats_out=stats)
    for i, (x, n) in enumerate(zip(xs, sizes)):
        assert got[i] is mine[i]
        mm = stats[i]
        mu_ref, m_ref = os2.s2fp8_statistics(x.double())
        assert abs(mm[0].item() - mu_ref.item()) <= 1e-6 * abs(mu_ref.item()), (i, n)
        assert abs(mm[1].item() - m_ref.item()) <= 2e-7 * abs(m_ref.item()), (i, n)
        r = torch.from_numpy(orng.floatq_fields(n, seed=77, offset=i, man_bits=2))
        want, _, _, _ = os2.s2fp8(x.to(DEV), r, mu=mm[0].clone(), m=mm[1].clone())
        diff = (got[i].view(torch.int32) != want.view(torch.int32)) & ~(torch.isnan(got[i]) & torch.isnan(want))
        assert int(diff.sum()) == 0, f"tensor {i} ({n} elements): {int(diff.sum())} differ"


def test_float_quantize_in_kernel_fields_match_the_oracle_generator():
    """FP8's performance path (no rand_bits): the fields the kernel adds are oracle/rng.py's, so the output is the
    oracle's bit for bit."""
    from oracle import rng as orng

    x, _ = interesting_values(100003, seed=8)
    for man, exp in ((2, 5), (3, 4), (10, 5), (7, 8)):
        got = cabi.float_quantize(x.to(DEV), cabi.floatq_params(exp, man, seed=5, offset=(1 << 35) + 3))
        r = torch.from_numpy(orng.floatq_fields(x.numel(), seed=5, offset=(1 << 35) + 3, man_bits=man))
        want = ofq.float_quantize(x, exp, man, r)
        assert torch.equal(got.cpu().view(torch.int32), want.view(torch.int32)), (exp, man)
