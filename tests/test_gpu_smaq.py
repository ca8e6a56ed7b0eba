"""GPU parity tests for the SmaQ round trip and its statistics, through the C ABI and the plugin.

Contract (BASELINE.json north_star):
  * given the reference's mean/std and the same uniform numbers, the decoded fp32 values are
    BIT-EXACT against the oracle / the reference-generated golden vectors (target; the stated
    tolerance is 1 ulp);
  * mean/std within 1e-6 relative (of the std scale for a near-zero mean).
"""
import math

import pytest
import torch

from oracle.smaq import SmaqConfig, full_mean_std, sample_mean_std, smaq_roundtrip, std_of
from tests import cabi
from tests.golden_util import assert_bit_equal, load_golden, uses_bn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = load_golden()


def make_outlier_tensor(n, seed=1234):
    """BASELINE.md §4 input recipe (generated on CPU so the oracle sees identical bits)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g)
    idx = torch.randperm(n, generator=g)[: max(1, n // 100)]
    x[idx] *= 10
    return x, g


# ---- round trip, explicit statistics and probs: bit-exact ------------------------------------------
@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if not c["same_object"]))
def test_roundtrip_matches_reference_golden(name):
    c = CASES[name]
    cfg = c["cfg"]
    x = c["x"]
    ref = smaq_roundtrip(x.clone(), cfg, probs=c["probs"], idx=c["idx"], **c["kwargs"])
    assert_bit_equal(ref.y, c["y"], "oracle drifted from golden")
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    params = cabi.codec_params(cfg, all_positive=c["kwargs"].get("all_positive", False))
    xd = x.to(DEV).contiguous().view(-1)
    pd = None if c["probs"] is None else c["probs"].to(DEV).contiguous().view(-1)
    if uses_bn(c):  # --use_batch_norm: the affine wrap runs inside the kernel (smart.py:136-149,174-179)
        gamma, beta = c["kwargs"]["batch_norm_stats"]
        if cfg.bn_scalar_params:  # smart.py:140-142 (the means as the reference's CPU run formed them)
            gamma, beta, channels, inner = gamma.mean().reshape(1), beta.mean().reshape(1), 1, x.numel()
        else:
            channels, inner = x.shape[1], x.shape[2] * x.shape[3]
        y = cabi.roundtrip_bn(xd, ms, params, gamma.to(DEV), beta.to(DEV), channels, inner, probs=pd)
    else:
        y = cabi.roundtrip(xd, ms, params, probs=pd)
    assert_bit_equal(y.cpu().view(x.shape), c["y"], name)


@pytest.mark.parametrize("n", [8, 9, 1000, (1 << 20) + 3, 1 << 24])
@pytest.mark.parametrize("stochastic", [True, False])
def test_roundtrip_random_sizes_bit_exact(n, stochastic):
    x, g = make_outlier_tensor(n, seed=n)
    probs = torch.rand(n, generator=g)
    cfg = SmaqConfig(stochastic_rounding=stochastic)
    ref = smaq_roundtrip(x, cfg, probs=probs)
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    y = cabi.roundtrip(x.to(DEV), ms, cabi.codec_params(cfg), probs=probs.to(DEV))
    assert_bit_equal(y.cpu(), ref.y, f"n={n}")


def test_config0_2p26_roundtrip_matches_oracle():
    """BASELINE configs[0] itself: 64 Mi elements, N(0,1) with 1 % x10, reference defaults — the fused round trip
    against the CPU oracle bit for bit (explicit probs, the oracle's statistics), the kernel's own statistics
    within 1e-6 of the oracle's, and the plugin's one-call path given the kernel's statistics."""
    n = 1 << 26
    x, g = make_outlier_tensor(n, seed=1234)
    probs = torch.rand(n, generator=g)
    cfg = SmaqConfig()
    ref = smaq_roundtrip(x, cfg, probs=probs)
    xd, pd = x.to(DEV), probs.to(DEV)
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    y = cabi.roundtrip(xd, ms, cabi.codec_params(cfg), probs=pd)
    assert_bit_equal(y.cpu(), ref.y, "2^26 round trip")
    del y
    yc, ms_k = cabi.compress(xd, cabi.codec_params(cfg), probs=pd)
    msc = ms_k.cpu()
    assert rel(msc[0].item(), ref.mean.item(), ref.std.item()) < 1e-6 and rel(msc[1].item(), ref.std.item()) < 1e-6
    ref_k = smaq_roundtrip(x, cfg, probs=probs, mean=msc[0], std=msc[1])
    assert_bit_equal(yc.cpu(), ref_k.y, "2^26 smaq_compress given its statistics")


@pytest.mark.parametrize("n", [8, 9, 1000, 32768, 40001, (1 << 20) + 3, 1 << 24])
def test_roundtrip_with_in_kernel_random_numbers_matches_oracle(n):
    """Performance path (no probs tensor): the kernel draws 16 random bits per element (Philox4x32-7, one call per
    16 elements).  oracle/rng.py restates the generator, so the output is checked bit for bit against the CPU oracle
    under the kernels' rounding rule — the large kernel, its unaligned form, the one-block kernel, with and without
    saturation, and a stream offset beyond 32 bits."""
    from oracle import rng as orng

    x, _ = make_outlier_tensor(n, seed=n + 1)
    cfg = SmaqConfig()
    seed, offset = 4242 + n, (3 << 33) + 17
    probs = torch.from_numpy(orng.probs_for(n, seed=seed, offset=offset))
    ref = smaq_roundtrip(x, cfg, probs=probs, rng_rule=True)
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    xd = x.to(DEV)
    y = cabi.roundtrip(xd, ms, cabi.codec_params(cfg, seed=seed, offset=offset))
    assert_bit_equal(y.cpu(), ref.y, f"n={n}")
    sat = smaq_roundtrip(x, cfg, probs=probs, rng_rule=True, saturate=True, all_positive=True)
    y = cabi.roundtrip(xd, ms, cabi.codec_params(cfg, seed=seed, offset=offset, saturate=True, all_positive=True))
    assert_bit_equal(y.cpu(), sat.y, f"n={n} saturate, all_positive")
    buf = torch.empty(n + 1, device=DEV)   # views off the 32-byte grid take the element-wise kernel
    buf[1:].copy_(xd)
    y = cabi.roundtrip(buf[1:], ms, cabi.codec_params(cfg, seed=seed, offset=offset))
    assert_bit_equal(y.cpu(), ref.y, f"n={n} unaligned")
    if n <= 32768:
        y, ms_k = cabi.roundtrip_small(xd, cabi.codec_params(cfg, seed=seed, offset=offset), want_stats=True)
        msc = ms_k.cpu()
        ref_k = smaq_roundtrip(x, cfg, probs=probs, mean=msc[0], std=msc[1], rng_rule=True)
        assert_bit_equal(y.cpu(), ref_k.y, f"n={n} one-block kernel")
    # the rule differs from the reference's expression only on ties |frac - p| <= 2^-25
    lit = smaq_roundtrip(x, cfg, probs=probs)
    assert float((lit.y != ref.y).float().mean()) < 1e-5


def test_bn_kernel_with_in_kernel_random_numbers_matches_oracle():
    from oracle import rng as orng

    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 7, 5, 9, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(7, generator=g) + 0.5, torch.randn(7, generator=g) * 0.2
    cfg = SmaqConfig(use_batch_norm=True)
    probs = torch.from_numpy(orng.probs_for(x.numel(), seed=99, offset=2)).view(x.shape)
    ref = smaq_roundtrip(x, cfg, probs=probs, rng_rule=True, batch_norm_stats=(gamma, beta), all_positive=True)
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    y = cabi.roundtrip_bn(x.to(DEV).view(-1), ms, cabi.codec_params(cfg, seed=99, offset=2, all_positive=True),
                          gamma.to(DEV), beta.to(DEV), 7, 45)
    assert_bit_equal(y.cpu().view(x.shape), ref.y, "BN kernel, in-kernel uniforms")


def test_roundtrip_saturate_matches_oracle():
    x, g = make_outlier_tensor(1 << 18, seed=5)
    probs = torch.rand(x.numel(), generator=g)
    cfg = SmaqConfig()
    ref = smaq_roundtrip(x, cfg, probs=probs, saturate=True)
    assert ref.code.abs().max() <= 63
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    y = cabi.roundtrip(x.to(DEV), ms, cabi.codec_params(cfg, saturate=True), probs=probs.to(DEV))
    assert_bit_equal(y.cpu(), ref.y, "saturate")


@pytest.mark.parametrize("kind", ["zero_mean_neg", "huge_std", "tiny_std", "inf_input", "zeros"])
def test_roundtrip_degenerate_statistics(kind):
    """Tensors for which the three-instruction division is not valid take the IEEE-divide branch."""
    g = torch.Generator().manual_seed(3)
    n = 5000
    x = torch.randn(n, generator=g)
    probs = torch.rand(n, generator=g)
    cfg = SmaqConfig()
    mean, std = None, None
    if kind == "zero_mean_neg":
        mean, std = torch.tensor(-0.0), torch.tensor(1.0)
        cfg = SmaqConfig(stochastic_rounding=False)
    elif kind == "huge_std":
        x = x * 1e30
    elif kind == "tiny_std":
        x = x * 1e-32
    elif kind == "inf_input":
        x[17] = float("inf")
        mean, std = torch.tensor(0.1), torch.tensor(1.3)
    elif kind == "zeros":
        x = torch.zeros(n)
    ref = smaq_roundtrip(x, cfg, probs=probs, mean=mean, std=std)
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    y = cabi.roundtrip(x.to(DEV), ms, cabi.codec_params(cfg), probs=probs.to(DEV))
    assert_bit_equal(y.cpu(), ref.y, kind)


def test_roundtrip_in_place_and_unaligned():
    x, g = make_outlier_tensor(100003, seed=9)
    probs = torch.rand(x.numel(), generator=g)
    cfg = SmaqConfig()
    ref = smaq_roundtrip(x, cfg, probs=probs)
    ms = cabi.mean_std_tensor(ref.mean, ref.std, DEV)
    params = cabi.codec_params(cfg)
    # in place
    xd = x.to(DEV)
    y = cabi.roundtrip(xd, ms, params, probs=probs.to(DEV), out=xd)
    assert y.data_ptr() == xd.data_ptr()
    assert_bit_equal(xd.cpu(), ref.y, "in place")
    # 4-byte-aligned (not 16) views of x, probs and y
    buf_x = torch.empty(x.numel() + 1, device=DEV)
    buf_p = torch.empty(x.numel() + 3, device=DEV)
    buf_y = torch.empty(x.numel() + 2, device=DEV)
    vx, vp, vy = buf_x[1:], buf_p[3:], buf_y[2:]
    vx.copy_(x)
    vp.copy_(probs)
    cabi.roundtrip(vx, ms, params, probs=vp, out=vy)
    assert_bit_equal(vy.cpu(), ref.y, "unaligned")


# ---- statistics --------------------------------------------------------------------------------
def rel(a, b, scale=None):
    scale = abs(b) if scale is None else scale
    return abs(a - b) / scale


@pytest.mark.parametrize("n", [8, 1000, 4099, (1 << 20) + 3, 1 << 24, (1 << 26) + 5])
@pytest.mark.parametrize("kind", ["outliers", "shifted"])
def test_full_statistics_within_1e6(n, kind):
    x, _ = make_outlier_tensor(n, seed=n + 1)
    if kind == "shifted":
        x = x * 0.02 + 3.0
    xd = x.to(DEV)
    got = cabi.stats_full(xd).cpu()
    xd64 = x.double()
    true_mean, true_std = xd64.mean().item(), xd64.std().item()
    ref_mean, ref_std = full_mean_std(x, SmaqConfig())
    # vs fp64 ground truth, and vs what the reference's fp32 torch ops return
    assert rel(got[0].item(), true_mean, max(abs(true_mean), true_std)) < 1e-6
    assert rel(got[1].item(), true_std) < 1e-6
    assert rel(got[0].item(), ref_mean.item(), max(abs(ref_mean.item()), ref_std.item())) < 1e-6
    assert rel(got[1].item(), ref_std.item()) < 1e-6
    # run-to-run reproducible (fixed combination order)
    again = cabi.stats_full(xd).cpu()
    assert torch.equal(got.view(torch.int32), again.view(torch.int32))


def test_statistics_edge_cases():
    xd = torch.full((1000,), 0.75, device=DEV)
    got = cabi.stats_full(xd).cpu()
    assert got[0].item() == 0.75 and got[1].item() == 0.0
    x = torch.randn(100)
    x[50] = float("nan")
    got = cabi.stats_full(x.to(DEV)).cpu()
    assert math.isnan(got[0].item()) and math.isnan(got[1].item())
    # unaligned view
    buf = torch.randn(1001, device=DEV)
    got = cabi.stats_full(buf[1:]).cpu()
    ref = buf[1:].double()
    assert rel(got[1].item(), ref.std().item()) < 1e-6
    # biased variant
    got = cabi.stats_full(buf[1:], unbiased=False).cpu()
    assert rel(got[1].item(), ref.std(unbiased=False).item()) < 1e-6


def test_sampled_statistics_explicit_indices():
    x, g = make_outlier_tensor(1 << 20, seed=77)
    for k in (16, 64, 1024):
        idx = torch.randperm(x.numel(), generator=g)[:k]
        cfg = SmaqConfig(use_sample_stats=True, num_samples=k)
        ref_mean, ref_std = sample_mean_std(x, idx, cfg)
        got = cabi.stats_sampled(x.to(DEV), idx).cpu()
        assert rel(got[0].item(), ref_mean.item(), max(abs(ref_mean.item()), ref_std.item())) < 1e-6
        assert rel(got[1].item(), ref_std.item()) < 1e-6


def test_sampled_statistics_device_draw():
    """Device-drawn indices: a uniform k-subset (the law of randperm(n)[:k]); deterministic per seed."""
    n, k = 1 << 16, 16
    x = torch.arange(n, dtype=torch.float32)  # value == index: statistics reveal the draw
    xd = x.to(DEV)
    a = cabi.stats_sampled_draw(xd, k, seed=1).cpu()
    b = cabi.stats_sampled_draw(xd, k, seed=1).cpu()
    c = cabi.stats_sampled_draw(xd, k, seed=2).cpu()
    assert torch.equal(a, b) and not torch.equal(a, c)
    means = torch.stack([cabi.stats_sampled_draw(xd, k, seed=s, offset=3) for s in range(400)]).cpu()[:, 0]
    # E[mean] = (n-1)/2, sd of the mean of 16 uniform draws = n/sqrt(12*16); 400 repetitions
    assert abs(means.mean().item() - (n - 1) / 2) < 4 * n / math.sqrt(12 * 16 * 400)
    # k == n must select every index exactly once
    small = torch.arange(16, dtype=torch.float32, device=DEV)
    got = cabi.stats_sampled_draw(small, 16, seed=5).cpu()
    assert abs(got[0].item() - 7.5) < 1e-6 and rel(got[1].item(), small.double().std(unbiased=False).item()) < 1e-6


def test_range_statistics():
    x, _ = make_outlier_tensor(300001, seed=4)
    cfg = SmaqConfig(use_range_std_dev=True)
    ref_std = std_of(x, cfg)
    got = cabi.stats_range(x.to(DEV)).cpu()
    assert rel(got[0].item(), x.double().mean().item(), 1.0) < 1e-6
    assert rel(got[1].item(), ref_std.item()) < 1e-6


# ---- small-tensor fused kernel --------------------------------------------------------------------
@pytest.mark.parametrize("n", [8, 10, 512, 4099, 32768])
def test_small_fused_matches_oracle_given_its_own_stats(n):
    x, g = make_outlier_tensor(n, seed=n + 3)
    probs = torch.rand(n, generator=g)
    cfg = SmaqConfig()
    y, ms = cabi.roundtrip_small(x.to(DEV), cabi.codec_params(cfg), probs=probs.to(DEV), want_stats=True)
    ms = ms.cpu()
    ref_mean, ref_std = full_mean_std(x, cfg)
    assert rel(ms[0].item(), ref_mean.item(), max(abs(ref_mean.item()), ref_std.item())) < 1e-6
    assert rel(ms[1].item(), ref_std.item()) < 1e-6
    ref = smaq_roundtrip(x, cfg, probs=probs, mean=ms[0], std=ms[1])
    assert_bit_equal(y.cpu(), ref.y, f"small n={n}")


# ---- plugin end to end ---------------------------------------------------------------------------
def make_plugin(argv=(), precision=32):
    from argparse import ArgumentParser

    from smart_compress.compress.smart import SmartFP

    args = SmartFP.add_argparse_args(ArgumentParser()).parse_args(list(argv))
    args.precision = precision
    return SmartFP(args)


@pytest.mark.parametrize("name", sorted(CASES))
def test_plugin_golden_cases(name):
    """Through SmartFP.__call__: the kernel's own statistics, the golden's probs / indices.
    Bit-exact against the oracle fed the kernel's statistics; the statistics within 1e-6."""
    c = CASES[name]
    fp = make_plugin(c["argv"].split(), c["precision"])
    xd = c["x"].to(DEV)
    kwargs = dict(c["kwargs"])
    oracle_kwargs = dict(c["kwargs"])
    if "batch_norm_stats" in kwargs:  # the layer's gamma / beta live on the device, as the hook passes them
        kwargs["batch_norm_stats"] = tuple(t.to(DEV) for t in kwargs["batch_norm_stats"])
        if c["cfg"].bn_scalar_params and c["cfg"].use_batch_norm:
            # the two means are formed on the device (as the reference would on CUDA); the oracle gets those
            g, b = kwargs["batch_norm_stats"]
            oracle_kwargs["batch_norm_stats"] = (g.mean().cpu().expand(g.numel()).clone(), b.mean().cpu().expand(b.numel()).clone())
    y = fp(xd, tag="t", _probs=c["probs"], _sample_idx=c["idx"], **kwargs)
    if c["same_object"]:
        assert y is xd
        return
    assert y is not xd and y.shape == xd.shape and y.dtype == xd.dtype and y.device == xd.device
    cfg = c["cfg"]
    flat = xd.contiguous().view(-1)
    ms = (fp.statistics(flat, c["idx"]) if (cfg.use_sample_stats or cfg.use_range_std_dev or uses_bn(c) or
                                              flat.numel() > 32768) else None)
    if ms is None:
        _, ms = cabi.roundtrip_small(flat, cabi.codec_params(cfg), probs=None if c["probs"] is None
                                     else c["probs"].to(DEV).view(-1), want_stats=True)
    ms = ms.cpu()
    if oracle_kwargs is not c["kwargs"] and c["cfg"].bn_scalar_params and uses_bn(c):
        import dataclasses
        cfg = dataclasses.replace(cfg, bn_scalar_params=False)  # the per-channel arrays already hold the device's means
    ref = smaq_roundtrip(c["x"].clone(), cfg, probs=c["probs"], idx=c["idx"], mean=ms[0], std=ms[1], **oracle_kwargs)
    assert_bit_equal(y.cpu(), ref.y, name)
    if not torch.isnan(c["y"]).any():
        # and close to the reference's own output (statistics may differ in the last bits)
        ref_stats = smaq_roundtrip(c["x"].clone(), cfg, probs=c["probs"], idx=c["idx"], **c["kwargs"])
        scale = max(abs(ref_stats.mean.item()), ref_stats.std.item(), 1e-30)
        assert rel(ms[0].item(), ref_stats.mean.item(), scale) < 1e-6
        if ref_stats.std.item() != 0:
            assert rel(ms[1].item(), ref_stats.std.item()) < 1e-6


def test_plugin_philox_path_properties():
    """Performance path (in-kernel Philox): seeded, unbiased, codes on the reference's grid."""
    fp = make_plugin()
    x, _ = make_outlier_tensor(1 << 22, seed=11)
    xd = x.to(DEV)
    torch.manual_seed(123)
    fp1 = make_plugin()
    y1 = fp1(xd)
    torch.manual_seed(123)
    fp2 = make_plugin()
    y2 = fp2(xd)
    assert torch.equal(y1, y2)          # same seed, same call index -> same stream
    y3 = fp2(xd)
    assert not torch.equal(y1, y3)      # next call -> next stream
    # every output must be one of the two values the oracle can produce (probs = 0+ and probs -> 1)
    ms = fp.statistics(xd).cpu()
    cfg = SmaqConfig()
    lo = smaq_roundtrip(x, cfg, probs=torch.full_like(x, 1.0 - 2 ** -24), mean=ms[0], std=ms[1]).y
    hi = smaq_roundtrip(x, cfg, probs=torch.zeros_like(x), mean=ms[0], std=ms[1]).y
    y = y1.cpu()
    assert bool(((y == lo) | (y == hi)).all())
    # stochastic rounding is unbiased: mean error is ~N(0, step^2/ (6 n))
    err = (y.double() - x.double()).mean().item()
    assert abs(err) < 5 * ms[1].item() / 15 / math.sqrt(x.numel())


def test_plugin_noncontiguous_and_shapes():
    fp = make_plugin(["--no_stochastic_rounding"])
    x = torch.randn(6, 5, 7, 3, device=DEV)
    xt = x.permute(0, 3, 1, 2)  # non-contiguous view, as cuDNN channels-last outputs can be
    y = fp(xt)
    assert y.shape == xt.shape and y.stride() == xt.stride()   # a dense view keeps its layout (as the reference's chain does)
    assert torch.equal(y, fp(x).permute(0, 3, 1, 2))            # processed in storage order: the contiguous call's bits
    ref = fp(xt.contiguous())                                   # logical order: the statistics may differ in the last bit
    step = float(x.std()) / 15
    assert float((y - ref).abs().max()) <= 1.01 * step and float(((y - ref).abs() > 1e-6).float().mean()) < 1e-3


def test_plugin_refuses_cpu_tensors():
    from smart_compress._native import NativeLibraryError

    fp = make_plugin()
    with pytest.raises(NativeLibraryError):
        fp(torch.randn(100))


def test_compression_ratio_logging():
    fp = make_plugin(["--measure_compression_ratio", "--no_stochastic_rounding"])
    seen = {}
    fp.log = lambda k, v, **kw: seen.__setitem__(k, v)
    x, _ = make_outlier_tensor(1 << 16, seed=21)
    fp(x.to(DEV), tag="forward_autograd")
    cfg = SmaqConfig(stochastic_rounding=False)
    ref = smaq_roundtrip(x, cfg)
    n_out = int((ref.hi | ref.lo).sum())
    want = n_out * 8 + (x.numel() - n_out) * 6
    assert abs(seen["new_size_forward_autograd"] - want) <= 8 * 4  # stats differ in the last bit at most
    assert seen["orig_size"] == x.numel() * 32
    assert 4.0 < seen["compression_ratio"] < 5.4


@pytest.mark.parametrize("scale", [1e-30, 1e-22, 1e-12, 1e12, 1e25])
@pytest.mark.parametrize("n", [600, 32768, 100000, (1 << 20) + 3])
def test_statistics_of_tiny_and_huge_magnitudes(n, scale):
    """The chunked pass squares deviations in fp32; where those squares would underflow (|x| ~ 1e-30: gradients
    late in training) or overflow it must fall back to fp64 — torch's CPU std accumulates in double."""
    g = torch.Generator().manual_seed(n)
    x = (torch.randn(n, generator=g) * scale + 0.25 * scale)
    xd = x.to(DEV)
    want_mean, want_std = x.double().mean().item(), x.double().std().item()
    if n > 32768:
        ms = cabi.stats_full(xd).cpu()
    else:
        _, ms = cabi.roundtrip_small(xd, cabi.codec_params(SmaqConfig(stochastic_rounding=False)), want_stats=True)
        ms = ms.cpu()
    assert rel(ms[0].item(), want_mean, max(abs(want_mean), want_std)) < 1e-6
    assert rel(ms[1].item(), want_std) < 1e-6
