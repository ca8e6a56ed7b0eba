"""Edge cases of the hot paths on a real device: special values inside otherwise ordinary tensors (they take
the per-chunk detours of the packed encoder), exact-zero probabilities, the saturation counter switch, the
fused statistics+round-trip entry point, workspace reuse across sizes, and concurrent callers."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

from oracle import pack as opack
from oracle.smaq import SmaqConfig, smaq_roundtrip
from smart_compress import _native as N
from tests import cabi, cabi_pack
from tests.golden_util import assert_bit_equal
from tests.test_gpu_smaq import make_outlier_tensor, make_plugin

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def encode_and_compare(x, cfg, probs, mean, std):
    res = smaq_roundtrip(x, cfg, probs=probs, mean=mean, std=std)
    ms = cabi.mean_std_tensor(res.mean, res.std, DEV)
    pd = None if probs is None else probs.to(DEV)
    buf, lay = cabi_pack.encode(x.to(DEV), ms, cabi.codec_params(cfg), cfg, probs=pd)
    p = opack.pack(res, cfg)
    hdr = cabi_pack.assert_stream_equals_oracle(buf, lay, p)
    assert hdr.n_saturated == p.n_saturated
    y = cabi_pack.decode(buf, lay)
    assert_bit_equal(y.cpu(), opack.decode(p), "decode vs oracle decode")
    return hdr


@pytest.mark.parametrize("stochastic", [True, False])
def test_special_values_inside_full_tiles(stochastic):
    """NaN, +-inf, x == mean (z = 0), -0.0, denormals and huge values scattered through aligned, full warp tiles:
    each one sends ONE 8-element chunk of the straight-line path to the IEEE-division re-run; everything must
    still be byte-identical to the oracle."""
    g = torch.Generator().manual_seed(21)
    n = 64 * 1024
    x = torch.randn(n, generator=g)
    mean, std = torch.tensor(0.25), torch.tensor(1.5)
    specials = [float("nan"), float("inf"), float("-inf"), 0.25, -0.0, 0.0, 1e-42, -1e-42, 3e38, -3e38, 0.25 + 1e-12,
                0.25 - 1.5 * 2.0 ** -41, 1.75, -1.25, 4.0, -3.5]   # incl. |z| == 1 and |z| == 2.5 exactly
    pos = torch.randperm(n, generator=g)[: 4 * len(specials)]
    for i, p in enumerate(pos):
        x[p] = specials[i % len(specials)]
    probs = torch.rand(n, generator=g) if stochastic else None
    cfg = SmaqConfig(stochastic_rounding=stochastic)
    hdr = encode_and_compare(x, cfg, probs, mean, std)
    assert hdr.n_saturated >= 4 * 5  # nan, +-inf, +-3e38


def test_zero_probability_rounds_like_the_reference():
    """p == 0 with a fractional part that rounds to 1.0 is the one case where rint(relu((frac - p) + 0.5)) is 2
    (SURVEY.md §7.3 H2): the explicit-probs path must reproduce it on the straight-line path."""
    n = 8192
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, generator=g)
    probs = torch.rand(n, generator=g)
    mean, std = torch.tensor(0.0), torch.tensor(1.0)
    x[100] = -2.0 ** -30 / 15.0     # c = -2^-30: floor -1, frac rounds to 1.0
    probs[100] = 0.0
    x[4000] = -1e-6
    probs[4000] = 0.0
    cfg = SmaqConfig()
    res = smaq_roundtrip(x, cfg, probs=probs, mean=mean, std=std)
    assert res.code[100] == 1.0     # the reference's own quirk
    encode_and_compare(x, cfg, probs, mean, std)
    ms = cabi.mean_std_tensor(mean, std, DEV)
    y = cabi.roundtrip(x.to(DEV), ms, cabi.codec_params(cfg), probs=probs.to(DEV))
    assert_bit_equal(y.cpu(), res.y, "round trip with zero probabilities")


def test_saturation_counter_is_optional():
    x, g = make_outlier_tensor(100000, seed=9)
    xd = x.to(DEV)
    cfg = SmaqConfig()
    ms = cabi.stats_full(xd)
    on, lay = cabi_pack.encode(xd, ms, cabi.codec_params(cfg, seed=4, count_saturated=True), cfg)
    off, _ = cabi_pack.encode(xd, ms, cabi.codec_params(cfg, seed=4, count_saturated=False), cfg)
    h_on, h_off = cabi_pack.sections(on, lay)[0], cabi_pack.sections(off, lay)[0]
    assert h_off.n_saturated == 2 ** 64 - 1 and 0 < h_on.n_saturated < x.numel()
    used = lay.extras_off + 4 * h_on.extras_words
    assert torch.equal(on[128:used], off[128:used])   # the stream itself does not depend on the switch
    msc = ms.cpu()
    z = (x - msc[0]) / msc[1]
    c = torch.where(z.abs() > 1, (z - torch.sign(z)) * cfg.range_outlier, z * cfg.range_normal)
    assert abs(int((c.abs() > 63).sum()) - h_on.n_saturated) <= 2   # fp32 vs this fp32-ish recomputation


# smaq_compress: the statistics kernel, then the round trip as a programmatic dependent launch that walks the
# tensor back to front (saturating calls and pointers that are not 32-byte aligned: an ordinary launch).  Same
# statistics kernel, element-indexed random numbers: the result must equal smaq_stats_full -> smaq_roundtrip bit
# for bit at every size, across the grid-sizing thresholds and with ragged tails.
FUSED_SIZES = [32769, 40000, 40003, (1 << 20) + 5, 909312, 909313 + 8, (1 << 22) + 1, 3 * (1 << 22) + 7]


def _check_compress(xd, params, probs=None):
    y, ms = cabi.compress(xd, params, probs=probs)
    ref = cabi.stats_full(xd)
    assert torch.equal(ms.view(torch.int32), ref.view(torch.int32)), (ms, ref)
    want = cabi.roundtrip(xd, ms, params, probs=probs)
    assert torch.equal(y.view(torch.int32), want.view(torch.int32))
    return y, ms


@pytest.mark.parametrize("n", FUSED_SIZES)
def test_fused_entry_point_equals_stats_then_roundtrip(n):
    x, _ = make_outlier_tensor(n, seed=n)
    xd = x.to(DEV)
    lib = N.load()
    params = cabi.codec_params(SmaqConfig(), seed=99, offset=3)
    y, ms = _check_compress(xd, params)
    # reused scratch, twice: same bits, and in place
    need = lib.smaq_compress_workspace_bytes(n)
    ws = torch.full((need + 4096,), 0xA5, dtype=torch.uint8, device=DEV)   # larger than needed, dirty
    N.check(lib.smaq_compress_workspace_init(ws.data_ptr(), ws.numel(), N.stream_ptr(xd.device)), "init")  # once
    for _ in range(2):
        y2, ms2 = cabi.compress(xd, params, ws=ws)
        assert torch.equal(ms2.view(torch.int32), ms.view(torch.int32))
        assert torch.equal(y2.view(torch.int32), y.view(torch.int32))
    inplace = xd.clone()
    cabi.compress(inplace, params, out=inplace, ws=ws)
    assert torch.equal(inplace.view(torch.int32), y.view(torch.int32))
    small = torch.empty(16, dtype=torch.uint8, device=DEV)
    assert lib.smaq_compress(xd.data_ptr(), y.data_ptr(), n, None, C.byref(params), small.data_ptr(), 16,
                             N.stream_ptr(xd.device)) == 3          # SMAQ_ERR_WORKSPACE
    assert b"workspace" in lib.smaq_b200_last_error()


@pytest.mark.parametrize("stochastic", [False, True])
@pytest.mark.parametrize("all_positive", [False, True])
def test_fused_entry_point_variants(stochastic, all_positive):
    n = (1 << 20) + 3
    x, _ = make_outlier_tensor(n, seed=5)
    xd = x.to(DEV)
    cfg = SmaqConfig(stochastic_rounding=stochastic)
    params = cabi.codec_params(cfg, seed=7, offset=11, all_positive=all_positive)
    _check_compress(xd, params)
    if stochastic:   # explicit uniforms (parity mode) through the same launches
        probs = torch.rand(n, generator=torch.Generator().manual_seed(3)).to(DEV)
        _check_compress(xd, params, probs=probs)
    # saturating calls and views that are not 32-byte aligned take the ordinary launch: same contract
    _check_compress(xd, cabi.codec_params(cfg, seed=7, offset=11, all_positive=all_positive, saturate=True))
    _check_compress(xd[1:], params)


def test_fused_entry_point_beyond_the_dependent_launch_size():
    """Above 2^27 elements smaq_compress takes two ordinary launches (measured faster there): same contract."""
    n = (1 << 27) + 8
    xd = torch.randn(n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))
    _check_compress(xd, cabi.codec_params(SmaqConfig(), seed=5, offset=1))


def test_fused_entry_point_against_the_oracle():
    """The whole call against the CPU oracle under the statistics it published, with explicit uniforms."""
    n = 70000
    x, _ = make_outlier_tensor(n, seed=17)
    probs = torch.rand(n, generator=torch.Generator().manual_seed(4))
    cfg = SmaqConfig()
    y, ms = cabi.compress(x.to(DEV), cabi.codec_params(cfg), probs=probs.to(DEV))
    msc = ms.cpu()
    want = smaq_roundtrip(x, cfg, probs=probs, mean=msc[0], std=msc[1])
    assert_bit_equal(y.cpu(), want.y, "smaq_compress vs oracle")


def test_fused_entry_point_special_tensors():
    params = cabi.codec_params(SmaqConfig(), seed=1)
    n = 1 << 18
    const = torch.full((n,), 3.25, device=DEV)                       # std 0 -> 1, codes 0 (smart.py:151-152)
    y, ms = cabi.compress(const, params)
    assert float(ms[1]) == 0.0 and torch.equal(y, const)
    bad = torch.randn(n, device=DEV)
    bad[n // 3] = float("nan")                                       # any NaN -> statistics NaN -> all NaN
    y, ms = cabi.compress(bad, params)
    assert torch.isnan(ms).all() and torch.isnan(y).all()
    bad[n // 3] = float("inf")
    y, ms = cabi.compress(bad, params)
    assert torch.isnan(y).all()


def test_plugin_scratch_reuse_across_sizes_and_threads():
    """Forward calls come from the main thread, backward calls from autograd's worker: two threads, each on
    its own stream, sharing one plugin instance, with tensor sizes that grow and shrink."""
    fp = make_plugin(["--no_stochastic_rounding"])   # deterministic: results comparable across threads
    sizes = [50000, 300000, 70000, 1 << 20, 33000]
    xs = [torch.randn(s, device=DEV) for s in sizes]
    want = [fp(x, tag="t").clone() for x in xs]
    torch.cuda.synchronize()
    errors = []

    def worker():
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(20):
                    for x, w in zip(xs, want):
                        y = fp(x, tag="t")
                        if not torch.equal(y, w):
                            errors.append("mismatch")
            st.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker) for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


@pytest.mark.parametrize("t_main,t_out", [(1.5, 3.0), (0.75, 2.0), (2.0, 5.0), (0.5, 1.25), (1.0, 4.0), (3.0, 3.5)])
@pytest.mark.parametrize("stochastic", [True, False])
def test_other_thresholds_take_the_right_path(t_main, t_out, stochastic):
    """The straight-line paths need a power-of-two main threshold (then (z -+ t) * range == fma(z, range, -+K)
    exactly); any other threshold must fall back to the generic sequence — and both must stay bit-exact, in the
    fused round trip and in the packed stream."""
    n = 40000 + 1024 * 8
    x, g = make_outlier_tensor(n, seed=int(100 * t_main + t_out))
    probs = torch.rand(n, generator=g) if stochastic else None
    cfg = SmaqConfig(main_std_dev_threshold=t_main, outlier_std_dev_threshold=t_out, stochastic_rounding=stochastic)
    res = smaq_roundtrip(x, cfg, probs=probs)
    ms = cabi.mean_std_tensor(res.mean, res.std, DEV)
    pd = None if probs is None else probs.to(DEV)
    y = cabi.roundtrip(x.to(DEV), ms, cabi.codec_params(cfg), probs=pd)
    assert_bit_equal(y.cpu(), res.y, "round trip")
    ysat = cabi.roundtrip(x.to(DEV), ms, cabi.codec_params(cfg, saturate=True), probs=pd)
    assert_bit_equal(ysat.cpu(), smaq_roundtrip(x, cfg, probs=probs, saturate=True).y, "saturated round trip")
    encode_and_compare(x, cfg, probs, res.mean, res.std)


def test_dependent_launches_inside_a_cuda_graph():
    """The two-kernel calls are programmatic dependent launches; captured into a CUDA graph (programmatic edges)
    and replayed they must give the bits of the eager calls — statistics + round trip, the packed encoder's two
    passes + decode, and S2FP8's statistics + apply."""
    n = (1 << 20) + 24
    x, _ = make_outlier_tensor(n, seed=31)
    xd = x.to(DEV)
    cfg = SmaqConfig()
    lib = N.load()
    params = cabi.codec_params(cfg, seed=77, offset=5)
    want_y, want_ms = cabi.compress(xd, params)
    ms = cabi.stats_full(xd)
    want_buf, lay = cabi_pack.encode(xd, ms, params, cfg)
    want_dec = cabi_pack.decode(want_buf, lay)
    p8 = cabi.floatq_params(5, 2, seed=9, offset=2)
    want_s2 = cabi.s2fp8_apply(xd, cabi.s2fp8_stats(xd), p8)

    y = torch.empty_like(xd)
    dec = torch.empty_like(xd)
    s2 = torch.empty_like(xd)
    mm = torch.empty(2, dtype=torch.float32, device=DEV)
    need = lib.smaq_compress_workspace_bytes(n)
    ws = torch.zeros(need, dtype=torch.uint8, device=DEV)
    sws_b = lib.smaq_stats_workspace_bytes(n)
    sws = torch.zeros(sws_b, dtype=torch.uint8, device=DEV)
    buf = torch.zeros(lay.total_capacity_bytes, dtype=torch.uint8, device=DEV)
    pws = torch.zeros(lay.workspace_bytes, dtype=torch.uint8, device=DEV)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            st = N.stream_ptr(xd.device)
            N.check(lib.smaq_compress(xd.data_ptr(), y.data_ptr(), n, None, C.byref(params), ws.data_ptr(), ws.numel(),
                                      st), "compress")
            N.check(lib.smaq_encode(xd.data_ptr(), n, ms.data_ptr(), None, C.byref(params), buf.data_ptr(), buf.numel(),
                                    pws.data_ptr(), pws.numel(), st), "encode")
            N.check(lib.smaq_decode(buf.data_ptr(), buf.numel(), n, cfg.num_bits_main, cfg.num_bits_outlier, 0,
                                    dec.data_ptr(), st), "decode")
            N.check(lib.smaq_s2fp8_stats(xd.data_ptr(), n, mm.data_ptr(), sws.data_ptr(), sws_b, st), "s2fp8_stats")
            N.check(lib.smaq_s2fp8_apply(xd.data_ptr(), s2.data_ptr(), n, mm.data_ptr(), None, C.byref(p8), st),
                    "s2fp8_apply")
    for _ in range(3):
        y.zero_(), dec.zero_(), s2.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(y.view(torch.int32), want_y.view(torch.int32))
        assert torch.equal(dec.view(torch.int32), want_dec.view(torch.int32))
        assert_bit_equal(s2, want_s2, "S2FP8 in a graph")
