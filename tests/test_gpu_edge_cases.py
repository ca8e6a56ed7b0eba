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


def test_dense_layouts_keep_their_strides():
    """a channels_last activation / weight must come back channels_last (the reference's elementwise chain
    preserves the memory format; `.contiguous()` would add a transpose copy and hand NCHW tensors to a channels_last
    network).  Processed in storage order: the result is the contiguous call's on the permuted tensor, bit for bit."""
    from argparse import ArgumentParser

    from smart_compress.compress.fp8 import FP8
    from smart_compress.compress.s2fp8 import S2FP8

    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 32, 24, 24, generator=g).to(DEV).to(memory_format=torch.channels_last)
    assert not x.is_contiguous()
    torch.manual_seed(4)
    y = make_plugin()(x, tag="forward_autograd")
    assert y.stride() == x.stride() and y.is_contiguous(memory_format=torch.channels_last)
    torch.manual_seed(4)
    flat = x.permute(0, 2, 3, 1).contiguous()          # the same elements in storage order
    want = make_plugin()(flat, tag="forward_autograd")
    assert torch.equal(y.permute(0, 2, 3, 1), want)
    # the batched optimizer path takes such tensors in place
    w = torch.randn(64, 16, 3, 3, generator=g).to(DEV).to(memory_format=torch.channels_last)
    before = w.clone()
    out = make_plugin().compress_many([w], None, tag="optimizer_weight")
    assert out[0] is w and w.stride() == before.stride() and not torch.equal(w, before)
    assert float((w - before).abs().max()) < float(before.std()) / 10
    for cls in (FP8, S2FP8):
        hp = cls.add_argparse_args(ArgumentParser()).parse_args([])
        hp.precision = 32
        z = cls(hp)(x, tag="t")
        assert z.stride() == x.stride() and bool(torch.isfinite(z).all())
    # a strided view that is NOT dense still works (copied)
    v = torch.randn(64, 200, generator=g).to(DEV)[:, ::2]
    assert make_plugin()(v, tag="t").shape == v.shape


def test_hparams_edits_act_as_in_the_reference():
    """the reference derives its ranges once, in __init__ (smart.py:75-84), and reads the threshold and the
    widths from hparams on every call.  An in-place edit of the threshold must therefore take effect on the next call
    (it silently did not in round 1), with the ranges unchanged — bit for bit the oracle's result for that mix."""
    import dataclasses

    fp = make_plugin(["--no_stochastic_rounding"])
    x, _ = make_outlier_tensor(50000, seed=2)
    xd = x.to(DEV)
    fp(xd, tag="t")
    fp.hparams.main_std_dev_threshold = 1.25
    y = fp(xd, tag="t")
    ms = fp.statistics(xd).cpu()
    base = SmaqConfig(stochastic_rounding=False)

    @dataclasses.dataclass
    class Mixed(SmaqConfig):   # threshold re-read, ranges as derived at construction
        @property
        def range_normal(self):
            return base.range_normal

        @property
        def range_outlier(self):
            return base.range_outlier

    ref = smaq_roundtrip(x, Mixed(stochastic_rounding=False, main_std_dev_threshold=1.25), mean=ms[0], std=ms[1])
    assert_bit_equal(y.cpu(), ref.y, "threshold edited in place")


def test_beyond_two_to_the_31_elements():
    """The header promises 64-bit element counts: 2^31 + 2^20 + 5 elements (8.6 GB) through the statistics, the fused
    round trip and the packed encode -> decode; the CPU oracle on slices from the start, the 2^31 boundary and the
    ragged end (the kernels' own random numbers via oracle/rng.py with 64-bit element indices)."""
    from oracle import pack as opack
    from oracle import rng as orng

    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~35 GB of device memory")
    n = (1 << 31) + (1 << 20) + 5
    g = torch.Generator(device=DEV).manual_seed(7)
    xd = torch.empty(n, device=DEV)
    for lo in range(0, n, 1 << 28):                      # generated in pieces: randn's own 2^31 limits are not the subject
        hi = min(n, lo + (1 << 28))
        xd[lo:hi] = torch.randn(hi - lo, generator=g, device=DEV) * 1.5 + 0.25
    cfg = SmaqConfig()
    ms = cabi.stats_full(xd)
    msc = ms.cpu()
    mean = sum(float(xd[lo:lo + (1 << 28)].double().sum()) for lo in range(0, n, 1 << 28)) / n
    sq = sum(float(((xd[lo:lo + (1 << 28)].double() - mean) ** 2).sum()) for lo in range(0, n, 1 << 28))
    std = (sq / (n - 1)) ** 0.5
    assert abs(msc[0].item() - mean) <= 1e-6 * std and abs(msc[1].item() - std) <= 1e-6 * std
    params = cabi.codec_params(cfg, seed=3, offset=11, saturate=True)
    y = cabi.roundtrip(xd, ms, params)
    buf, lay = cabi_pack.encode(xd, ms, params, cfg)
    dec = cabi_pack.decode(buf, lay)
    assert torch.equal(dec[-(1 << 22):].view(torch.int32), y[-(1 << 22):].view(torch.int32))
    assert torch.equal(dec[: 1 << 22].view(torch.int32), y[: 1 << 22].view(torch.int32))
    for first, count in ((0, 4096), ((1 << 31) - 2048, 4096), (n - 3000, 3000)):
        xs = xd[first:first + count].cpu()
        probs = torch.from_numpy(orng.probs_for(count, seed=3, offset=11, first=first))
        ref = smaq_roundtrip(xs, cfg, probs=probs, mean=msc[0], std=msc[1], rng_rule=True, saturate=True)
        assert_bit_equal(y[first:first + count].cpu(), ref.y, f"round trip at {first}")
        assert_bit_equal(dec[first:first + count].cpu(), ref.y, f"decode at {first}")
    h = N.PackedHeader.from_buffer_copy(bytes(buf[:128].cpu().numpy())[: C.sizeof(N.PackedHeader)])
    assert h.n == n and h.status == 0 and 0.25 * n < h.n_outlier < 0.40 * n
    del y, dec, buf, xd
    torch.cuda.empty_cache()
