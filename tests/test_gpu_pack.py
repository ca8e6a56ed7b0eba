"""GPU parity tests for the packed SmaQ stream (smaq_encode / smaq_decode).

Contract: given the reference's mean/std and the same uniform numbers, the packed payload (tag
words, base fields, extras stream, tile table, counters) is BYTE-IDENTICAL to oracle/pack.py
applied to the reference's integer codes, and decode() is bit-identical to the reference's
round trip with codes saturated at the field width."""
import numpy as np
import pytest
import torch

from oracle import pack as opack
from oracle.smaq import SmaqConfig, compressed_bits, smaq_roundtrip
from tests import cabi, cabi_pack
from tests.golden_util import assert_bit_equal, load_golden
from tests.test_gpu_smaq import make_outlier_tensor, make_plugin

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = load_golden()


def run_case(x, cfg, probs, *, mean=None, std=None, all_positive=False, idx=None):
    res = smaq_roundtrip(x, cfg, probs=probs, idx=idx, mean=mean, std=std)
    ms = cabi.mean_std_tensor(res.mean, res.std, DEV)
    xd = x.to(DEV).contiguous().view(-1)
    pd = None if probs is None else probs.to(DEV).contiguous().view(-1)
    buf, lay = cabi_pack.encode(xd, ms, cabi.codec_params(cfg), cfg, probs=pd)
    y = cabi_pack.decode(buf, lay, all_positive=all_positive)
    want = smaq_roundtrip(x, cfg, probs=probs, idx=idx, mean=res.mean, std=res.std, saturate=True,
                          all_positive=all_positive)
    p = opack.pack(res, cfg)
    hdr, table, planes, extras = cabi_pack.sections(buf, lay)
    assert hdr.status == 0 and hdr.magic == opack.MAGIC and hdr.n == x.numel()
    assert np.array_equal(planes[:, 0, :], p.planes[:, 0, :]), "outlier bitmap differs"
    assert np.array_equal(planes, p.planes), "base fields differ"
    assert np.array_equal(table, p.table), "tile table differs"
    assert np.array_equal(extras, p.extras), "extras stream differs"
    assert hdr.n_outlier == p.n_outlier and hdr.n_saturated == p.n_saturated
    assert hdr.extras_words == int(p.table[-1])
    nan_free = not torch.isnan(want.y).any()
    if nan_free:
        assert_bit_equal(y.cpu().view(x.shape), want.y, "decode")
    # size contract (smart.py:184-187)
    assert cfg.num_bits_outlier * hdr.n_outlier + cfg.num_bits_main * (hdr.n - hdr.n_outlier) == compressed_bits(res, cfg)
    return buf, lay, p


@pytest.mark.parametrize("n", [8, 33, 1000, 1024, 8191, 8192, 8193, 100003, (1 << 20) + 3])
@pytest.mark.parametrize("stochastic", [True, False])
def test_encode_decode_byte_exact(n, stochastic):
    x, g = make_outlier_tensor(n, seed=n + 17)
    probs = torch.rand(n, generator=g) if stochastic else None
    run_case(x, SmaqConfig(stochastic_rounding=stochastic), probs)


def test_encode_decode_16m():
    x, g = make_outlier_tensor(1 << 24, seed=3)
    run_case(x, SmaqConfig(), torch.rand(x.numel(), generator=g))


@pytest.mark.parametrize("bm,bo", [(4, 6), (5, 9), (6, 6), (8, 12), (4, 4), (7, 8), (6, 9)])
def test_other_bit_widths(bm, bo):
    x, g = make_outlier_tensor(50001, seed=bm * 31 + bo)
    cfg = SmaqConfig(num_bits_main=bm, num_bits_outlier=bo)
    run_case(x, cfg, torch.rand(x.numel(), generator=g))


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items()
                                        if not c["same_object"] and 4 <= c["cfg"].num_bits_main <= 8
                                        and 0 <= c["cfg"].num_bits_outlier - c["cfg"].num_bits_main <= 4))
def test_golden_inputs_through_the_packed_path(name):
    c = CASES[name]
    run_case(c["x"], c["cfg"], c["probs"], idx=c["idx"], all_positive=c["kwargs"].get("all_positive", False))


def test_degenerate_statistics_and_negative_zero():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(20000, generator=g)
    cfg = SmaqConfig(stochastic_rounding=False)
    run_case(x, cfg, None, mean=torch.tensor(-0.0), std=torch.tensor(1.0))          # -0 codes, IEEE-divide branch
    run_case(x * 1e30, SmaqConfig(), torch.rand(20000, generator=g))                   # std outside the fast range
    run_case(torch.zeros(5000), SmaqConfig(), torch.rand(5000, generator=g))           # std == 0 -> 1
    xi = x.clone()
    xi[5], xi[6], xi[7] = float("inf"), float("-inf"), float("nan")
    run_case(xi, cfg, None, mean=torch.tensor(0.0), std=torch.tensor(1.0))   # inf saturates, NaN -> 0; counted


def test_decoder_reads_an_oracle_written_stream():
    x, g = make_outlier_tensor(70001, seed=99)
    probs = torch.rand(x.numel(), generator=g)
    cfg = SmaqConfig()
    res = smaq_roundtrip(x, cfg, probs=probs)
    p = opack.pack(res, cfg)
    lay = cabi_pack.layout(x.numel(), cfg)
    buf = cabi_pack.upload_oracle_packed(p, lay, DEV)
    y = cabi_pack.decode(buf, lay)
    assert_bit_equal(y.cpu(), smaq_roundtrip(x, cfg, probs=probs, saturate=True).y, "decode(oracle stream)")


def test_unaligned_input_and_output():
    x, g = make_outlier_tensor(40001, seed=5)
    probs = torch.rand(x.numel(), generator=g)
    cfg = SmaqConfig()
    res = smaq_roundtrip(x, cfg, probs=probs)
    ms = cabi.mean_std_tensor(res.mean, res.std, DEV)
    bx = torch.empty(x.numel() + 1, device=DEV)
    bp = torch.empty(x.numel() + 2, device=DEV)
    by = torch.empty(x.numel() + 3, device=DEV)
    bx[1:].copy_(x)
    bp[2:].copy_(probs)
    buf, lay = cabi_pack.encode(bx[1:], ms, cabi.codec_params(cfg), cfg, probs=bp[2:])
    cabi_pack.decode(buf, lay, out=by[3:])
    assert_bit_equal(by[3:].cpu(), smaq_roundtrip(x, cfg, probs=probs, saturate=True).y, "unaligned")


def test_philox_encode_decode_equals_fused_roundtrip_and_is_deterministic():
    """Performance path: the packed pipeline and the fused fake-quant kernel draw the same Philox
    numbers for the same (seed, offset), so they must agree bit for bit; and the stream is
    byte-identical run to run (deterministic placement)."""
    x, _ = make_outlier_tensor((1 << 22) + 11, seed=8)
    xd = x.to(DEV)
    cfg = SmaqConfig()
    ms = cabi.stats_full(xd)
    params = cabi.codec_params(cfg, seed=77, offset=5, saturate=True)
    buf1, lay = cabi_pack.encode(xd, ms, params, cfg)
    buf2, _ = cabi_pack.encode(xd, ms, params, cfg)
    used = lay.extras_off + 4 * cabi_pack.sections(buf1, lay)[0].extras_words
    assert torch.equal(buf1[:used], buf2[:used])
    y = cabi_pack.decode(buf1, lay)
    fused = cabi.roundtrip(xd, ms, params)
    assert torch.equal(y.view(torch.int32), fused.view(torch.int32))


def test_plugin_encode_decode_and_size_accounting():
    fp = make_plugin()
    x, _ = make_outlier_tensor(1 << 20, seed=4)
    xd = x.to(DEV).view(64, 128, 128)
    packed = fp.encode(xd)
    y = fp.decode(packed)
    assert y.shape == xd.shape
    h = packed.header()
    assert h.status == 0 and h.n == x.numel()
    ratio = 32 * x.numel() / packed.payload_bits()
    assert 4.6 < ratio < 5.3
    # stored bytes: payload + table + alignment only
    assert packed.used_bytes() * 8 - packed.payload_bits() < 8 * (128 + 4 * (packed.layout.n_cta_tiles + 1) + 128) + 32 * packed.layout.n_warp_tiles
    err = (y - xd).abs()
    ms = fp.statistics(xd.view(-1)).cpu()
    assert float(err[(xd - ms[0]).abs() <= 2.5 * ms[1]].max()) <= ms[1].item() / 15 + 1e-6


@pytest.mark.parametrize("n", [1 << 28, 1 << 30])
def test_full_size_properties(n):
    """Config-2 scale (256 Mi elements): no oracle at this size; size-independent properties instead —
    packed pipeline == fused kernel under the same Philox stream; header count == counting kernel;
    table monotone and consistent with the bitmap popcounts."""
    import ctypes as C
    from smart_compress import _native as N

    g = torch.Generator(device=DEV).manual_seed(1234)
    xd = torch.randn(n, generator=g, device=DEV)
    xd[torch.randint(0, n, (n // 100,), generator=g, device=DEV)] *= 10
    cfg = SmaqConfig()
    ms = cabi.stats_full(xd)
    params = cabi.codec_params(cfg, seed=5, offset=9, saturate=True)
    buf, lay = cabi_pack.encode(xd, ms, params, cfg)
    y = cabi_pack.decode(buf, lay)
    fused = cabi.roundtrip(xd, ms, params)
    assert torch.equal(y.view(torch.int32), fused.view(torch.int32))
    del fused, y
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    N.check(N.load().smaq_count_outliers(xd.data_ptr(), n, ms.data_ptr(), C.byref(params), counter.data_ptr(),
                                         N.stream_ptr(DEV)), "count")
    hdr_raw = bytes(buf[:128].cpu().numpy())
    hdr = N.PackedHeader.from_buffer_copy(hdr_raw[: C.sizeof(N.PackedHeader)])
    assert hdr.status == 0 and hdr.n_outlier == int(counter.item())
    table = buf[lay.table_off: lay.table_off + 4 * (lay.n_cta_tiles + 1)].view(torch.int32).long()
    tags = buf[lay.planes_off: lay.planes_off + lay.planes_bytes].view(torch.int32).view(lay.n_warp_tiles, 6, 32)[:, 0, :]
    pop = torch.zeros(lay.n_warp_tiles, dtype=torch.int64, device=DEV)
    t = tags.long() & 0xFFFFFFFF
    for b in range(32):
        pop += ((t >> b) & 1).sum(dim=1)
    words = ((pop * 2 + 31) // 32).view(-1, 8).sum(dim=1)  # word-aligned segment per warp tile
    assert torch.equal(table[1:] - table[:-1], words)
    assert int(pop.sum()) == hdr.n_outlier and int(table[-1]) == hdr.extras_words
