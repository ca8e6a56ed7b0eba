"""GPU parity tests for the packed SmaQ stream (smaq_encode / smaq_decode).

Contract: given the reference's mean/std and the same uniform numbers, the packed payload (tag
words, base fields, extras stream, tile table, counters) is BYTE-IDENTICAL to oracle/pack.py
applied to the reference's integer codes, and decode() is bit-identical to the reference's
round trip with codes saturated at the field width."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import pack as opack
from oracle.smaq import SmaqConfig, compressed_bits, smaq_roundtrip
from smart_compress import _native as N
from tests import cabi, cabi_pack
from tests.golden_util import assert_bit_equal, load_golden, uses_bn
from tests.test_gpu_smaq import make_outlier_tensor, make_plugin

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = load_golden()


def run_case(x, cfg, probs, *, mean=None, std=None, all_positive=False, idx=None, rng=None):
    """rng=(seed, offset): the kernel draws its own uniforms (performance path) and the oracle is fed the same
    numbers from oracle/rng.py under the kernels' rounding rule; otherwise `probs` goes to both (parity mode)."""
    rule = rng is not None
    if rule:
        from oracle import rng as orng
        probs = torch.from_numpy(orng.probs_for(x.numel(), seed=rng[0], offset=rng[1])).view(x.shape)
    res = smaq_roundtrip(x, cfg, probs=probs, idx=idx, mean=mean, std=std, rng_rule=rule)
    ms = cabi.mean_std_tensor(res.mean, res.std, DEV)
    xd = x.to(DEV).contiguous().view(-1)
    pd = None if (probs is None or rule) else probs.to(DEV).contiguous().view(-1)
    params = cabi.codec_params(cfg) if not rule else cabi.codec_params(cfg, seed=rng[0], offset=rng[1])
    buf, lay = cabi_pack.encode(xd, ms, params, cfg, probs=pd)
    y = cabi_pack.decode(buf, lay, all_positive=all_positive)
    want = smaq_roundtrip(x, cfg, probs=probs, idx=idx, mean=res.mean, std=res.std, saturate=True,
                          all_positive=all_positive, rng_rule=rule)
    p = opack.pack(res, cfg)
    hdr = cabi_pack.assert_stream_equals_oracle(buf, lay, p)
    assert hdr.n_saturated == p.n_saturated
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


@pytest.mark.parametrize("n", [33, 1024, 8193, 100003, (1 << 22) + 11])
def test_encode_with_in_kernel_random_numbers_byte_exact(n):
    """Performance path: the kernel draws its own uniforms (Philox4x32-7, 16 bits per element).  oracle/rng.py
    restates the generator on the CPU, so the stream is still checked byte for byte against oracle/pack.py and the
    decoder against the oracle's saturated round trip — not against another CUDA kernel."""
    x, _ = make_outlier_tensor(n, seed=n + 5)
    run_case(x, SmaqConfig(), None, rng=(77 + n, 5))


def test_config0_2p26_in_kernel_random_numbers_byte_exact():
    x, _ = make_outlier_tensor(1 << 26, seed=1234)
    run_case(x, SmaqConfig(), None, rng=(1234, (1 << 40) + 9))   # a stream offset beyond 32 bits


def test_config0_2p26_encode_decode_byte_exact():
    """BASELINE configs[0] itself (64 Mi elements, the bench recipe's seed): every byte of the stream against
    oracle/pack.py, the decoder against the oracle's saturated round trip."""
    x, g = make_outlier_tensor(1 << 26, seed=1234)
    run_case(x, SmaqConfig(), torch.rand(x.numel(), generator=g))


@pytest.mark.parametrize("bm,bo", [(4, 6), (5, 9), (6, 6), (8, 12), (4, 4), (7, 8), (6, 9)])
def test_other_bit_widths(bm, bo):
    x, g = make_outlier_tensor(50001, seed=bm * 31 + bo)
    cfg = SmaqConfig(num_bits_main=bm, num_bits_outlier=bo)
    run_case(x, cfg, torch.rand(x.numel(), generator=g))


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items()
                                        if not c["same_object"] and not uses_bn(c) and 4 <= c["cfg"].num_bits_main <= 8
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
    """Performance path: the packed pipeline and the fused fake-quant kernel draw the same numbers for the same
    (seed, offset, element), so they agree bit for bit (both are also checked against the oracle on their own);
    and the stream is byte-identical run to run."""
    x, _ = make_outlier_tensor((1 << 22) + 11, seed=8)
    xd = x.to(DEV)
    cfg = SmaqConfig()
    ms = cabi.stats_full(xd)
    params = cabi.codec_params(cfg, seed=77, offset=5, saturate=True)
    buf1, lay = cabi_pack.encode(xd, ms, params, cfg)
    buf2, _ = cabi_pack.encode(xd, ms, params, cfg)
    assert torch.equal(buf1, buf2)   # poison included: the same bytes are written, the same are left alone
    y = cabi_pack.decode(buf1, lay)
    fused = cabi.roundtrip(xd, ms, params)
    assert torch.equal(y.view(torch.int32), fused.view(torch.int32))


@pytest.mark.parametrize("kind", ["relu", "relu_shifted", "lower_outlier", "far"])
def test_zero_on_grid_keeps_exact_zeros_and_matches_the_oracle(kind):
    """zero_on_grid (packed activation storage; not in the reference): the encoder moves the mean by at most half a
    quantisation step so that 0.0 is a grid point.  The mean it used is the oracle's restatement of the rule bit
    for bit; the stream is byte-identical to oracle/pack.py given that mean; and every exact zero decodes to
    exactly 0.0 (up to the ~1e-6 of them whose scaled value rounds across an integer)."""
    from oracle import rng as orng
    from oracle.smaq import snap_mean_to_zero

    g = torch.Generator().manual_seed(12)
    n = 300007
    x = torch.randn(n, generator=g)
    if kind == "relu":
        x = x.relu()                          # zero inside the main range (z0 ~ -0.68)
    elif kind == "relu_shifted":
        x = (x * 0.01 + 0.004).relu()
    elif kind == "lower_outlier":
        x = torch.where(torch.rand(n, generator=g) < 0.03, torch.zeros(n), x * 0.3 + 0.5)   # zero ~ 1.6 sigma below the mean
    else:
        x = torch.where(torch.rand(n, generator=g) < 0.001, torch.zeros(n), x * 0.05 + 3.0)  # zero far outside: no-op
    cfg = SmaqConfig()
    xd = x.to(DEV)
    ms = cabi.stats_full(xd)
    msc = ms.cpu()
    buf, lay = cabi_pack.encode(xd, ms, cabi.codec_params(cfg, seed=31, offset=4, zero_on_grid=True), cfg)
    hdr, _, _ = cabi_pack.sections(buf, lay)
    want_mean = snap_mean_to_zero(msc[0], msc[1], cfg)
    assert np.float32(hdr.mean).view(np.uint32) == want_mean.numpy().view(np.uint32), (hdr.mean, float(want_mean))
    step = float(msc[1]) / 15
    if kind == "far":
        assert hdr.mean == float(msc[0])
    else:
        assert abs(hdr.mean - float(msc[0])) <= 0.51 * step * (42 / 15 if kind == "lower_outlier" else 1.0) + 1e-12
    probs = torch.from_numpy(orng.probs_for(n, seed=31, offset=4))
    res = smaq_roundtrip(x, cfg, probs=probs, mean=want_mean, std=msc[1], rng_rule=True)
    cabi_pack.assert_stream_equals_oracle(buf, lay, opack.pack(res, cfg))
    y = cabi_pack.decode(buf, lay).cpu()
    zeros = x == 0
    if kind != "far":
        wrong = int((y[zeros] != 0).sum())
        assert wrong <= max(2, int(2e-5 * int(zeros.sum()))), (wrong, int(zeros.sum()))
    want = smaq_roundtrip(x, cfg, probs=probs, mean=want_mean, std=msc[1], rng_rule=True, saturate=True)
    assert_bit_equal(y, want.y, "decode")


@pytest.mark.parametrize("n", [1000, 8193, 300007, (1 << 22) + 11])
def test_split_buffers_and_compacted_extras_match_the_oracle(n):
    """Exact-size storage: header + planes in one buffer, the extras compacted to their used words behind a
    per-warp-tile table (smaq_extras_compact).  Table = exclusive prefix of the oracle's segment
    sizes, dense extras = the oracle's used words in tile order, decode_split of either form = the plain decode, and
    the allocation is the reference's accounting (smart.py:184-187) plus table and padding."""
    fp = make_plugin()
    x, _ = make_outlier_tensor(n, seed=n + 2)
    xd = x.to(DEV)
    torch.manual_seed(9)
    whole = make_plugin().encode(xd)
    torch.manual_seed(9)
    packed = fp.encode(xd, split=True)
    lay = packed.layout
    assert packed.buffer.numel() == lay.extras_off and packed.extras.numel() == lay.extras_capacity_bytes
    hb = C.sizeof(N.PackedHeader)   # (the rest of the 128-byte header slot is never written)
    assert torch.equal(packed.buffer[:hb], whole.buffer[:hb])
    assert torch.equal(packed.buffer[lay.planes_off:], whole.buffer[lay.planes_off: lay.extras_off])
    y_fixed = fp.decode(packed)
    assert torch.equal(y_fixed, make_plugin().decode(whole))
    h = packed.header()
    # the oracle's view of the same stream
    msc = torch.tensor([h.mean, h.std_raw])
    probs = torch.from_numpy(__import__("oracle.rng", fromlist=["x"]).probs_for(n, seed=9, offset=0))
    res = smaq_roundtrip(x, SmaqConfig(), probs=probs, mean=msc[0], std=msc[1], rng_rule=True)
    p = opack.pack(res, SmaqConfig())
    fp.compact(packed, h.extras_words)
    assert packed.table is not None and h.extras_words == p.extras_words
    table = packed.table[: lay.n_warp_tiles + 1].cpu().numpy().astype(np.int64)
    want_table = np.concatenate([[0], np.cumsum(p.seg_used)])
    assert np.array_equal(table, want_table)
    dense = packed.extras[: 4 * h.extras_words].view(torch.int32).cpu().numpy().view(np.uint32)
    want_dense = np.concatenate([p.extras[t, : p.seg_used[t]] for t in range(p.planes.shape[0])]) if h.extras_words else np.zeros(0, np.uint32)
    assert np.array_equal(dense, want_dense)
    assert torch.equal(fp.decode(packed), y_fixed)
    bits = packed.allocated_bytes() * 8
    assert bits - packed.payload_bits() <= 8 * (128 + 64 + 4) + 64 * lay.n_warp_tiles + 6 * 1024   # header, look-ahead pad, table + padding per tile, ragged last tile


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
    # stored bytes: payload + header + word alignment per warp tile only
    assert packed.used_bytes() * 8 - packed.payload_bits() < 8 * 128 + 32 * packed.layout.n_warp_tiles
    err = (y - xd).abs()
    ms = fp.statistics(xd.view(-1)).cpu()
    assert float(err[(xd - ms[0]).abs() <= 2.5 * ms[1]].max()) <= ms[1].item() / 15 + 1e-6


@pytest.mark.parametrize("n", [1 << 28, 1 << 30])
def test_full_size_properties(n):
    """Config-2 scale (256 Mi and 1 Gi elements): the full oracle does not finish in seconds here, so —
    size-independent properties (packed pipeline == fused kernel under the same stream; header counters == counting
    kernel == bitmap popcounts) plus the CPU ORACLE on 64 warp tiles sampled across the tensor (the last tile, past
    element 2^30 - 1024, included: 64-bit indexing and Philox counters beyond 2^27 calls)."""
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
    tags = buf[lay.planes_off: lay.planes_off + lay.planes_bytes].view(torch.int32).view(lay.n_warp_tiles, 6, 32)[:, 0, :]
    pop = torch.zeros(lay.n_warp_tiles, dtype=torch.int64, device=DEV)
    t = tags.long() & 0xFFFFFFFF
    for b in range(32):
        pop += ((t >> b) & 1).sum(dim=1)
    words = (pop * 2 + 31) // 32   # word-aligned segment per warp tile
    assert int(pop.sum()) == hdr.n_outlier and int(words.sum()) == hdr.extras_words
    # the sampled oracle check: 64 warp tiles spread over the tensor, each re-encoded by the CPU oracle from the
    # same statistics and the kernel's own random numbers (oracle/rng.py), compared word for word
    from oracle import rng as orng
    msc = ms.cpu()
    raw_planes = buf[lay.planes_off: lay.planes_off + lay.planes_bytes].view(torch.int32).view(lay.n_warp_tiles, 6, 32)
    raw_extras = buf[lay.extras_off: lay.extras_off + lay.n_warp_tiles * lay.extras_stride_bytes].view(torch.int32).view(lay.n_warp_tiles, -1)
    for wt in torch.linspace(0, lay.n_warp_tiles - 1, 64).long().tolist():
        xs = xd[wt * 1024: (wt + 1) * 1024].cpu()
        probs = torch.from_numpy(orng.probs_for(1024, seed=5, offset=9, first=wt * 1024))
        res = smaq_roundtrip(xs, cfg, probs=probs, mean=msc[0], std=msc[1], rng_rule=True)
        p = opack.pack(res, cfg)
        assert np.array_equal(raw_planes[wt].cpu().numpy().view(np.uint32), p.planes[0]), wt
        u = int(p.seg_used[0])
        assert np.array_equal(raw_extras[wt, :u].cpu().numpy().view(np.uint32), p.extras[0, :u]), wt
