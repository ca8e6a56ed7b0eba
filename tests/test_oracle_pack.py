"""Self-consistency of the packed-stream oracle (oracle/pack.py) — CPU only.

pack -> decode must reproduce, bit for bit, the reference's round trip with codes saturated at the
field width (oracle/smaq.py, saturate=True), and the stream must have exactly the size the
reference accounts for (smart.py:184-187) plus the documented alignment overhead."""
import numpy as np
import pytest
import torch

from oracle import pack as opack
from oracle.smaq import SmaqConfig, compressed_bits, smaq_roundtrip
from tests.golden_util import assert_bit_equal


def make(n, seed, outliers=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, generator=g)
    if outliers:
        x[torch.randperm(n, generator=g)[: max(1, n // 100)]] *= 10
    return x, torch.rand(n, generator=g)


def test_pack_with_the_kernels_own_random_numbers():
    """The rule the kernels apply to their own uniforms (oracle/rng.py, rng_rule) through the packer."""
    from oracle import rng

    x, _ = make(30000, 5)
    cfg = SmaqConfig()
    probs = torch.from_numpy(rng.probs_for(x.numel(), seed=77, offset=3))
    res = smaq_roundtrip(x, cfg, probs=probs, rng_rule=True)
    want = smaq_roundtrip(x, cfg, probs=probs, rng_rule=True, saturate=True)
    assert_bit_equal(opack.decode(opack.pack(res, cfg)), want.y, "decode(pack(x)), rng rule")
    # the rule is the reference's expression except on ties: on this input they agree everywhere
    lit = smaq_roundtrip(x, cfg, probs=probs, saturate=True)
    assert float((lit.y != want.y).float().mean()) < 1e-4


@pytest.mark.parametrize("n", [8, 31, 1024, 1025, 8192, 8193, 50000])
@pytest.mark.parametrize("cfgkw", [dict(), dict(stochastic_rounding=False), dict(num_bits_main=4, num_bits_outlier=6),
                                   dict(num_bits_main=5, num_bits_outlier=9), dict(num_bits_main=6, num_bits_outlier=6)])
def test_pack_decode_equals_saturated_roundtrip(n, cfgkw):
    x, probs = make(n, n)
    cfg = SmaqConfig(**cfgkw)
    res = smaq_roundtrip(x, cfg, probs=probs)
    p = opack.pack(res, cfg)
    y = opack.decode(p)
    want = smaq_roundtrip(x, cfg, probs=probs, saturate=True)
    assert_bit_equal(y, want.y, "decode(pack(x))")
    # exact size: planes + extras (before word alignment) == the reference's accounting
    assert p.payload_bits == compressed_bits(res, cfg)
    n_pad = p.planes.shape[0] * 1024
    xb = cfg.num_bits_outlier - cfg.num_bits_main
    assert p.planes.size * 32 == n_pad * cfg.num_bits_main
    assert 0 <= p.extras_words * 32 - xb * p.n_outlier < 32 * p.planes.shape[0] + 1  # < 32 pad bits per warp tile
    for t in range(p.planes.shape[0]):   # nothing outside a segment's used words
        assert not p.extras[t, p.seg_used[t]:].any()
    # n_saturated is a pure function of the data: scaled values the field cannot hold
    c = res.extras["c"]
    assert p.n_saturated == int((c.abs() > torch.where(res.hi | res.lo, float(cfg.max_code_outlier),
                                                        float(cfg.max_code_main))).sum())


def test_negative_zero_code_survives_truncation():
    cfg = SmaqConfig(stochastic_rounding=False)
    x = torch.tensor([-0.01, 0.01, -3.0, 3.0, 0.5, -0.5, 1.2, -1.2, 0.0])
    res = smaq_roundtrip(x, cfg, mean=torch.tensor(-0.0), std=torch.tensor(1.0))
    assert_bit_equal(opack.decode(opack.pack(res, cfg)), smaq_roundtrip(x, cfg, mean=torch.tensor(-0.0),
                     std=torch.tensor(1.0), saturate=True).y, "-0 mean")


def test_non_finite_codes_are_saturated_or_zeroed_and_counted():
    cfg = SmaqConfig(stochastic_rounding=False)
    x = torch.randn(64).clamp(-2, 2)
    x[3], x[4], x[5] = float("inf"), float("-inf"), float("nan")
    res = smaq_roundtrip(x, cfg, mean=torch.tensor(0.0), std=torch.tensor(1.0))
    p = opack.pack(res, cfg)
    assert p.n_saturated == 3
    tag, S = opack.unpack_values(p)
    assert tag[3] and S[3] == 63          # +inf: upper side, code 63
    assert tag[4] and S[4] == -64         # -inf: lower side, code -63
    assert not tag[5] and S[5] == 0       # NaN: main, code 0


def test_lane_order_is_a_permutation():
    perm = opack.lane_order_index(4096).reshape(-1)
    assert np.array_equal(np.sort(perm), np.arange(4096))
    assert perm[0] == 0 and perm[1] == 1 and perm[8] == 256 and perm[32] == 8
