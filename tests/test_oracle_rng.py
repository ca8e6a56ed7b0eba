"""oracle/rng.py — the CPU restatement of the kernels' random numbers — pinned and characterised (CPU only).

* Philox4x32 against the published Random123 known-answer vectors (ten rounds: kat_vectors, philox4x32 10);
* the 16-bit extraction: every value uniform, the eight values of a group disjoint bit fields of one call;
* the rounding rule the kernels apply to these numbers against the reference's expression."""
import numpy as np
import torch

from oracle import rng
from oracle.smaq import round_stochastic


def test_philox4x32_10_known_answers():
    # Random123-1.14 examples/kat_vectors: philox4x32 10 <counter> <key> <result>
    kat = [((0, 0, 0, 0), 0, (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, 0xFFFFFFFFFFFFFFFF, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), 0xA4093822 | (0x299F31D0 << 32),
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in kat:
        got = tuple(int(w[0]) for w in rng.philox4x32(*ctr, key, rounds=10))
        assert got == want


def test_seven_rounds_is_a_prefix_of_the_same_iteration():
    """The kernels run seven rounds (Salmon et al.: the fewest that pass BigCrush): same round function, same key
    schedule — ten rounds from a counter equal three more rounds from the seven-round state with the advanced key."""
    c = (123, 456, 789, 1011)
    seed = 0x0123456789ABCDEF
    s7 = rng.philox4x32(*c, seed, rounds=7)
    k0 = ((seed & 0xFFFFFFFF) + 7 * rng.W0) & 0xFFFFFFFF
    k1 = ((seed >> 32) + 7 * rng.W1) & 0xFFFFFFFF
    s10 = rng.philox4x32(*(int(w[0]) for w in s7), k0 | (k1 << 32), rounds=3)
    assert [int(w[0]) for w in s10] == [int(w[0]) for w in rng.philox4x32(*c, seed, rounds=10)]


def test_rnd16_layout():
    seed, off = 99, 7
    k = rng.rnd16(4096, seed, off)
    # group g and g ^ 32 share a call; inside a group the eight values are the call's 128 bits cut into 16-bit fields
    for g in (0, 5, 31, 64, 100):
        call = ((g >> 6) << 5) | (g & 31)
        w = [int(x[0]) for x in rng.philox4x32(call, 0, off, 0, seed)]
        raw = b"".join(int(x).to_bytes(4, "little") for x in w)
        a = k[8 * g: 8 * g + 8]
        b = k[8 * (g ^ 32): 8 * (g ^ 32) + 8]
        first, second = (a, b) if not (g >> 5) & 1 else (b, a)
        assert b"".join(int(v).to_bytes(2, "little") for v in first) == raw          # the disjoint halves, in order
        for q in range(4):
            by = raw[4 * q: 4 * q + 4]
            assert int(second[2 * q]) == by[1] | (by[2] << 8) and int(second[2 * q + 1]) == by[3] | (by[0] << 8)
    # first / offset arguments: a slice of the stream is the stream's slice
    assert np.array_equal(rng.rnd16(100, seed, off, first=1000), k[1000:1100])
    assert not np.array_equal(rng.rnd16(4096, seed, off + 1), k)


def test_rnd16_is_uniform_and_uncorrelated():
    n = 1 << 22
    k = rng.rnd16(n, 2024, 3).astype(np.int64)
    counts = np.bincount(k >> 8, minlength=256)
    chi2 = float(((counts - n / 256) ** 2 / (n / 256)).sum())
    assert chi2 < 350            # 255 degrees of freedom: mean 255, sd 22.6
    counts = np.bincount(k & 0xFF, minlength=256)
    assert float(((counts - n / 256) ** 2 / (n / 256)).sum()) < 350
    u = (k + 0.5) / 65536 - 0.5
    for lag in (1, 2, 7, 8, 255, 257, 512, 1024):
        r = float((u[:-lag] * u[lag:]).mean() / u.var())
        assert abs(r) < 5 / np.sqrt(n), (lag, r)
    # the documented dependence: an element of a group with bit 5 clear and the element 256 later share one byte
    # (its high byte is the other's low byte), a linear correlation of 2^-8 for those pairs, half that over all pairs
    r256 = float((u[:-256] * u[256:]).mean() / u.var())
    assert 0.5 / 512 < r256 < 2.0 / 512, r256


def test_rng_rule_against_the_reference_expression():
    """floor(c) + [frac >= p] == floor(c) + round(relu((frac - p) + 0.5)) except where |frac - p| <= 2^-25; and
    P(round up) is the fractional part to within 2^-17, with zero mean."""
    g = torch.Generator().manual_seed(0)
    c = (torch.rand(1 << 20, generator=g) - 0.5) * 126
    p = torch.from_numpy(rng.probs_for(c.numel(), 5, 1))
    a, b = round_stochastic(c, p, rng_rule=True), round_stochastic(c, p)
    frac = c - c.floor()
    differ = a != b
    assert bool(((frac - p).abs()[differ] <= 2.0 ** -25).all()) and int(differ.sum()) < 8
    # exhaustive over k for a few fractions: #{k : frac >= p_k} / 65536 vs frac
    allp = torch.from_numpy(((65535.5 - np.arange(65536)) / 65536.0).astype(np.float32))
    errs = []
    for f in (0.0, 2.0 ** -20, 0.1, 0.25, 0.5, 0.7, 1 - 2.0 ** -20, 1 - 2.0 ** -24):
        up = float((torch.tensor(f, dtype=torch.float32) >= allp).float().mean())
        assert abs(up - f) <= 2.0 ** -17 + 1e-12, (f, up)
        errs.append(up - f)
    assert float((torch.tensor(0.0) >= allp).sum()) == 0      # an integer never moves
    fr = torch.rand(4096, generator=g)
    bias = torch.stack([(f >= allp).float().mean() - f for f in fr]).mean()
    assert abs(float(bias)) < 2.0 ** -22
