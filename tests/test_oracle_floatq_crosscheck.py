"""Independent cross-check of the UNPINNED float oracle (oracle/floatq.py): qtorch is not available, but torch's own
dtype casts (float8_e5m2, float8_e4m3fn, float16, bfloat16) are independent implementations of "round an fp32 at m
mantissa bits".  Inside the target's normal range and away from exact ties (qtorch rounds ties away from zero by
adding half an ulp to the bit pattern, IEEE casts round them to even) the two must agree bit for bit; on the ties the
restatement must be the cast's neighbour one target-ulp further from zero whenever they differ.  This anchors the
rounding position and the field arithmetic of the restatement; the clip constants (MAX_E, no subnormals) remain
qtorch 0.2.0's published ones and stay unpinned."""
import numpy as np
import pytest
import torch

from oracle.floatq import qtorch_float_quantize

FORMATS = [  # exp, man, torch dtype, smallest normal, largest finite of the torch dtype
    (5, 2, torch.float8_e5m2, 2.0 ** -14, 57344.0),
    (4, 3, torch.float8_e4m3fn, 2.0 ** -6, 448.0),
    (5, 10, torch.float16, 2.0 ** -14, 65504.0),
    (8, 7, torch.bfloat16, 2.0 ** -126, 3.38e38),
]


def _inputs(lo, hi, n=200_000, seed=7):
    g = np.random.default_rng(seed)
    mag = np.exp(g.uniform(np.log(lo * 1.01), np.log(hi * 0.49), n)).astype(np.float32)
    return (mag * g.choice(np.float32([-1, 1]), n)).astype(np.float32)


@pytest.mark.parametrize("exp,man,dtype,lo,hi", FORMATS)
def test_nearest_agrees_with_torch_casts_off_ties(exp, man, dtype, lo, hi):
    x = _inputs(lo, hi)
    drop = 23 - man
    low = x.view(np.uint32) & np.uint32((1 << drop) - 1)
    tie = low == np.uint32(1 << (drop - 1))
    ours = qtorch_float_quantize(x, exp, man, "nearest")
    cast = torch.from_numpy(x).to(dtype).float().numpy()
    assert np.array_equal(ours[~tie].view(np.uint32), cast[~tie].view(np.uint32))


@pytest.mark.parametrize("exp,man,dtype,lo,hi", FORMATS)
def test_ties_round_away_from_zero(exp, man, dtype, lo, hi):
    x = _inputs(lo, hi, n=50_000, seed=11)
    drop = 23 - man
    bits = (x.view(np.uint32) & ~np.uint32((1 << drop) - 1)) | np.uint32(1 << (drop - 1))  # exact ties
    x = bits.view(np.float32)
    ours = qtorch_float_quantize(x, exp, man, "nearest")
    cast = torch.from_numpy(x).to(dtype).float().numpy()
    up = (bits + np.uint32(1 << (drop - 1))).view(np.float32)  # the neighbour further from zero
    assert np.array_equal(ours.view(np.uint32), up.view(np.uint32))
    differ = ours != cast
    # where the even neighbour is the nearer-to-zero one the two differ by exactly one target ulp
    assert np.all(np.abs(ours[differ]) > np.abs(cast[differ]))
    assert 0.3 < differ.mean() < 0.7


@pytest.mark.parametrize("exp,man,dtype,lo,hi", FORMATS)
def test_stochastic_brackets_and_is_unbiased(exp, man, dtype, lo, hi):
    x = _inputs(lo, hi, n=100_000, seed=3)
    g = np.random.default_rng(5)
    r = g.integers(0, 2**31 - 1, x.shape, dtype=np.int64).astype(np.int32)
    drop = 23 - man
    q = qtorch_float_quantize(x, exp, man, "stochastic", r)
    down = (x.view(np.uint32) & ~np.uint32((1 << drop) - 1)).view(np.float32)  # truncation toward zero
    up = ((x.view(np.uint32) & ~np.uint32((1 << drop) - 1)) + np.uint32(1 << drop)).view(np.float32)
    assert np.all((q == down) | (q == up))
    # P(up) = low bits / 2^drop: the mean relative error vanishes
    rel = ((q.astype(np.float64) - x) / np.abs(x)).mean()
    assert abs(rel) < 4 * 2.0 ** -(man + 1) / np.sqrt(x.size)
