"""CPU restatement of the in-kernel random numbers of the SmaQ kernels (TEST INFRASTRUCTURE ONLY).

The reference draws its uniforms with ``torch.rand_like`` (smart.py:94); with explicit ``probs`` the CUDA kernels
consume those bit for bit (parity mode).  Without them the kernels draw their own, and this file restates how, so
that the performance path too is checked against the CPU oracle element for element:

* Philox4x32 with SEVEN rounds (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
  multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85) —
  smart-quantization_b200/csrc/common.cuh ``philox4x32``; key = the 64-bit seed, counter = (call index lo, hi,
  stream offset lo, hi).  ``tests/test_oracle_rng.py`` pins this file to the published Random123 known-answer
  vectors (ten rounds) before anything relies on it;
* 16 random bits per element, one call per 16 elements (common.cuh ``rnd16_*``): elements are taken in groups of
  eight (group g = elements 8g .. 8g+7); groups g and g ^ 32 share the call with index ((g >> 6) << 5) | (g & 31);
  word q of the call serves elements 2q and 2q+1 of both groups; of a word b3 b2 b1 b0 the group with bit 5 clear
  takes the halves (b1 b0), (b3 b2), the other one the windows in between, (b2 b1), (b0 b3);
* the rounding rule with these numbers (smaq_math.cuh ``encode_pair`` kRng): code = floor(c + q),
  q = (k + 1/2) / 2^16 — i.e. ``floor(c) + [frac >= p]`` with p = 1 - q, the reference's expression
  ``floor(c) + round(relu((frac - p) + 0.5))`` except on the set |frac - p| <= 2^-25 (``oracle.smaq`` ``rng_rule``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

ROUNDS = 7
M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, seed: int, rounds: int = ROUNDS):
    """Vectorised Philox4x32-`rounds`: uint32 arrays (or scalars) c0..c3 -> four uint32 arrays."""
    c0, c1, c2, c3 = (np.atleast_1d(np.asarray(c, dtype=np.uint64)) & MASK32 for c in (c0, c1, c2, c3))
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0          # 32 x 32 -> 64 bits, exact in uint64
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & MASK32, n2, p0 & MASK32
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def rnd16(n: int, seed: int, offset: int = 0, first: int = 0) -> np.ndarray:
    """The 16-bit value k of elements first .. first+n-1 of stream `offset` (uint16 array)."""
    e = np.arange(first, first + n, dtype=np.uint64)
    g = e >> np.uint64(3)
    j = (e & np.uint64(7)).astype(np.int64)
    call = ((g >> np.uint64(6)) << np.uint64(5)) | (g & np.uint64(31))
    sub = ((g >> np.uint64(5)) & np.uint64(1)).astype(np.int64)
    ones = np.ones_like(call)
    words = philox4x32(call & MASK32, call >> np.uint64(32), ones * np.uint64(offset & 0xFFFFFFFF),
                       ones * np.uint64((offset >> 32) & 0xFFFFFFFF), seed)
    w = np.choose(j >> 1, words).astype(np.uint64)
    t = (2 * (j & 1) + sub).astype(np.uint64)
    lo = (w >> (np.uint64(8) * t)) & np.uint64(0xFF)
    hi = (w >> (np.uint64(8) * ((t + np.uint64(1)) & np.uint64(3)))) & np.uint64(0xFF)
    return (lo | (hi << np.uint64(8))).astype(np.uint16)


def probs_for(n: int, seed: int, offset: int = 0, first: int = 0) -> np.ndarray:
    """p = 1 - q = (65535.5 - k) / 2^16 as fp32 (exact): what ``oracle.smaq.smaq_roundtrip(rng_rule=True)`` takes."""
    k = rnd16(n, seed, offset, first).astype(np.float64)
    return ((65535.5 - k) / 65536.0).astype(np.float32)


def floatq_fields(n: int, seed: int, offset: int = 0, man_bits: int = 2, first: int = 0) -> np.ndarray:
    """What the float-emulation kernels (csrc/float_quantize.cu ``fq_k16`` / ``rand_field``) add to an element's bit
    pattern before truncating when they draw their own numbers: one Philox4x32-7 call per eight elements (counter =
    element index / 8), 16 bits per element — half (j & 1) of word j >> 1 — placed at the top of the (23 - man)-bit
    field plus half a step of the 16-bit grid.  Returned as int32: feeding it to ``oracle.floatq.float_quantize`` as
    its ``r`` reproduces the kernel (the oracle masks with the same field mask)."""
    e = np.arange(first, first + n, dtype=np.uint64)
    g = e >> np.uint64(3)
    j = (e & np.uint64(7)).astype(np.int64)
    ones = np.ones_like(g)
    words = philox4x32(g & MASK32, g >> np.uint64(32), ones * np.uint64(offset & 0xFFFFFFFF),
                       ones * np.uint64((offset >> 32) & 0xFFFFFFFF), seed)
    w = np.choose(j >> 1, words).astype(np.uint64)
    k16 = np.where(j & 1, w >> np.uint64(16), w & np.uint64(0xFFFF))
    rshift = (23 - man_bits) - 16
    if rshift >= 0:
        field = (k16 << np.uint64(rshift)) | np.uint64((1 << (rshift - 1)) if rshift > 0 else 0)
    else:
        field = k16 >> np.uint64(-rshift)
    return field.astype(np.int64).astype(np.int32)
