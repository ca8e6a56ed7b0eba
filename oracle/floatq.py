"""CPU oracle for the FP8 / FP16 / BF16 emulation (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED.  The arithmetic is not in the reference: it is the third-party package
``qtorch==0.2.0`` (pyproject.toml:10, poetry.lock:773-781), imported at
smart_compress/util/pytorch/quantization.py:3 and called at :147-149 (rounding="nearest", to
find the largest representable value) and :191-193 (rounding="stochastic").  qtorch is neither
vendored under /root/reference nor installed in this image, and the reference holds no test or
golden vector at that boundary, so this file can only restate qtorch 0.2.0's PUBLISHED
algorithm (quant_cuda/bit_helper.cu + float_kernel.cu of the QPyTorch project) from knowledge of
that source; it has not been executed against qtorch.  What IS pinned is the reference's own
wrapper around it (quantization.py:187-204), restated in ``float_quantize`` below.

qtorch 0.2.0, per element (x an fp32, r an int32 random number, man/exp the target widths):
    bits  = bit pattern of x
    mask  = (1 << (23 - man)) - 1
    q     = (bits + (r & mask)) & ~mask            stochastic   (round_bitwise_stochastic)
    q     = (bits + (1 << (22 - man))) & ~mask     nearest      (round_bitwise_nearest)
    clip_exponent(q):
        q == 0                     -> 0
        e = (q << 1) >> 24         stored exponent
        e > MAX_E                  -> sign(x) | MAX_E << 23 | (top `man` mantissa bits set)   saturate, never inf
        e < MIN_E                  -> sign(x) | MIN_E << 23                                  no subnormals
    MAX_E = 127 + 2^(exp-1)  (0.2.0 reserves no exponent code for inf/NaN; newer releases use
    2^(exp-1) - 1 and add subnormals — ``max_exp_bias=-1`` selects the newer maximum),
    MIN_E = 127 - (2^(exp-1) - 2).
On CUDA, r comes from ``randint_like(x, INT_MAX)``; on CPU from a std::mt19937 seeded by
random_device, i.e. the reference's FP8 path is not reproducible on CPU even with a torch seed.
Both the oracle and the kernel therefore take r as an explicit input.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import numpy as np
import torch

FLT_MAX = np.float32(np.finfo(np.float32).max)
FLT_EPS = np.float32(np.finfo(np.float32).eps)


def _consts(exp: int, man: int, max_exp_bias: int = 0):
    mask = np.uint32((1 << (23 - man)) - 1)
    max_e = (1 << (exp - 1)) + 127 + max_exp_bias
    min_e = -((1 << (exp - 1)) - 2) + 127
    max_man = (((0xFFFFFFFF << 9) & 0xFFFFFFFF) >> 9 >> (23 - man)) << (23 - man)
    return mask, max_e, min_e, np.uint32((max_e << 23) | max_man), np.uint32(min_e << 23)


def qtorch_float_quantize(x: np.ndarray, exp: int, man: int, rounding: str, r: np.ndarray | None = None,
                          max_exp_bias: int = 0) -> np.ndarray:
    """Restatement of qtorch 0.2.0 ``float_quantize`` on a numpy fp32 array."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    bits = x.view(np.uint32)
    mask, max_e, min_e, max_bits, min_bits = _consts(exp, man, max_exp_bias)
    with np.errstate(over="ignore"):
        if rounding == "stochastic":
            assert r is not None
            add = np.ascontiguousarray(r).astype(np.int64).astype(np.uint32) & mask
        elif rounding == "nearest":
            add = np.uint32(1 << (22 - man))
        else:
            raise ValueError(rounding)
        q = (bits + add).astype(np.uint32) & ~mask
    e = ((q << np.uint32(1)) >> np.uint32(24)).astype(np.int64)
    sign = bits & np.uint32(0x80000000)
    out = q.copy()
    hi = (q != 0) & (e > max_e)
    lo = (q != 0) & (e < min_e)
    out[hi] = sign[hi] | max_bits
    out[lo] = sign[lo] | min_bits
    return out.view(np.float32).reshape(x.shape)


def max_value(exp: int, man: int, max_exp_bias: int = 0) -> np.float32:
    """quantization.py:138-150: quantize(finfo(float32).max, exp, man, rounding="nearest")."""
    return qtorch_float_quantize(np.array([FLT_MAX]), exp, man, "nearest", max_exp_bias=max_exp_bias)[0]


def float_quantize(x: torch.Tensor, exp: int, man: int, r: torch.Tensor, *, check_inf: bool = True,
                   precision: int = 32, max_exp_bias: int = 0) -> torch.Tensor:
    """The reference's wrapper, quantization.py:187-204, with qtorch replaced by the restatement."""
    is_16_bit = precision == 16
    src = x.float() if is_16_bit else x
    rv = torch.from_numpy(
        qtorch_float_quantize(src.detach().cpu().numpy(), exp, man, "stochastic", r.detach().cpu().numpy(),
                              max_exp_bias)
    ).clone()
    if check_inf:  # :195-199
        mv = torch.tensor(max_value(exp, man, max_exp_bias))
        should_be_inf = torch.abs(rv - mv) <= torch.tensor(FLT_EPS)
        rv[should_be_inf] = float("inf")
    return rv.half() if is_16_bit else rv
