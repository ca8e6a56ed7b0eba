"""CPU/torch oracle for S2FP8 (TEST INFRASTRUCTURE ONLY).

Op-for-op restatement of reference smart_compress/compress/s2fp8.py:31-48 with torch operators
on whatever device the input lives on.  The e5m2 rounding inside it is oracle/floatq.py
(PARITY UNPINNED, see there: qtorch 0.2.0 is absent).  log2/pow differ in the last ulp between
CPU (SLEEF) and CUDA (libdevice), so bit-level comparisons of a CUDA kernel are only meaningful
against this oracle evaluated with CUDA tensors on the same GPU; against the
CPU evaluation the tests use a stated tolerance.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import torch

from .floatq import float_quantize


def s2fp8_statistics(x: torch.Tensor):
    """s2fp8.py:33-37 -> (mu, m)."""
    x_abs = x.abs()
    l2 = torch.where(x_abs == 0.0, x_abs, torch.log2(x_abs))
    return torch.mean(l2), torch.max(l2)


@torch.no_grad()
def s2fp8(x: torch.Tensor, r: torch.Tensor, *, mu=None, m=None, check_inf: bool = True, max_exp_bias: int = 0):
    """s2fp8.py:31-48.  ``r``: int32 random numbers for the stochastic rounding.  Returns (y, mu, m, T)."""
    signs = torch.sign(x)
    x_abs = x.abs()
    if mu is None or m is None:
        mu, m = s2fp8_statistics(x)
    mu = torch.as_tensor(mu, dtype=x.dtype, device=x.device)
    m = torch.as_tensor(m, dtype=x.dtype, device=x.device)
    alpha = 15.0 / (m - mu)
    beta = -alpha * mu
    beta_pow2 = 2.0 ** beta
    pre = x_abs.clone().pow_(alpha).mul_(beta_pow2)
    truncated = float_quantize(pre, 5, 2, r, check_inf=check_inf, max_exp_bias=max_exp_bias).to(x.device)
    y = ((truncated * beta_pow2.reciprocal_()) ** alpha.reciprocal_()) * signs
    return y, mu, m, truncated
