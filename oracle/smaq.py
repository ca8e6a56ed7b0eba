"""CPU oracle for the SmaQ compress->decompress path (TEST INFRASTRUCTURE ONLY).

A restatement, op for op, of the reference's ``SmartFP.__call__``
(/root/reference/smart_compress/compress/smart.py:110-190) using the same torch
CPU operators in the same order, so that every intermediate rounds exactly as
the reference's does.  Unlike the reference it (1) takes the random numbers and
the sample indices as explicit inputs, (2) returns the intermediates a packed
codec needs to be checked against (class masks, integer codes), and (3) never
touches a profiler or a logger.

Pinned: ``tests/test_oracle_pinned.py`` checks this file bit-for-bit against the
fixtures in ``tests/golden/`` that ``oracle/gen_golden.py`` produced by running
the unmodified reference in the build container (and, when /root/reference is
present, against the live reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  The product (``smart-quantization_b200/``) never does.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch


@dataclass
class SmaqConfig:
    """The hparams SmartFP reads (smart.py:17-69) with the reference defaults."""

    num_samples: int = 16
    use_sample_stats: bool = False
    stochastic_rounding: bool = True
    num_bits_main: int = 6
    num_bits_outlier: int = 8
    main_std_dev_threshold: float = 1.0
    outlier_std_dev_threshold: float = 2.5
    min_size: int = 8
    use_range_std_dev: bool = False
    use_batch_norm: bool = False
    bn_scalar_params: bool = False
    precision: int = 32

    # smart.py:75-80
    @property
    def range_outlier(self) -> float:
        return ((2 ** (self.num_bits_outlier - 2)) - 1) / (
            self.outlier_std_dev_threshold - self.main_std_dev_threshold
        )

    @property
    def range_normal(self) -> float:
        return ((2 ** (self.num_bits_main - 2)) - 1) / self.main_std_dev_threshold

    # smart.py:82-84
    @property
    def clamped_range(self):
        return (1e-4, 1e4) if self.precision == 16 else (1e-38, 1e38)

    # largest magnitude a packed code can hold (tag + sign/side + magnitude bits)
    @property
    def max_code_main(self) -> int:
        return (1 << (self.num_bits_main - 2)) - 1

    @property
    def max_code_outlier(self) -> int:
        return (1 << (self.num_bits_outlier - 2)) - 1


@dataclass
class SmaqResult:
    y: torch.Tensor
    mean: Optional[torch.Tensor] = None
    std: Optional[torch.Tensor] = None  # as returned by the statistics step (before the ==0 fix-up)
    hi: Optional[torch.Tensor] = None  # z > +t   (bool)
    lo: Optional[torch.Tensor] = None  # z < -t   (bool)
    code: Optional[torch.Tensor] = None  # fp32 holding the rounded integer code (smart.py:166-169)
    passthrough: bool = False  # numel < min_size: reference returns its input object (smart.py:125-128)
    extras: dict = field(default_factory=dict)


def sample_mean_std(data: torch.Tensor, idx: torch.Tensor, cfg: SmaqConfig):
    """smart.py:86-91 with the permutation prefix ``idx`` passed in (k = min(N, num_samples))."""
    sample = data.view(-1)[idx]
    return sample.mean(), std_of(sample, cfg, unbiased=False)


def std_of(data: torch.Tensor, cfg: SmaqConfig, **kw):
    """smart.py:100-108."""
    if cfg.use_range_std_dev:
        range_ = data.max() - data.min()
        c = 1 / torch.sqrt(2.0 * torch.log(torch.tensor(data.numel()).type_as(range_)))
        return range_ * c
    return data.std(**kw)


def full_mean_std(data: torch.Tensor, cfg: SmaqConfig):
    """smart.py:130-132 (default path: full-tensor mean, unbiased std)."""
    return data.mean(), std_of(data, cfg)


def round_stochastic(c: torch.Tensor, probs: torch.Tensor, rng_rule: bool = False) -> torch.Tensor:
    """smart.py:93-98 with ``probs`` supplied by the caller instead of rand_like.

    ``rng_rule`` (NOT in the reference): the rule the CUDA kernels apply to their OWN random numbers
    (oracle/rng.py): floor(c) + [frac >= p].  It is the reference's expression except where
    |frac - p| <= 2^-25, i.e. where ``(frac - p) + 0.5`` lands on the tie of round-half-even.  With
    caller-supplied ``probs`` (parity mode) the kernels evaluate the reference's expression itself."""
    floored = c.floor()
    fractions = c - floored
    if rng_rule:
        return floored + (fractions >= probs).to(c.dtype)
    return floored + torch.relu((fractions - probs) + 0.5).round()


@torch.no_grad()
def smaq_roundtrip(
    data: torch.Tensor,
    cfg: SmaqConfig = SmaqConfig(),
    *,
    probs: Optional[torch.Tensor] = None,
    idx: Optional[torch.Tensor] = None,
    mean: Optional[torch.Tensor] = None,
    std: Optional[torch.Tensor] = None,
    all_positive: bool = False,
    saturate: bool = False,
    batch_norm_stats=None,
    rng_rule: bool = False,
) -> SmaqResult:
    """The whole fake-quantisation call, smart.py:121-190.

    ``batch_norm_stats`` (gamma, beta) of the producing BatchNorm2d; honoured only under
                 --use_batch_norm (smart.py:121): the NCHW map is un-affined per channel AFTER the
                 statistics were taken (smart.py:136-149) and re-affined before ``all_positive``
                 (smart.py:174-182).
    ``probs``   uniform [0,1) numbers, same shape as data (needed when stochastic_rounding)
    ``idx``     sample indices for --use_sample_stats (the first k entries of the permutation)
    ``mean/std`` override the statistics step (used to feed a kernel's stats back in)
    ``saturate`` NOT in the reference: clamp the rounded code to what the packed
                 format can hold (|code| <= 2^(bits-2)-1).  This is the H1 rule of
                 SURVEY.md §7.3: the reference never clamps at
                 --outlier_std_dev_threshold, a packed byte must.
    """
    numel = data.numel()
    if numel < cfg.min_size:  # smart.py:125-128
        return SmaqResult(y=data, passthrough=True)

    if mean is None or std is None:
        if cfg.use_sample_stats:
            assert idx is not None, "sampled statistics need explicit indices"
            mean, std = sample_mean_std(data, idx, cfg)  # smart.py:133
        else:
            mean, std = full_mean_std(data, cfg)  # smart.py:131
    mean = torch.as_tensor(mean, dtype=data.dtype)
    std_raw = torch.as_tensor(std, dtype=data.dtype)

    use_bn = cfg.use_batch_norm and batch_norm_stats is not None  # smart.py:121
    if use_bn:  # smart.py:136-149
        gamma, beta = batch_norm_stats
        if cfg.bn_scalar_params:
            gamma, beta = gamma.mean(), beta.mean()
        data = ((data.permute(0, 3, 2, 1).clone() - beta) / gamma).permute(0, 3, 2, 1).clone()

    std_dev = std_raw
    if std_dev == 0:  # smart.py:151-152
        std_dev = torch.ones_like(std_dev)

    t = cfg.main_std_dev_threshold
    z = (data - mean) / std_dev.clamp(*cfg.clamped_range)  # smart.py:154
    hi = z > t  # smart.py:155
    lo = z < -t  # smart.py:156
    is_outlier = hi | lo  # smart.py:157
    scalars = (hi * -t) + (lo * t)  # smart.py:159-161
    ranges = torch.where(is_outlier, cfg.range_outlier, cfg.range_normal)  # smart.py:162

    c = (z + scalars) * ranges  # smart.py:164
    if cfg.stochastic_rounding:  # smart.py:166-169
        assert probs is not None, "stochastic rounding needs explicit probs"
        code = round_stochastic(c, probs, rng_rule)
    else:
        code = c.trunc()

    if saturate:
        lim = torch.where(is_outlier, float(cfg.max_code_outlier), float(cfg.max_code_main))
        # NaN codes are left alone (clamp keeps NaN); the packer maps them explicitly
        code = torch.maximum(torch.minimum(code, lim), -lim)

    y = (code / ranges) - scalars  # smart.py:171
    y = (y * std_dev) + mean  # smart.py:172
    if use_bn:  # smart.py:174-179
        y = ((y.permute(0, 3, 2, 1).clone() * gamma) + beta).permute(0, 3, 2, 1).clone()
    if all_positive:  # smart.py:181-182
        y = y.clamp_min(0.0)

    return SmaqResult(y=y, mean=mean, std=std_raw, hi=hi, lo=lo, code=code, extras={"z": z, "c": c, "probs": probs, "rng_rule": rng_rule})


def snap_mean_to_zero(mean: torch.Tensor, std: torch.Tensor, cfg: SmaqConfig) -> torch.Tensor:
    """NOT in the reference — the packed encoder's ``zero_on_grid`` option (csrc/params.cuh): the mean m' nearest to
    ``mean`` for which x == 0 decodes to exactly 0: with c0 the integer code nearest to zero's scaled z-score,
    m' = -(((c0 / range) - shift) * std), each step rounded to fp32 as smart.py:171-172 rounds it."""
    mean = torch.as_tensor(mean, dtype=torch.float32)
    std = torch.as_tensor(std, dtype=torch.float32)
    std_dev = torch.ones_like(std) if std == 0 else std
    t = cfg.main_std_dev_threshold
    z0 = (torch.zeros_like(mean) - mean) / std_dev.clamp(*cfg.clamped_range)
    if not bool(z0.abs() <= 1e30) or not bool(std_dev > 0) or not bool(std_dev.abs() <= 1e30):
        return mean
    hi, lo = z0 > t, z0 < -t
    shift = (hi * -t) + (lo * t)
    rng = torch.where(hi | lo, torch.tensor(cfg.range_outlier), torch.tensor(cfg.range_normal)).to(torch.float32)
    lim = cfg.max_code_outlier if bool(hi | lo) else cfg.max_code_main
    c0 = ((z0 + shift) * rng).round()
    if not bool(c0.abs() <= lim):
        return mean
    return -(((c0 / rng) - shift) * std_dev)


def compressed_bits(res: SmaqResult, cfg: SmaqConfig) -> int:
    """smart.py:184-187: the size the reference reports for one call."""
    n_out = int((res.hi | res.lo).sum())
    n_main = res.hi.numel() - n_out
    return n_out * cfg.num_bits_outlier + n_main * cfg.num_bits_main
