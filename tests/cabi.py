"""Thin test-side helpers that call the C ABI (include/smaq_b200.h) directly through ctypes."""
import ctypes as C

import torch

from smart_compress import _native as N


def codec_params(cfg, *, all_positive=False, saturate=False, seed=1234, offset=0, count_saturated=True,
                 zero_on_grid=False) -> N.CodecParams:
    """oracle SmaqConfig -> smaq_codec_params (the host-side float->fp32 narrowing happens in ctypes)."""
    p = N.CodecParams()
    p.threshold = cfg.main_std_dev_threshold
    p.range_main = cfg.range_normal
    p.range_outlier = cfg.range_outlier
    p.clamp_lo, p.clamp_hi = cfg.clamped_range
    p.bits_main, p.bits_outlier = cfg.num_bits_main, cfg.num_bits_outlier
    p.stochastic = int(cfg.stochastic_rounding)
    p.all_positive = int(all_positive)
    p.saturate = int(saturate)
    p.count_saturated = int(count_saturated)
    p.zero_on_grid = int(zero_on_grid)
    p.seed, p.offset = seed, offset
    return p


def _ws(n, dev):
    lib = N.load()
    nbytes = lib.smaq_stats_workspace_bytes(n)
    return torch.empty(nbytes, dtype=torch.uint8, device=dev), nbytes


def stats_full(x, unbiased=True):
    lib = N.load()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    ws, nb = _ws(x.numel(), x.device)
    N.check(lib.smaq_stats_full(x.data_ptr(), x.numel(), int(unbiased), out.data_ptr(), ws.data_ptr(), nb,
                                N.stream_ptr(x.device)), "stats_full")
    return out


def stats_range(x):
    lib = N.load()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    ws, nb = _ws(x.numel(), x.device)
    N.check(lib.smaq_stats_range(x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), nb,
                                 N.stream_ptr(x.device)), "stats_range")
    return out


def stats_sampled(x, idx, range_std=False):
    lib = N.load()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    idx = idx.to(x.device, torch.int64).contiguous()
    N.check(lib.smaq_stats_sampled(x.data_ptr(), x.numel(), idx.data_ptr(), idx.numel(), int(range_std), out.data_ptr(),
                                   N.stream_ptr(x.device)), "stats_sampled")
    return out


def stats_sampled_draw(x, k, seed, offset=0, range_std=False):
    lib = N.load()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    N.check(lib.smaq_stats_sampled_draw(x.data_ptr(), x.numel(), k, int(range_std), seed, offset, out.data_ptr(),
                                        N.stream_ptr(x.device)), "stats_sampled_draw")
    return out


def s2fp8_stats(x):
    lib = N.load()
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    ws, nb = _ws(x.numel(), x.device)
    N.check(lib.smaq_s2fp8_stats(x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), nb,
                                 N.stream_ptr(x.device)), "s2fp8_stats")
    return out


def roundtrip(x, mean_std, params, probs=None, out=None):
    lib = N.load()
    y = torch.empty_like(x) if out is None else out
    N.check(lib.smaq_roundtrip(x.data_ptr(), y.data_ptr(), x.numel(), mean_std.data_ptr(),
                               None if probs is None else probs.data_ptr(), C.byref(params),
                               N.stream_ptr(x.device)), "roundtrip")
    return y


def roundtrip_bn(x, mean_std, params, gamma, beta, channels, inner, probs=None):
    """smaq_roundtrip_bn on a flat NCHW tensor (--use_batch_norm)."""
    lib = N.load()
    y = torch.empty_like(x)
    N.check(lib.smaq_roundtrip_bn(x.data_ptr(), y.data_ptr(), x.numel(), mean_std.data_ptr(),
                                  None if probs is None else probs.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                  channels, inner, C.byref(params), N.stream_ptr(x.device)), "roundtrip_bn")
    return y


def compress(x, params, probs=None, out=None, ws=None):
    """smaq_compress (statistics + round trip behind one entry point) -> (y, the mean/std it used)."""
    lib = N.load()
    n = x.numel()
    y = torch.empty_like(x) if out is None else out
    need = lib.smaq_compress_workspace_bytes(n)
    if ws is None:
        ws = torch.full((need,), 0xA5, dtype=torch.uint8, device=x.device)   # dirty scratch
        N.check(lib.smaq_compress_workspace_init(ws.data_ptr(), ws.numel(), N.stream_ptr(x.device)), "init")
    N.check(lib.smaq_compress(x.data_ptr(), y.data_ptr(), n, None if probs is None else probs.data_ptr(),
                              C.byref(params), ws.data_ptr(), ws.numel(), N.stream_ptr(x.device)), "compress")
    sb = lib.smaq_stats_workspace_bytes(n)
    return y, ws[sb:sb + 8].view(torch.float32).clone()


def roundtrip_small(x, params, probs=None, want_stats=False):
    lib = N.load()
    y = torch.empty_like(x)
    ms = torch.empty(2, dtype=torch.float32, device=x.device) if want_stats else None
    N.check(lib.smaq_roundtrip_small(x.data_ptr(), y.data_ptr(), x.numel(),
                                     None if probs is None else probs.data_ptr(), C.byref(params),
                                     None if ms is None else ms.data_ptr(), N.stream_ptr(x.device)), "roundtrip_small")
    return (y, ms) if want_stats else y


def floatq_params(exp, man, *, rounding=1, check_inf=True, max_exp_bias=0, seed=7, offset=0):
    p = N.FloatqParams()
    p.exp_bits, p.man_bits, p.rounding = exp, man, rounding
    p.check_inf, p.max_exp_bias = int(check_inf), max_exp_bias
    p.seed, p.offset = seed, offset
    return p


def float_quantize(x, params, rand_bits=None, out=None):
    lib = N.load()
    y = torch.empty_like(x) if out is None else out
    N.check(lib.smaq_float_quantize(x.data_ptr(), y.data_ptr(), x.numel(),
                                    None if rand_bits is None else rand_bits.data_ptr(), C.byref(params),
                                    N.stream_ptr(x.device)), "float_quantize")
    return y


def s2fp8_apply(x, mu_max, params, rand_bits=None):
    lib = N.load()
    y = torch.empty_like(x)
    N.check(lib.smaq_s2fp8_apply(x.data_ptr(), y.data_ptr(), x.numel(), mu_max.data_ptr(),
                                 None if rand_bits is None else rand_bits.data_ptr(), C.byref(params),
                                 N.stream_ptr(x.device)), "s2fp8_apply")
    return y


def selftest_pow(a, y):
    """(a ** y, accepted-per-quad) through the S2FP8 fast path's pow (test hook)."""
    lib = N.load()
    out = torch.empty_like(a)
    acc = torch.empty((a.numel() + 3) // 4, dtype=torch.int32, device=a.device)
    yt = torch.tensor([float(y)], dtype=torch.float32, device=a.device)
    N.check(lib.smaq_selftest_pow(a.data_ptr(), yt.data_ptr(), out.data_ptr(), acc.data_ptr(), a.numel(),
                                  N.stream_ptr(a.device)), "selftest_pow")
    return out, acc


def selftest_s2_screen(samples, dev="cuda:0"):
    """[lg2 abs err, lg2 rel err, ex2 rel err, bound ratio, largest margin, samples used] (test hook)."""
    out = torch.zeros(6, dtype=torch.float32, device=dev)
    N.check(N.load().smaq_selftest_s2_screen(out.data_ptr(), samples, N.stream_ptr(out.device)), "selftest_s2_screen")
    torch.cuda.synchronize()
    return out.cpu().tolist()


def mean_std_tensor(mean, std, dev):
    return torch.tensor([float(mean), float(std)], dtype=torch.float32, device=dev)
