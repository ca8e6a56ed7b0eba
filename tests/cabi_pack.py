"""Test-side helpers for the packed stream: call smaq_encode / smaq_decode through the C ABI and
compare with oracle/pack.py."""
import ctypes as C

import numpy as np
import torch

from oracle import pack as opack
from oracle.smaq import smaq_roundtrip
from smart_compress import _native as N
from tests import cabi


def layout(n, cfg):
    lay = N.PackedLayout()
    N.check(N.load().smaq_packed_layout_for(n, cfg.num_bits_main, cfg.num_bits_outlier, C.byref(lay)), "layout")
    return lay


def encode(xd, ms, params, cfg, probs=None):
    lib = N.load()
    lay = layout(xd.numel(), cfg)
    buf = torch.full((lay.total_capacity_bytes,), 0xAB, dtype=torch.uint8, device=xd.device)  # poison
    ws = torch.full((lay.workspace_bytes,), 0x5A, dtype=torch.uint8, device=xd.device)   # dirty scratch
    N.check(lib.smaq_encode_workspace_init(ws.data_ptr(), ws.numel(), N.stream_ptr(xd.device)), "encode ws init")
    N.check(lib.smaq_encode(xd.data_ptr(), xd.numel(), ms.data_ptr(), None if probs is None else probs.data_ptr(),
                            C.byref(params), buf.data_ptr(), buf.numel(), ws.data_ptr(), ws.numel(),
                            N.stream_ptr(xd.device)), "encode")
    assert not bool(ws.any()), "smaq_encode must leave its scratch zero"
    return buf, lay


def decode(buf, lay, all_positive=False, out=None):
    lib = N.load()
    y = torch.empty(lay.n, dtype=torch.float32, device=buf.device) if out is None else out
    N.check(lib.smaq_decode(buf.data_ptr(), buf.numel(), lay.n, lay.bits_main, lay.bits_outlier, int(all_positive),
                            y.data_ptr(), N.stream_ptr(buf.device)), "decode")
    return y


def sections(buf, lay):
    """Host copies: (header struct, planes uint32[n_wt, 1+pm, 32], extras uint32[n_wt, stride words])."""
    raw = buf.cpu().numpy()
    hdr = N.PackedHeader.from_buffer_copy(bytes(raw[: C.sizeof(N.PackedHeader)]))
    pm = lay.bits_main - 1
    planes = raw[lay.planes_off: lay.planes_off + lay.planes_bytes].view(np.uint32).reshape(lay.n_warp_tiles, 1 + pm, 32).copy()
    sw = max(lay.extras_stride_bytes // 4, 1)
    if lay.extras_stride_bytes:
        extras = raw[lay.extras_off: lay.extras_off + lay.n_warp_tiles * lay.extras_stride_bytes].view(np.uint32).reshape(lay.n_warp_tiles, sw).copy()
    else:
        extras = np.zeros((lay.n_warp_tiles, 1), dtype=np.uint32)
    return hdr, planes, extras


def assert_stream_equals_oracle(buf, lay, p: opack.Packed, poison=0xABABABAB):
    """Every byte of the stream: header counters, planes, each segment's used words — and nothing written outside
    them (the encode helper poisons the buffer first)."""
    hdr, planes, extras = sections(buf, lay)
    assert hdr.status == 0 and hdr.magic == opack.MAGIC and hdr.n == p.n
    assert np.array_equal(planes[:, 0, :], p.planes[:, 0, :]), "outlier bitmap differs"
    assert np.array_equal(planes, p.planes), "base fields differ"
    used = np.arange(extras.shape[1])[None, :] < p.seg_used[:, None]
    assert np.array_equal(np.where(used, extras, 0), np.where(used, p.extras, 0)), "extras segments differ"
    if lay.extras_stride_bytes:
        assert np.all(extras[~used] == poison), "the encoder wrote outside a segment's used words"
    assert hdr.n_outlier == p.n_outlier and hdr.extras_words == p.extras_words
    return hdr


def upload_oracle_packed(p: opack.Packed, lay, dev):
    """Serialise an oracle Packed into the on-device layout (header included)."""
    raw = np.zeros(lay.total_capacity_bytes, dtype=np.uint8)
    h = N.PackedHeader()
    cfg = p.cfg
    h.magic, h.bits_main, h.bits_outlier = opack.MAGIC, cfg.num_bits_main, cfg.num_bits_outlier
    h.stochastic, h.n = int(cfg.stochastic_rounding), p.n
    h.mean, h.std_raw = float(p.mean), float(p.std_raw)
    h.threshold, h.range_main, h.range_outlier = cfg.main_std_dev_threshold, cfg.range_normal, cfg.range_outlier
    h.clamp_lo, h.clamp_hi = cfg.clamped_range
    h.n_outlier, h.n_saturated, h.extras_words = p.n_outlier, p.n_saturated, p.extras_words
    raw[: C.sizeof(h)] = np.frombuffer(bytes(h), dtype=np.uint8)
    raw[lay.planes_off: lay.planes_off + p.planes.size * 4] = p.planes.reshape(-1).view(np.uint8)
    if lay.extras_stride_bytes:
        # unused words of a segment are unspecified: fill them with noise so the decoder cannot rely on zeros
        ex = p.extras.copy()
        noise = np.random.default_rng(0).integers(0, 2**32, ex.shape, dtype=np.uint64).astype(np.uint32)
        used = np.arange(ex.shape[1])[None, :] < p.seg_used[:, None]
        ex = np.where(used, ex, noise)
        raw[lay.extras_off: lay.extras_off + ex.size * 4] = ex.reshape(-1).view(np.uint8)
    return torch.from_numpy(raw).to(dev)


def smoke_check(xd, ms, probs_d, cfg, x, probs):
    """Used by __graft_entry__.smoke(): encode -> decode on the GPU against the oracle."""
    params = cabi.codec_params(cfg)
    buf, lay = encode(xd, ms, params, cfg, probs=probs_d)
    y = decode(buf, lay)
    msc = ms.cpu()
    res = smaq_roundtrip(x, cfg, probs=probs, mean=msc[0], std=msc[1])
    want = smaq_roundtrip(x, cfg, probs=probs, mean=msc[0], std=msc[1], saturate=True)
    assert torch.equal(y.cpu().view(torch.int32), want.y.view(torch.int32)), "packed decode != oracle"
    p = opack.pack(res, cfg)
    hdr = assert_stream_equals_oracle(buf, lay, p)
    assert hdr.n_saturated == p.n_saturated
