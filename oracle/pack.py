"""CPU oracle for the PACKED SmaQ stream "SQB1" (TEST INFRASTRUCTURE ONLY).

The reference never materialises codes: SmartFP is fake quantisation (smart.py:154-172) and only
*accounts* for a size of 6 bits per main element and 8 per outlier (smart.py:184-187).  The
packed stream is therefore new work named by the north star; its oracle is this bit-packer
applied to the integer codes of oracle/smaq.py (which is pinned to the reference), with the H1
saturation rule of SURVEY.md §7.3.  DESIGN.md "Packed layout" is the normative text; this file
is its executable form and the CUDA encoder must reproduce it byte for byte.

Per element the code is split as
    payload P = (|code| << 1) | s      s = sign bit of the code (main)  /  1 for a LOWER outlier
    main   : P < 2^pm,  pm = bits_main - 1        (tag bit + pm bits   = bits_main)
    outlier: P < 2^po,  po = bits_outlier - 1     (tag bit + po bits   = bits_outlier)
    base = P & (2^pm - 1)   stored for every element at a fixed position
    ext  = P >> pm          xb = po - pm bits, stored only for outliers, densely
so the stream holds exactly  n + pm*n + xb*n_out = bits_main*n_main + bits_outlier*n_out  bits
(+ the per-tile table and word alignment, reported as overhead).

Geometry (chosen so one warp reads its 1024 values with eight coalesced 128-bit loads and every
lane packs its own 32 values in registers):
    warp tile = 1024 consecutive elements; lane l (0..31) owns local element i = 8k + j
                (k = 0..3, j = 0..7)  <->  tile element 256*k + 8*l + j
    planes    : per warp tile (1+pm) rows of 32 uint32, row-major [row][lane]:
                row 0 = tag word (bit i = element i is an outlier),
                rows 1..pm = the lane's 32 base fields concatenated LSB-first (field i at bit pm*i)
    extras    : per warp tile, its outliers' ext fields concatenated in (lane, i) order, LSB-first,
                padded to a whole uint32 (so a warp places them without a block barrier); warp-tile
                segments follow each other in tile order.
    CTA tile  = 8 warp tiles; table[t] = first word of CTA tile t's first segment in the extras
                section, table[n_cta_tiles] = total words (a warp finds its own segment by adding
                the word counts of the warps before it, which follow from the tag words).
Elements past n (padding of the last tile) are main elements with payload 0.
Infinite codes saturate like any other over-range code; NaN codes (NaN statistics or inputs) are
stored as code 0; both are counted in n_saturated together with the clamped ones.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .smaq import SmaqConfig, SmaqResult

WARP_TILE = 1024
WARPS_PER_CTA = 8
CTA_TILE = WARP_TILE * WARPS_PER_CTA
MAGIC = 0x31425153  # 'SQB1' little endian


@dataclass
class Packed:
    n: int
    cfg: SmaqConfig
    mean: np.float32
    std_raw: np.float32
    planes: np.ndarray   # uint32 [n_warp_tiles, 1+pm, 32]
    table: np.ndarray    # uint32 [n_cta_tiles + 1]
    extras: np.ndarray   # uint32 [table[-1]]
    n_outlier: int
    n_saturated: int

    @property
    def payload_bits(self) -> int:
        """The size the reference accounts for (smart.py:184-187)."""
        c = self.cfg
        return c.num_bits_outlier * self.n_outlier + c.num_bits_main * (self.n - self.n_outlier)

    @property
    def stored_bits(self) -> int:
        return 32 * (self.planes.size + self.table.size + self.extras.size)


def lane_order_index(n_padded: int) -> np.ndarray:
    """perm[w, l, i] = flat element index held by lane l, local slot i of warp tile w."""
    w = np.arange(n_padded // WARP_TILE)[:, None, None]
    l = np.arange(32)[None, :, None]
    i = np.arange(32)[None, None, :]
    k, j = i // 8, i % 8
    return w * WARP_TILE + 256 * k + 8 * l + j


def codes_from_result(res: SmaqResult, cfg: SmaqConfig):
    """(outlier mask, sign/side bit, magnitude, n_saturated) from an UNSATURATED oracle result."""
    code = res.code.detach().reshape(-1).numpy().astype(np.float32)
    hi = res.hi.reshape(-1).numpy()
    lo = res.lo.reshape(-1).numpy()
    outlier = hi | lo
    lim = np.where(outlier, cfg.max_code_outlier, cfg.max_code_main).astype(np.float32)
    finite = ~np.isnan(code)  # +-inf saturates to +-lim, NaN becomes 0
    clipped = finite & (np.abs(code) > lim)
    sat = np.clip(np.where(finite, code, 0.0), -lim, lim)
    mag = np.abs(sat).astype(np.uint32)
    # main: the code's own sign bit (so trunc's -0.0 survives); outlier: which side of the mean
    s = np.where(outlier, lo, np.signbit(sat) & finite).astype(np.uint32)
    return outlier, s, mag, int(clipped.sum() + (~finite).sum())


def pack(res: SmaqResult, cfg: SmaqConfig) -> Packed:
    pm, po = cfg.num_bits_main - 1, cfg.num_bits_outlier - 1
    xb = po - pm
    assert pm >= 2 and xb >= 0
    n = res.code.numel()
    outlier, s, mag, n_sat = codes_from_result(res, cfg)
    payload = (mag.astype(np.uint64) << np.uint64(1)) | s.astype(np.uint64)
    assert np.all(payload[~outlier] < (1 << pm)) and np.all(payload < (1 << po))

    n_wt = -(-n // WARP_TILE)
    n_ct = -(-n_wt // WARPS_PER_CTA)
    n_pad = n_wt * WARP_TILE
    pay = np.zeros(n_pad, dtype=np.uint64)
    tag = np.zeros(n_pad, dtype=bool)
    pay[:n] = payload
    tag[:n] = outlier
    perm = lane_order_index(n_pad)          # [w, l, i]
    pay_l = pay[perm]
    tag_l = tag[perm]

    planes = np.zeros((n_wt, 1 + pm, 32), dtype=np.uint32)
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    planes[:, 0, :] = (tag_l.astype(np.uint64) * weights).sum(axis=2).astype(np.uint32)
    base = pay_l & np.uint64((1 << pm) - 1)
    # concatenate 32 pm-bit fields LSB-first into pm words, via a bit matrix
    bits = ((base[..., None] >> np.arange(pm, dtype=np.uint64)) & np.uint64(1)).reshape(n_wt, 32, 32 * pm)
    words = (bits.reshape(n_wt, 32, pm, 32).astype(np.uint64) * weights).sum(axis=3).astype(np.uint32)
    planes[:, 1:, :] = np.transpose(words, (0, 2, 1))

    # extras: one word-aligned segment per warp tile, fields in (lane, i) order == flattened [w, l, i]
    ext = (pay_l >> np.uint64(pm)).reshape(n_wt, -1)
    tflat = tag_l.reshape(n_wt, -1)
    seg_words = np.zeros(n_ct * WARPS_PER_CTA, dtype=np.int64)
    chunks = []
    for w in range(n_wt):
        e = ext[w][tflat[w]]
        nbits = e.size * xb
        nwords = -(-nbits // 32)
        seg_words[w] = nwords
        if nwords:
            b = ((e[:, None] >> np.arange(xb, dtype=np.uint64)) & np.uint64(1)).reshape(-1)
            b = np.concatenate([b, np.zeros(nwords * 32 - nbits, dtype=np.uint64)])
            chunks.append((b.reshape(nwords, 32) * weights).sum(axis=1).astype(np.uint32))
    table = np.zeros(n_ct + 1, dtype=np.uint32)
    table[1:] = np.cumsum(seg_words.reshape(n_ct, WARPS_PER_CTA).sum(axis=1))
    extras = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint32)
    return Packed(n=n, cfg=cfg, mean=np.float32(res.mean), std_raw=np.float32(res.std), planes=planes,
                  table=table, extras=extras, n_outlier=int(outlier.sum()), n_saturated=n_sat)


def unpack_codes(p: Packed):
    """Packed -> (outlier mask, s bit, magnitude) in flat element order."""
    cfg = p.cfg
    pm, po = cfg.num_bits_main - 1, cfg.num_bits_outlier - 1
    xb = po - pm
    n_wt = p.planes.shape[0]
    n_pad = n_wt * WARP_TILE
    shifts = np.arange(32, dtype=np.uint32)
    tag_l = ((p.planes[:, 0, :, None] >> shifts) & 1).astype(bool)          # [w, l, i]
    words = np.transpose(p.planes[:, 1:, :], (0, 2, 1))                      # [w, l, pm]
    bits = ((words[..., None] >> shifts) & 1).reshape(n_wt, 32, 32 * pm)     # LSB-first bit string
    fields = bits.reshape(n_wt, 32, 32, pm).astype(np.uint64)
    base = (fields << np.arange(pm, dtype=np.uint64)).sum(axis=3)            # [w, l, i]
    ext = np.zeros(n_wt * 32 * 32, dtype=np.uint64)
    tflat = tag_l.reshape(-1)
    word = 0
    for w in range(n_wt):
        sl = slice(w * WARP_TILE, (w + 1) * WARP_TILE)
        k = int(tflat[sl].sum())
        nwords = -(-(k * xb) // 32)
        if w % WARPS_PER_CTA == 0:
            assert word == int(p.table[w // WARPS_PER_CTA]), "table does not match the tag words"
        if k and xb:
            seg = p.extras[word: word + nwords]
            b = ((seg[:, None] >> shifts) & 1).reshape(-1)[: k * xb].reshape(k, xb).astype(np.uint64)
            ext[np.nonzero(tflat[sl])[0] + sl.start] = (b << np.arange(xb, dtype=np.uint64)).sum(axis=1)
        word += nwords
    pay_l = base.reshape(-1) | (ext << np.uint64(pm))
    perm = lane_order_index(n_pad).reshape(-1)
    pay = np.zeros(n_pad, dtype=np.uint64)
    tag = np.zeros(n_pad, dtype=bool)
    pay[perm] = pay_l
    tag[perm] = tflat
    pay, tag = pay[: p.n], tag[: p.n]
    return tag, (pay & np.uint64(1)).astype(np.uint32), (pay >> np.uint64(1)).astype(np.uint32)


@torch.no_grad()
def decode(p: Packed, all_positive: bool = False) -> torch.Tensor:
    """Packed -> fp32, finishing with the reference's own inverse (smart.py:171-172,181-182)."""
    cfg = p.cfg
    tag, s, mag = unpack_codes(p)
    tag_t = torch.from_numpy(tag)
    s_t = torch.from_numpy(s.astype(np.int64)).bool()
    magf = torch.from_numpy(mag.astype(np.float32))
    code = torch.where(s_t, -magf, magf)  # -0.0 when s and mag == 0
    t = cfg.main_std_dev_threshold
    hi = tag_t & ~s_t
    lo = tag_t & s_t
    scalars = (hi * -t) + (lo * t)
    ranges = torch.where(tag_t, cfg.range_outlier, cfg.range_normal)
    mean = torch.tensor(p.mean)
    std = torch.tensor(p.std_raw)
    if std == 0:
        std = torch.ones_like(std)
    y = (code / ranges) - scalars
    y = (y * std) + mean
    if all_positive:
        y = y.clamp_min(0.0)
    return y
