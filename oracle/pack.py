"""CPU oracle for the PACKED SmaQ stream "SQB3" (TEST INFRASTRUCTURE ONLY).

The reference never materialises codes: SmartFP is fake quantisation (smart.py:154-172) and only
*accounts* for a size of 6 bits per main element and 8 per outlier (smart.py:184-187).  The
packed stream is therefore new work named by the north star; its oracle is this bit-packer
applied to the integer codes of oracle/smaq.py (which is pinned to the reference), with the H1
saturation rule of SURVEY.md §7.3.  DESIGN.md "Packed layout" is the normative text; this file
is its executable form and the CUDA encoder must reproduce it byte for byte.

Per element, with L = 2^(bits-2) - 1 the largest code magnitude its width holds
(bits = num_bits_main or num_bits_outlier), pm = bits_main - 1, po = bits_outlier - 1, xb = po - pm:
    c'   = clamp((z + shift) * range, -L, L)            the H1 rule; NaN -> 0 (both counted in
                                                         n_saturated — a pure function of the data)
    code = round(c')                                     the reference's rounding (smart.py:93-98 / :169);
                                                         == clamp(reference code, -L, L)
    S    = code - [sign bit of z]                        z < 0 implies code <= 0 and z >= 0 implies
                                                         code >= 0, so S is a two's-complement number of
                                                         bits-1 bits and (side, code) is recovered from it
    U    = S + 2^(po-1)
    base = U mod 2^pm    stored for every element at a fixed position (main: S in two's complement)
    ext  = U >> pm       xb bits, stored only for outliers, densely   (outlier: U is S, offset binary)
so the stream holds exactly  n + pm*n + xb*n_out = bits_main*n_main + bits_outlier*n_out  bits
(+ word alignment per warp tile, reported as overhead).

Geometry (chosen so one warp reads its 1024 values as one 4 KB bulk copy and every lane packs its own
32 values in registers with packed fp32 arithmetic):
    warp tile = 1024 consecutive elements; lane l (0..31) owns four CHUNKS of eight consecutive
                elements: chunk k (0..3), element j (0..7)  <->  tile element 256*k + 8*l + j
    planes    : per warp tile (1+pm) rows of 32 uint32, row-major [row][lane]:
                row 0 = tag word (bit 8k+j = element (k, j) is an outlier),
                rows 1..pm = the lane's 32 base fields as a 32*pm-bit string, LSB-first; element
                (k, j) sits at bit  8*pm*k + 4*pm*(j & 1) + pm*(j >> 1)  (even elements of a chunk
                first, then the odd ones: the two lanes of the packed fp32 accumulator)
    extras    : per warp tile, lanes in order, per lane its chunks in order (LSB-first); inside a
                chunk the outliers' ext fields are concatenated with the FIRST outlier in the MOST
                significant position (E = E * 2^xb + ext); the warp-tile segment is padded to a
                whole uint32 and starts at a FIXED place, word t * (1024 * xb / 32) of the extras section
                for warp tile t: only its used words (their number follows from the tile's tag words) are
                ever written or read, the rest of the stride is unspecified.  (Round 1's "SQB2" packed the
                segments densely across tiles behind a per-CTA-tile table; that made the encoder two
                passes — see DESIGN.md §3.)
Elements past n (padding of the last tile) are main elements with S = 0.

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
MAGIC = 0x33425153  # 'SQB3' little endian


@dataclass
class Packed:
    n: int
    cfg: SmaqConfig
    mean: np.float32
    std_raw: np.float32
    planes: np.ndarray   # uint32 [n_warp_tiles, 1+pm, 32]
    extras: np.ndarray   # uint32 [n_warp_tiles, seg_words]; words past seg_used[t] are zero here, unspecified on the device
    seg_used: np.ndarray  # int64 [n_warp_tiles] words of each segment that are part of the stream
    n_outlier: int
    n_saturated: int

    @property
    def payload_bits(self) -> int:
        """The size the reference accounts for (smart.py:184-187)."""
        c = self.cfg
        return c.num_bits_outlier * self.n_outlier + c.num_bits_main * (self.n - self.n_outlier)

    @property
    def extras_words(self) -> int:
        return int(self.seg_used.sum())

    @property
    def stored_bits(self) -> int:
        return 32 * (self.planes.size + self.extras_words)


def lane_order_index(n_padded: int) -> np.ndarray:
    """perm[w, l, i] = flat element index held by lane l, local slot i of warp tile w."""
    w = np.arange(n_padded // WARP_TILE)[:, None, None]
    l = np.arange(32)[None, :, None]
    i = np.arange(32)[None, None, :]
    k, j = i // 8, i % 8
    return w * WARP_TILE + 256 * k + 8 * l + j


def field_bit_positions(pm: int) -> np.ndarray:
    """Bit position of local element i = 8k + j inside the lane's 32*pm-bit base string."""
    i = np.arange(32)
    k, j = i // 8, i % 8
    return 8 * pm * k + 4 * pm * (j & 1) + pm * (j >> 1)


def stored_values(res: SmaqResult, cfg: SmaqConfig):
    """(outlier mask, S, n_saturated) from an UNSATURATED oracle result (needs extras c, z)."""
    from .smaq import round_stochastic

    c = res.extras["c"].detach().reshape(-1).clone()
    z = res.extras["z"].detach().reshape(-1)
    outlier = (res.hi | res.lo).reshape(-1)
    lim = torch.where(outlier, float(cfg.max_code_outlier), float(cfg.max_code_main))
    isn = torch.isnan(c)
    n_sat = int((isn | (c.abs() > lim)).sum())
    cc = torch.where(isn, torch.zeros_like(c), torch.maximum(torch.minimum(c, lim), -lim))
    if cfg.stochastic_rounding:
        code = round_stochastic(cc, res.extras["probs"].reshape(-1), res.extras.get("rng_rule", False))
    else:
        code = cc.trunc()
    # where nothing was clipped this IS the reference's code (pinned), elsewhere its clamp
    ref = res.code.detach().reshape(-1)
    ok = ~isn & ~torch.isnan(ref)
    assert torch.equal(code[ok], torch.maximum(torch.minimum(ref, lim), -lim)[ok]), "clamp(c) then round != clamp(code)"
    below = torch.signbit(z) & ~isn
    S = code.to(torch.int64) - below.to(torch.int64)
    return outlier.numpy(), S.numpy(), n_sat


def pack(res: SmaqResult, cfg: SmaqConfig) -> Packed:
    pm, po = cfg.num_bits_main - 1, cfg.num_bits_outlier - 1
    xb = po - pm
    assert pm >= 2 and xb >= 0
    n = res.code.numel()
    outlier, S, n_sat = stored_values(res, cfg)
    lo_m, hi_m = -(cfg.max_code_main + 1), cfg.max_code_main
    lo_o, hi_o = -(cfg.max_code_outlier + 1), cfg.max_code_outlier
    assert np.all((S[~outlier] >= lo_m) & (S[~outlier] <= hi_m)) and np.all((S[outlier] >= lo_o) & (S[outlier] <= hi_o))
    U = (S + (1 << (po - 1))).astype(np.uint64)

    n_wt = -(-n // WARP_TILE)
    n_pad = n_wt * WARP_TILE
    upad = np.full(n_pad, 1 << (po - 1), dtype=np.uint64)   # padding: main, S = 0
    tag = np.zeros(n_pad, dtype=bool)
    upad[:n] = U
    tag[:n] = outlier
    perm = lane_order_index(n_pad)          # [w, l, i]
    u_l = upad[perm]
    tag_l = tag[perm]

    planes = np.zeros((n_wt, 1 + pm, 32), dtype=np.uint32)
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    planes[:, 0, :] = (tag_l.astype(np.uint64) * weights).sum(axis=2).astype(np.uint32)
    base = u_l & np.uint64((1 << pm) - 1)
    # the 32 pm-bit fields of each lane go into its 32*pm-bit string (pm words, LSB-first)
    words = np.zeros((n_wt, 32, pm + 1), dtype=np.uint64)
    pos = field_bit_positions(pm)
    for i in range(32):
        wi, sh = int(pos[i]) >> 5, int(pos[i]) & 31
        v = base[:, :, i] << np.uint64(sh)
        words[:, :, wi] |= v & np.uint64(0xFFFFFFFF)
        words[:, :, wi + 1] |= v >> np.uint64(32)
    assert not words[:, :, pm].any()
    planes[:, 1:, :] = np.transpose(words[:, :, :pm].astype(np.uint32), (0, 2, 1))

    # extras: one word-aligned segment per warp tile, at a fixed stride
    seg_w = WARP_TILE * xb // 32
    extras = np.zeros((n_wt, max(seg_w, 1)), dtype=np.uint32)
    seg_used = np.zeros(n_wt, dtype=np.int64)
    if xb:
        t4 = tag_l.reshape(n_wt, 32, 4, 8)
        per_chunk = t4.sum(axis=3).reshape(n_wt, 128)                       # outliers per (lane, chunk), stream order
        before = np.cumsum(per_chunk, axis=1) - per_chunk                    # ... in the chunks before this one
        after = np.flip(np.cumsum(np.flip(t4, axis=3), axis=3), axis=3) - t4  # ... after this element in its chunk
        slot = (before.reshape(n_wt, 32, 4, 1) + after).reshape(n_wt, 32, 32)  # first outlier of a chunk: highest slot
        ext = (u_l >> np.uint64(pm)).astype(np.int64)
        w_idx, l_idx, i_idx = np.nonzero(tag_l)
        bitpos = slot[w_idx, l_idx, i_idx] * xb
        val = ext[w_idx, l_idx, i_idx]
        for t in range(xb):
            bp = bitpos + t
            np.bitwise_or.at(extras, (w_idx, bp >> 5), (((val >> t) & 1) << (bp & 31)).astype(np.uint32))
        seg_used = -(-(tag_l.reshape(n_wt, -1).sum(axis=1) * xb) // 32)
    return Packed(n=n, cfg=cfg, mean=np.float32(res.mean), std_raw=np.float32(res.std), planes=planes,
                  extras=extras, seg_used=seg_used.astype(np.int64), n_outlier=int(outlier.sum()), n_saturated=n_sat)


def unpack_values(p: Packed):
    """Packed -> (outlier mask, S) in flat element order."""
    cfg = p.cfg
    pm, po = cfg.num_bits_main - 1, cfg.num_bits_outlier - 1
    xb = po - pm
    n_wt = p.planes.shape[0]
    n_pad = n_wt * WARP_TILE
    shifts = np.arange(32, dtype=np.uint32)
    tag_l = ((p.planes[:, 0, :, None] >> shifts) & 1).astype(bool)          # [w, l, i]
    words = np.transpose(p.planes[:, 1:, :], (0, 2, 1)).astype(np.uint64)    # [w, l, pm]
    words = np.concatenate([words, np.zeros((n_wt, 32, 1), dtype=np.uint64)], axis=2)
    pos = field_bit_positions(pm)
    base = np.zeros((n_wt, 32, 32), dtype=np.uint64)
    for i in range(32):
        wi, sh = int(pos[i]) >> 5, int(pos[i]) & 31
        two = words[:, :, wi] | (words[:, :, wi + 1] << np.uint64(32))
        base[:, :, i] = (two >> np.uint64(sh)) & np.uint64((1 << pm) - 1)
    ext = np.zeros((n_wt, 32, 32), dtype=np.uint64)
    if xb:
        t4 = tag_l.reshape(n_wt, 32, 4, 8)
        per_chunk = t4.sum(axis=3).reshape(n_wt, 128)
        before = np.cumsum(per_chunk, axis=1) - per_chunk
        after = np.flip(np.cumsum(np.flip(t4, axis=3), axis=3), axis=3) - t4
        slot = (before.reshape(n_wt, 32, 4, 1) + after).reshape(n_wt, 32, 32)
        used = -(-(tag_l.reshape(n_wt, -1).sum(axis=1) * xb) // 32)
        assert np.array_equal(used, p.seg_used), "segment sizes do not match the tag words"
        w_idx, l_idx, i_idx = np.nonzero(tag_l)
        bitpos = slot[w_idx, l_idx, i_idx] * xb
        v = np.zeros(w_idx.shape, dtype=np.uint64)
        for t in range(xb):
            bp = bitpos + t
            v |= ((p.extras[w_idx, bp >> 5].astype(np.uint64) >> (bp & 31).astype(np.uint64)) & np.uint64(1)) << np.uint64(t)
        ext[w_idx, l_idx, i_idx] = v
    u_l = (base | (ext << np.uint64(pm))).reshape(-1)
    tflat = tag_l.reshape(-1)
    perm = lane_order_index(n_pad).reshape(-1)
    U = np.zeros(n_pad, dtype=np.int64)
    tag = np.zeros(n_pad, dtype=bool)
    U[perm] = u_l.astype(np.int64)
    tag[perm] = tflat
    U, tag = U[: p.n], tag[: p.n]
    S_out = U - (1 << (po - 1))
    if xb > 0:   # main: two's complement on pm bits
        S_main = np.where(U >= (1 << (pm - 1)), U - (1 << pm), U)
    else:
        S_main = U - (1 << (pm - 1))
    return tag, np.where(tag, S_out, S_main)


@torch.no_grad()
def decode(p: Packed, all_positive: bool = False) -> torch.Tensor:
    """Packed -> fp32, finishing with the reference's own inverse (smart.py:171-172,181-182)."""
    cfg = p.cfg
    tag, S = unpack_values(p)
    tag_t = torch.from_numpy(tag)
    S_t = torch.from_numpy(S)
    below = S_t < 0
    code = torch.where(below, S_t + 1, S_t).to(torch.float32)
    if not cfg.stochastic_rounding:   # trunc yields -0.0 for a zero code on the lower side
        code = torch.where(below & (code == 0), torch.tensor(-0.0), code)
    t = cfg.main_std_dev_threshold
    hi = tag_t & ~below
    lo = tag_t & below
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
