"""The materialised SmaQ stream ("SQB3"): what ``SmartFP.encode`` returns and ``SmartFP.decode`` reads.

Not in the reference (which only fake-quantises, smart.py:154-172); named by the build's north
star.  The buffer lives in device memory; ``header()`` is the only call that synchronises."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from .. import _native as N


def packed_layout(n: int, bits_main: int, bits_outlier: int) -> N.PackedLayout:
    lay = N.PackedLayout()
    N.check(N.load().smaq_packed_layout_for(n, bits_main, bits_outlier, C.byref(lay)), "smaq_packed_layout_for")
    return lay


@dataclass
class PackedSmaq:
    """``buffer`` alone: the whole stream in one capacity-sized allocation (header | planes | fixed-stride extras).
    With ``extras`` set the stream is SPLIT: ``buffer`` = header + planes (exact), ``extras`` a separate allocation —
    capacity-sized and fixed-stride right after ``SmartFP.encode(split=True)``, exact and dense (with ``table``: the
    word offset of every warp tile's segment) after ``SmartFP.compact``."""
    buffer: torch.Tensor          # uint8, device
    layout: N.PackedLayout
    shape: torch.Size
    extras: torch.Tensor = None   # uint8, device (split form)
    table: torch.Tensor = None    # int32[n_warp_tiles + 1], device (compacted form)

    @property
    def numel(self) -> int:
        return int(self.layout.n)

    def header(self) -> N.PackedHeader:
        """Copies the 128-byte header to the host (synchronises the stream)."""
        raw = bytes(self.buffer[: C.sizeof(N.PackedHeader)].cpu().numpy())
        return N.PackedHeader.from_buffer_copy(raw)

    def section(self, name: str) -> torch.Tensor:
        """uint32 view of 'planes' or 'extras' (device tensor; the extras hold one fixed-stride segment per warp
        tile, of which only the used words are specified)."""
        lay = self.layout
        off, nbytes = {
            "planes": (lay.planes_off, lay.planes_bytes),
            "extras": (lay.extras_off, lay.n_warp_tiles * lay.extras_stride_bytes),
        }[name]
        return self.buffer[off: off + nbytes].view(torch.int32)

    def allocated_bytes(self) -> int:
        """What the stream occupies in device memory right now."""
        return sum(int(t.numel() * t.element_size()) for t in (self.buffer, self.extras, self.table) if t is not None)

    def used_bytes(self) -> int:
        """Bytes that are part of the stream: header + planes + the extras words actually written (the buffer
        itself is capacity-sized: every warp tile's segment has a fixed place)."""
        lay = self.layout
        return int(lay.extras_off + 4 * self.header().extras_words)

    def payload_bits(self) -> int:
        """The reference's own size accounting (smart.py:184-187)."""
        h = self.header()
        return int(h.bits_outlier * h.n_outlier + h.bits_main * (h.n - h.n_outlier))
