"""Gradient compression fused with the all-reduce — SURVEY.md §8 f-3.  NOT in the reference and OFF by default.

The reference compresses gradients AFTER DistributedDataParallel has all-reduced them in fp32
(smart_compress/util/pytorch/optimizer.py:135-141): the codec saves memory, not NVLink traffic.  This module moves
the codec in front of the wire as a DDP communication hook; it changes numerics (every rank's contribution is
quantised before the sum, and the mean is quantised once more on its way back), so it must be asked for::

    register_compressed_allreduce(ddp_model, codec)          # codec: a SmartFP instance

Per gradient bucket of B fp32 elements on N ranks (mean-reduction, like DDP's default hook):

  1. every rank ENCODES its bucket once into a packed SmaQ stream (6/8 bits per element) that lives in SYMMETRIC
     memory — mapped into every peer over NVLink / NVSwitch;                         [smaq_stats_full + smaq_encode]
  2. barrier; rank r owns the r-th shard of CTA tiles and runs ONE kernel that decodes that shard of ALL N streams
     straight out of the peers' memory and sums them: the kernel's loads are the reduce-scatter — no receive
     buffer, no NCCL call; the fixed-stride stream (SQB3) makes a tile range a contiguous slice;   [smaq_decode_sum]
  3. rank r re-encodes its reduced shard into a second symmetric buffer;           [smaq_stats_full + smaq_encode]
  4. barrier; every rank decodes the N reduced shards straight out of the peers' memory into its bucket: the
     decoder's loads are the all-gather.                                                          [smaq_decode x N]

Wire traffic per rank: 2 (N-1)/N x ~0.8 B per element instead of 2 (N-1)/N x 4 B for the fp32 ring all-reduce.
Every rank decodes the same reduced streams, so replicas stay bit-identical.  Buckets smaller than ``min_numel``
(and anything that is not CUDA fp32) take the ordinary fp32 all-reduce.

``transport="nccl"`` is the fallback without symmetric memory: the same four steps with an all-to-all and an
all-gather of the packed bytes in between (NCCL moves them, the kernels read local copies).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from ... import _native as N
from ...compress.packed import packed_layout

CTA_TILE = 8192


def shard_tiles(n_cta_tiles: int, world: int):
    """(first tile, tile count) of every rank's shard: contiguous, sizes differ by at most one tile."""
    base, extra = divmod(n_cta_tiles, world)
    out, first = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


class _Arena:
    """The wire buffers, allocated ONCE (on the caller's thread, at registration): region A holds this rank's packed
    bucket, region B its packed reduced shard.  With transport "p2p" they are symmetric memory — the rendezvous is a
    collective and must not happen inside DDP's hook, which runs on autograd's thread — and every bucket of a step
    reuses them (the barriers order the reuse)."""

    def __init__(self, max_numel: int, world: int, device, group, transport: str, bits):
        lay = packed_layout(max(max_numel, 1), *bits)
        self.cap_a = (int(lay.total_capacity_bytes) + 255) // 256 * 256
        self.cap_b = self.cap_a   # a shard is never larger than the bucket
        self.max_numel = max_numel
        if transport == "p2p":
            import torch.distributed._symmetric_memory as symm

            self.buf = symm.empty(self.cap_a + self.cap_b, dtype=torch.uint8, device=device)
            self.hdl = symm.rendezvous(self.buf, group)
            self.peer_a = [int(p) for p in self.hdl.buffer_ptrs]
            self.peer_b = [int(p) + self.cap_a for p in self.hdl.buffer_ptrs]
        else:
            self.buf = torch.empty(self.cap_a + self.cap_b, dtype=torch.uint8, device=device)
            self.hdl = None
            self.recv_a = torch.empty(world * self.cap_a, dtype=torch.uint8, device=device)   # every rank's stream
            self.recv_b = torch.empty(world * self.cap_b, dtype=torch.uint8, device=device)
            self.peer_a = [self.recv_a.data_ptr() + r * self.cap_a for r in range(world)]
            self.peer_b = [self.recv_b.data_ptr() + r * self.cap_b for r in range(world)]
        self.ptr_array = (C.c_void_p * world)(*self.peer_a)


class _BucketPlan:
    """Everything that depends only on a bucket's size: layouts and shard geometry (pure host arithmetic)."""

    def __init__(self, n: int, world: int, bits):
        self.n = n
        self.lay = packed_layout(n, *bits)
        self.bytes_a = int(self.lay.total_capacity_bytes)
        self.shards = shard_tiles(int(self.lay.n_cta_tiles), world)
        self.shard_elems = [max(0, min(n, (f + c) * CTA_TILE) - f * CTA_TILE) for f, c in self.shards]
        self.shard_bytes = [int(packed_layout(max(e, 1), *bits).total_capacity_bytes) for e in self.shard_elems]


class CompressedAllReduce:
    def __init__(self, codec, max_numel: int, group=None, transport: str = "p2p", min_numel: int = 1 << 16, device=None):
        if transport not in ("p2p", "nccl"):
            raise ValueError(transport)
        self.codec = codec
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise NotImplementedError("smaq_decode_sum takes at most 8 sources per call")
        self.transport = transport
        self.min_numel = min_numel
        self.plans = {}
        self.bits = (codec.hparams.num_bits_main, codec.hparams.num_bits_outlier)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.arena = _Arena(max_numel, self.world, device, self.group, transport, self.bits)
        self.stats = {"compressed_buckets": 0, "plain_buckets": 0, "wire_bytes": 0, "fp32_wire_bytes": 0}

    # ---------------------------------------------------------------------------------------------------------------
    def _encode_into(self, x: torch.Tensor, dst_ptr: int, dst_bytes: int):
        """statistics + packed encode of a flat fp32 tensor into raw device memory (a slice of the symmetric buffer)."""
        lib = N.load()
        codec = self.codec
        stream = N.stream_ptr(x.device)
        ms = codec.statistics(x)
        key = (x.device.index, stream)
        ws = codec._encode_ws.get(key)
        if ws is None:
            ws = codec._encode_ws[key] = torch.empty(64, dtype=torch.uint8, device=x.device)
            N.check(lib.smaq_encode_workspace_init(N.ptr(ws), ws.numel(), stream), "smaq_encode_workspace_init")
        params = codec._params(all_positive=False)
        N.check(lib.smaq_encode(N.ptr(x), x.numel(), N.ptr(ms), None, C.byref(params), dst_ptr, dst_bytes, N.ptr(ws),
                                ws.numel(), stream), "smaq_encode")

    def _barrier(self):
        if self.arena.hdl is not None:
            self.arena.hdl.barrier(channel=0)   # a device-side barrier over the signal pads, on the current stream

    def allreduce_mean_(self, flat: torch.Tensor) -> torch.Tensor:
        """In place: flat <- mean over ranks of flat, through the packed wire format."""
        n = flat.numel()
        ar = self.arena
        if (n < self.min_numel or n > ar.max_numel or not flat.is_cuda or flat.dtype != torch.float32
                or not flat.is_contiguous() or self.world == 1):
            self.stats["plain_buckets"] += 1
            dist.all_reduce(flat, group=self.group)
            return flat.div_(self.world)
        lib = N.load()
        plan = self.plans.get(n)
        if plan is None:
            plan = self.plans[n] = _BucketPlan(n, self.world, self.bits)
        stream = N.stream_ptr(flat.device)
        base = ar.buf.data_ptr()
        first, count = plan.shards[self.rank]
        # 1. my bucket as a packed stream where every peer can read it
        self._barrier()                         # nobody is still reading the buffers of the previous bucket
        self._encode_into(flat, base, ar.cap_a)
        if ar.hdl is not None:
            self._barrier()
        else:
            dist.all_gather_into_tensor(ar.recv_a, ar.buf[:ar.cap_a], group=self.group)
        # 2. reduce my shard straight out of the peers' streams
        if count > 0:
            N.check(lib.smaq_decode_sum(ar.ptr_array, self.world, ar.cap_a, n, self.bits[0], self.bits[1], first, count,
                                        1.0 / self.world, flat.data_ptr(), stream), "smaq_decode_sum")
            # 3. the reduced shard goes back on the wire packed
            mine = flat[first * CTA_TILE: first * CTA_TILE + plan.shard_elems[self.rank]]
            self._encode_into(mine, base + ar.cap_a, ar.cap_b)
        if ar.hdl is not None:
            self._barrier()
        else:
            dist.all_gather_into_tensor(ar.recv_b, ar.buf[ar.cap_a: ar.cap_a + ar.cap_b], group=self.group)
        # 4. everybody decodes everybody's reduced shard (my own included: replicas stay bit-identical)
        for r in range(self.world):
            f, c = plan.shards[r]
            e = plan.shard_elems[r]
            if c == 0 or e == 0:
                continue
            N.check(lib.smaq_decode(ar.peer_b[r], ar.cap_b, e, self.bits[0], self.bits[1], 0,
                                    flat.data_ptr() + 4 * f * CTA_TILE, stream), "smaq_decode")
        self.stats["compressed_buckets"] += 1
        payload = 0.8 * n   # ~6.3 bits per element on gradients; the exact figure is in the stream headers
        self.stats["wire_bytes"] += int(2 * (self.world - 1) / self.world * payload)
        self.stats["fp32_wire_bytes"] += int(2 * (self.world - 1) / self.world * 4 * n)
        return flat

    # DDP communication hook: hook(state, bucket) -> Future[Tensor]
    def hook(self, _state, bucket):  # (no annotations: DDP compares them with the classes themselves)
        buf = bucket.buffer()
        with N.on_device_of(buf):
            self.allreduce_mean_(buf)
        fut = torch.futures.Future()
        fut.set_result(buf)
        return fut


def register_compressed_allreduce(ddp_model, codec, transport: str = "p2p", min_numel: int = 1 << 16,
                                  group: Optional["dist.ProcessGroup"] = None) -> CompressedAllReduce:
    """Install the compressed all-reduce on a DistributedDataParallel model (opt-in; see the module docstring)."""
    params = [p for p in ddp_model.parameters() if p.requires_grad]
    max_numel = sum(p.numel() for p in params)   # no bucket is larger than all gradients together
    car = CompressedAllReduce(codec, max_numel, group=group, transport=transport, min_numel=min_numel,
                              device=params[0].device)
    ddp_model.register_comm_hook(None, car.hook)
    return car
