"""SmaQ plugin — the API of reference smart_compress/compress/smart.py:10-190 on sm_100a kernels.

One call = (at most) two kernel launches and no host synchronisation:
  1. statistics  -> ``smaq_stats_full`` / ``_sampled`` / ``_range``   (smart.py:130-134)
  2. round trip  -> ``smaq_roundtrip``                                 (smart.py:151-182)
Tensors of at most ``smaq_fused_small_max()`` elements with default statistics take one launch
(``smaq_roundtrip_small``).  The reference needs ~28 launches, a blocking ``std_dev == 0`` read
and two scalar uploads for the same call.

Flags, defaults, kwargs and return conventions are the reference's.  Extra keyword arguments
understood here (the reference swallows unknown kwargs, smart.py:117) exist for parity tests:
``_probs`` (explicit uniform numbers), ``_sample_idx`` (explicit sample indices), ``_out``.
"""
from __future__ import annotations

import ctypes as C
import itertools
import threading
from argparse import ArgumentParser, Namespace
from typing import Optional, Tuple, Union

import torch

from .. import _native as N
from ..util.globals import Globals
from .base import CompressionAlgorithmBase, chain_parser

_FLAGS = (
    # (flag, kwargs) in the reference's order (smart.py:17-69)
    ("--num_samples", dict(type=int, default=16, help="number of samples to use for mean/std_dev calculation")),
    ("--use_sample_stats", dict(action="store_true", help="use sample mean and std for smart compression")),
    ("--no_stochastic_rounding",
     dict(action="store_false", dest="stochastic_rounding", help="use stochastic rounding when quantizing")),
    ("--num_bits_main", dict(type=int, default=6, help="number of bits for main data (within 1 std dev)")),
    ("--num_bits_outlier", dict(type=int, default=8, help="number of bits for outlier data (more than 1 std dev)")),
    ("--main_std_dev_threshold", dict(type=float, default=1.0, help="std dev to consider something main")),
    ("--outlier_std_dev_threshold",
     dict(type=float, default=2.5, help="max std dev for outliers (everything else is clamped to this)")),
    ("--min_size", dict(type=int, default=8)),
    ("--use_range_std_dev", dict(action="store_true", help="use range std dev (from range batch norm paper)")),
    ("--use_batch_norm", dict(action="store_true", help="support BN acceleration")),
    ("--bn_scalar_params", dict(action="store_true", help="BN params should be scalar")),
)


class SmartFP(CompressionAlgorithmBase):
    @staticmethod
    def add_argparse_args(parent_parser: ArgumentParser):
        parser = chain_parser(CompressionAlgorithmBase.add_argparse_args(parent_parser))
        for flag, kw in _FLAGS:
            parser.add_argument(flag, **kw)
        return parser

    def __init__(self, hparams: Namespace):
        super().__init__(hparams)
        self._derive(self.hparams)
        self._calls = itertools.count()  # Philox stream offset: one stream per call
        self._tls = threading.local()
        self._small_max = int(N.load().smaq_fused_small_max())
        self._call_ws = {}               # (device, stream) -> scratch of the fused statistics + round-trip call
        self._lib = N.load()
        self._ws_bytes = None
        self._desc_cache = {}            # compress_many: device descriptor arrays by (pointers, sizes)
        self._multi_ws = {}
        self._encode_ws = {}             # (device, stream) -> the packed encoder's 64-byte scratch
        self._graph_keep = []            # buffers a captured CUDA graph reads at replay

    def _derive(self, hp):
        # smart.py:75-84, evaluated in Python floats exactly as there
        self.range_outlier = ((2 ** (hp.num_bits_outlier - 2)) - 1) / (
            hp.outlier_std_dev_threshold - hp.main_std_dev_threshold
        )
        self.range_normal = ((2 ** (hp.num_bits_main - 2)) - 1) / hp.main_std_dev_threshold
        self.clamped_range = (1e-4, 1e4) if getattr(hp, "precision", 32) == 16 else (1e-38, 1e38)

    def _lean_flags(self) -> bool:
        """Whether the flags allow the lean path of __call__ (read per call: tests and callers mutate hparams)."""
        hp = self.hparams
        try:
            return not (hp.use_sample_stats or hp.use_range_std_dev or hp.use_batch_norm
                        or hp.measure_compression_ratio) and hp.min_size <= self._small_max
        except AttributeError:  # a hand-made Namespace without every flag: the general path uses getattr defaults
            return False

    # ------------------------------------------------------------------------------------------
    def _params(self, all_positive: bool, saturate: bool = False, offset: Optional[int] = None) -> N.CodecParams:
        """The per-call constants (smart.py:72-84 + the call's kwargs).  The flag-derived part is filled once
        per thread (forward calls come from the main thread, backward calls from autograd's worker); a call
        only stamps its own kwargs, the seed and its Philox stream number."""
        hp = self.hparams
        p = getattr(self._tls, "params", None)
        if p is None:
            p = N.CodecParams()
            # like the reference: the ranges and the clamp are derived ONCE, in __init__ (smart.py:75-84) ...
            p.range_main = self.range_normal
            p.range_outlier = self.range_outlier
            p.clamp_lo, p.clamp_hi = self.clamped_range
            self._tls.params = p
        # ... while the threshold (smart.py:155-160) and the widths (smart.py:184-187) are read from hparams on
        # every call, so update_hparams() or an in-place edit acts exactly as it does there
        p.threshold = hp.main_std_dev_threshold
        p.bits_main = hp.num_bits_main
        p.bits_outlier = hp.num_bits_outlier
        p.stochastic = 1 if hp.stochastic_rounding else 0
        p.all_positive = 1 if all_positive else 0
        p.saturate = 1 if saturate else 0
        # the packed encoder counts what it clipped only when the size accounting is on (like the
        # reference, which only pays for its accounting under --measure_compression_ratio, base.py:79)
        p.count_saturated = 1 if getattr(hp, "measure_compression_ratio", False) else 0
        p.zero_on_grid = 0
        # torch.manual_seed() governs the stream, as it governs the reference's rand_like
        p.seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        sc = N.active_counter()   # inside N.counted_step(): numbered from the step's start + a device counter
        if sc is None:
            p.offset = next(self._calls) if offset is None else offset
            p.offset_base = None
        else:
            p.offset = sc.next() if offset is None else offset
            p.offset_base = sc.base_ptr
        return p

    def _next_stream(self) -> int:
        sc = N.active_counter()
        return next(self._calls) if sc is None else sc.next()

    def statistics(self, flat: torch.Tensor, sample_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device float[2] = (mean, std) per smart.py:130-134 — no host round trip."""
        lib = N.load()
        hp = self.hparams
        n = flat.numel()
        out = torch.empty(2, dtype=torch.float32, device=flat.device)
        stream = N.stream_ptr(flat.device)
        if hp.use_sample_stats:
            k = min(n, hp.num_samples)
            rng = 1 if hp.use_range_std_dev else 0  # _get_std(sample): the range estimate over the k samples
            if sample_idx is not None:
                idx = sample_idx.to(device=flat.device, dtype=torch.int64).contiguous()
                N.check(lib.smaq_stats_sampled(N.ptr(flat), n, N.ptr(idx), int(idx.numel()), rng, N.ptr(out), stream),
                        "smaq_stats_sampled")
            elif k <= 1024:
                N.check(lib.smaq_stats_sampled_draw(N.ptr(flat), n, k, rng, torch.initial_seed() & (2**64 - 1),
                                                    (1 << 62) + self._next_stream(), N.ptr(out), stream),
                        "smaq_stats_sampled_draw")
            else:
                idx = torch.randperm(n, device=flat.device)[:k]
                N.check(lib.smaq_stats_sampled(N.ptr(flat), n, N.ptr(idx), k, rng, N.ptr(out), stream),
                        "smaq_stats_sampled")
            return out
        ws_bytes = lib.smaq_stats_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=flat.device)
        if hp.use_range_std_dev:
            N.check(lib.smaq_stats_range(N.ptr(flat), n, N.ptr(out), N.ptr(ws), ws_bytes, stream), "smaq_stats_range")
        else:
            N.check(lib.smaq_stats_full(N.ptr(flat), n, 1, N.ptr(out), N.ptr(ws), ws_bytes, stream),
                    "smaq_stats_full")
        return out

    def __call__(
        self,
        data: torch.Tensor,
        tag: str = None,
        all_positive=False,
        batch_norm_stats: Union[Tuple[torch.Tensor, torch.Tensor], None] = None,
        **extra,
    ):
        # The call the training hooks make hundreds of times per step — default statistics, no size accounting,
        # no test kwargs, a contiguous fp32 CUDA tensor above the single-block size — goes straight to
        # smaq_compress: the small configs are bound by the host's cost per call, and the general path below
        # spends 10 us of Python around 7 us of launches.  Nothing here records autograd history (a fresh
        # output buffer and raw pointers), so it needs no no_grad scope.
        if not extra and batch_norm_stats is None and Globals.profiler is None and self._lean_flags():
            numel = data.numel()
            if (numel > self._small_max and data.dtype is torch.float32 and data.is_cuda and data.is_contiguous()
                    and not data.is_sparse and data.device.index == N._get_device()):
                device = data.device
                stream = N.stream_ptr(device)
                key = (device.index, stream)
                ws = self._call_ws.get(key)
                lib = self._lib
                if ws is None or ws.numel() < self._ws_need(numel):
                    ws = self._grow_ws(key, numel, device, stream)
                out = torch.empty_like(data)
                rc = lib.smaq_compress(data.data_ptr(), out.data_ptr(), numel, None,
                                       C.byref(self._params(bool(all_positive))), ws.data_ptr(), ws.numel(), stream)
                if rc:
                    N.check(rc, "smaq_compress")
                return out
        with torch.no_grad(), (N.on_device_of(data) if data.is_cuda else N._NO_SWITCH):
            profiler = Globals.profiler
            if profiler is not None:
                with profiler.profile("smaq"):
                    return self._call(data, tag, all_positive, batch_norm_stats, extra)
            return self._call(data, tag, all_positive, batch_norm_stats, extra)

    def _ws_need(self, numel: int) -> int:
        # smaq_compress_workspace_bytes does not depend on the element count (per-block partials of a capped
        # grid): ask once
        need = self._ws_bytes
        if need is None:
            need = self._ws_bytes = int(self._lib.smaq_compress_workspace_bytes(numel))
        return need

    def _grow_ws(self, key, numel, device, stream):
        need = int(self._lib.smaq_compress_workspace_bytes(numel))
        ws = self._call_ws[key] = torch.empty(need, dtype=torch.uint8, device=device)
        N.check(self._lib.smaq_compress_workspace_init(N.ptr(ws), ws.numel(), stream), "smaq_compress_workspace_init")
        return ws

    def _call(self, data, tag, all_positive, batch_norm_stats, extra):
        hp = self.hparams
        numel = data.numel()
        orig_size = numel * 32
        if numel < hp.min_size:  # smart.py:125-128: the SAME tensor object comes back
            self.log_ratio(tag, orig_size, 32, 32)
            return data

        N.require_cuda_f32(data, "SmartFP")
        lib = N.load()
        use_bn = bool(getattr(hp, "use_batch_norm", False)) and batch_norm_stats is not None
        probs = extra.get("_probs")
        sample_idx = extra.get("_sample_idx")

        # dense layouts (channels_last, permuted views) are processed in storage order and keep their strides; the
        # batch-norm mode indexes channels from the NCHW position and explicit sample indices are logical positions
        if data.is_contiguous() or (not use_bn and sample_idx is None and probs is None and N.is_dense(data)):
            src = data
        else:
            src = data.contiguous()
        flat = N.storage_order(src)
        out = torch.empty_like(src)
        stream = N.stream_ptr(data.device)
        params = self._params(all_positive, offset=extra.get("_offset"))
        probs_ptr = None
        if probs is not None:
            probs = probs.to(device=data.device, dtype=torch.float32).contiguous()
            assert probs.numel() == numel
            probs_ptr = N.ptr(probs)

        default_stats = not hp.use_sample_stats and not hp.use_range_std_dev
        mean_std = None
        if default_stats and not use_bn and numel > self._small_max and not hp.measure_compression_ratio:
            # the common call of the training hooks: statistics + round trip behind ONE entry point, on a
            # grow-only scratch buffer per (device, stream) — calls on one stream are ordered
            key = (data.device.index, stream)
            ws = self._call_ws.get(key)
            if ws is None or ws.numel() < lib.smaq_compress_workspace_bytes(numel):
                ws = self._grow_ws(key, numel, data.device, stream)
            N.check(lib.smaq_compress(N.ptr(flat), N.ptr(out), numel, probs_ptr, C.byref(params), N.ptr(ws), ws.numel(),
                                      stream), "smaq_compress")
            return out
        if default_stats and not use_bn and numel <= self._small_max:
            if hp.measure_compression_ratio:
                mean_std = torch.empty(2, dtype=torch.float32, device=data.device)
            N.check(
                lib.smaq_roundtrip_small(N.ptr(flat), N.ptr(out), numel, probs_ptr, C.byref(params),
                                         None if mean_std is None else N.ptr(mean_std), stream),
                "smaq_roundtrip_small",
            )
        else:
            mean_std = self.statistics(flat, sample_idx)  # statistics precede the BN un-affine (smart.py:130-149)
            if use_bn:
                gamma, beta, channels, inner = self._bn_affine_params(src, batch_norm_stats)
                N.check(
                    lib.smaq_roundtrip_bn(N.ptr(flat), N.ptr(out), numel, N.ptr(mean_std), probs_ptr, N.ptr(gamma),
                                          N.ptr(beta), channels, inner, C.byref(params), stream),
                    "smaq_roundtrip_bn",
                )
            else:
                N.check(
                    lib.smaq_roundtrip(N.ptr(flat), N.ptr(out), numel, N.ptr(mean_std), probs_ptr, C.byref(params),
                                       stream),
                    "smaq_roundtrip",
                )

        if hp.measure_compression_ratio:
            self.log_size(tag, orig_size, lambda: self._compressed_bits(flat, mean_std, params))
        return out

    # -- many tensors, one launch (the optimizer side) ------------------------------------------
    @torch.no_grad()
    def compress_many(self, tensors, kwargs_list=None, tag: str = None, stats_out: Optional[dict] = None):
        """``[self(t, tag=tag, **kw) for t, kw in zip(tensors, kwargs_list)]`` in at most four launches
        (``smaq_roundtrip_multi``: per-tensor full statistics + round trip, one Philox stream per tensor;
        tensors up to ``smaq_fused_small_max()`` elements take one block each, larger ones are cut into
        work items).  OptimLP loops the codec over every
        parameter, gradient and state tensor (reference optimizer.py:69-127); most of them are tiny
        (median 512 elements for ResNet-18), so per-tensor launches are pure latency.  The batched
        tensors are updated IN PLACE and returned as the same objects (the reference re-binds ``.data``
        to a fresh tensor; nothing else aliases optimizer tensors, so the effect is the same).
        ``stats_out`` (a dict, parity tests): receives ``{index: device float[2]}`` with the (mean, std)
        each batched tensor was quantised with."""
        hp = self.hparams
        lib = N.load()
        small_max = lib.smaq_fused_small_max()
        per_tensor = (hp.use_sample_stats or hp.use_range_std_dev or hp.measure_compression_ratio
                      or Globals.profiler is not None)
        results = list(tensors)
        batch = []
        # Philox streams are numbered like the per-tensor loop numbers them: one per tensor that is
        # actually quantised (numel >= min_size), in order
        first = None
        numbered = 0
        for i, t in enumerate(tensors):
            kw = kwargs_list[i] if kwargs_list is not None else {}
            n = t.numel()
            if n < hp.min_size:
                results[i] = self(t, tag=tag, **kw)  # comes back untouched (smart.py:125-128)
                continue
            if first is None:
                first = self._next_stream()
            else:
                self._next_stream()
            stream_no = numbered
            numbered += 1
            ok = (not per_tensor and t.is_cuda and t.dtype == torch.float32
                  and N.is_dense(t) and "batch_norm_stats" not in kw and "_probs" not in kw)
            if ok:
                batch.append((i, t, bool(kw.get("all_positive", False)), stream_no))
            else:
                results[i] = self(t, tag=tag, _offset=first + stream_no, **kw)
        if not batch:
            return results
        by_device = {}
        for item in batch:  # an optimizer whose parameters span GPUs: one launch set per device
            by_device.setdefault(item[1].device, []).append(item)
        for device, items in by_device.items():
            with N.on_device_of(items[0][1]):
                self._launch_many(device, items, first, stats_out)
        return results

    def _launch_many(self, device, batch, first, stats_out=None):
        hp = self.hparams
        lib = self._lib
        key = (device, tuple((t.data_ptr(), t.numel(), ap, sn) for _, t, ap, sn in batch))
        descs = self._desc_cache.get(key)
        if descs is None:
            host = (N.TensorDesc * len(batch))()
            for j, (_, t, ap, sn) in enumerate(batch):
                host[j].x = host[j].y = t.data_ptr()
                host[j].n = t.numel()
                host[j].all_positive = int(ap)
                host[j].stream = sn
            N.ensure_pinned_arena()
            capturing = torch.cuda.is_current_stream_capturing()
            raw = (N.pinned_arena_take(bytes(host)) if capturing
                   else torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).pin_memory())
            descs = raw.to(device, non_blocking=True)
            if capturing:
                # the captured copy reads `raw` at every replay: both live as long as the codec
                self._graph_keep.append((raw, descs))
            else:
                if len(self._desc_cache) > 64:
                    self._desc_cache.clear()
                self._desc_cache[key] = (descs, raw)
        else:
            descs = descs[0]
        params = self._params(all_positive=False, offset=first)
        total = sum(t.numel() for _, t, _, _ in batch)
        need = lib.smaq_multi_workspace_bytes(len(batch), total)
        stream = N.stream_ptr(device)
        ws_key = (device, stream)
        ws = self._multi_ws.get(ws_key)
        if ws is None or ws.numel() < need:  # grow-only scratch; calls on one stream are ordered
            ws = self._multi_ws[ws_key] = torch.empty(need, dtype=torch.uint8, device=device)
        ms = None
        if stats_out is not None:  # parity tests: the (mean, std) each tensor was quantised with
            ms = torch.zeros(len(batch), 2, dtype=torch.float32, device=device)
            for j, (i, _, _, _) in enumerate(batch):
                stats_out[i] = ms[j]
        N.check(
            lib.smaq_roundtrip_multi(N.ptr(descs), len(batch), max(t.numel() for _, t, _, _ in batch), total,
                                     C.byref(params), int(hp.min_size), N.ptr(ws), ws.numel(),
                                     None if ms is None else N.ptr(ms), stream),
            "smaq_roundtrip_multi",
        )

    # -- materialised stream: encode / decode ---------------------------------------------------
    @torch.no_grad()
    def encode(self, data: torch.Tensor, mean_std: Optional[torch.Tensor] = None, **extra):
        """Quantise and PACK ``data`` (6-bit main / 8-bit outlier codes with the default flags) into a
        device buffer; statistics as in ``__call__`` unless ``mean_std`` (device float[2]) is given.
        ``zero_on_grid=True`` (not in the reference): the mean is moved by at most half a step so that exact zeros
        decode to exact zeros — for saved ReLU outputs (``packed_saved_tensors``)."""
        from .packed import PackedSmaq, packed_layout

        N.require_cuda_f32(data, "SmartFP.encode")
        if N.wrong_device(data):
            with N.on_device_of(data):
                return self.encode(data, mean_std, **extra)
        lib = N.load()
        hp = self.hparams
        src = data if data.is_contiguous() else data.contiguous()
        flat = src.view(-1)
        n = flat.numel()
        lay = packed_layout(n, hp.num_bits_main, hp.num_bits_outlier)
        if mean_std is None:
            mean_std = self.statistics(flat, extra.get("_sample_idx"))
        split = bool(extra.get("split"))
        if split:  # header + planes (exact) and the extras (capacity) as two allocations: the second can be compacted
            buf = torch.empty(lay.extras_off, dtype=torch.uint8, device=data.device)
            ext = torch.empty(lay.extras_capacity_bytes, dtype=torch.uint8, device=data.device)
        else:
            buf = torch.empty(lay.total_capacity_bytes, dtype=torch.uint8, device=data.device)
            ext = None
        stream = N.stream_ptr(data.device)
        key = (data.device.index, stream)
        ws = self._encode_ws.get(key)
        if ws is None:  # 64 bytes, zeroed once: every encode leaves it zero (calls on one stream are ordered)
            ws = self._encode_ws[key] = torch.empty(max(int(lay.workspace_bytes), 64), dtype=torch.uint8, device=data.device)
            N.check(lib.smaq_encode_workspace_init(N.ptr(ws), ws.numel(), stream), "smaq_encode_workspace_init")
        params = self._params(all_positive=False)
        params.zero_on_grid = 1 if extra.get("zero_on_grid") else 0
        probs = extra.get("_probs")
        probs_ptr = None
        if probs is not None:
            probs = probs.to(device=data.device, dtype=torch.float32).contiguous()
            probs_ptr = N.ptr(probs)
        if split:
            N.check(
                lib.smaq_encode_split(N.ptr(flat), n, N.ptr(mean_std), probs_ptr, C.byref(params), N.ptr(buf), buf.numel(),
                                      N.ptr(ext), ext.numel(), N.ptr(ws), ws.numel(), stream),
                "smaq_encode_split",
            )
        else:
            N.check(
                lib.smaq_encode(N.ptr(flat), n, N.ptr(mean_std), probs_ptr, C.byref(params), N.ptr(buf), buf.numel(),
                                N.ptr(ws), ws.numel(), stream),
                "smaq_encode",
            )
        return PackedSmaq(buffer=buf, layout=lay, shape=data.shape, extras=ext)

    @torch.no_grad()
    def compact(self, packed, extras_words: int):
        """Replace a split stream's capacity-sized extras by the used words only (``extras_words`` = the header's
        count, which the caller has read — ``packed.header()`` or an asynchronous copy): afterwards the stream occupies
        bits_main * n_main + bits_outlier * n_outlier bits (+ table and padding), the reference's own accounting
        (smart.py:184-187), in memory.  Two small kernels on the current stream; the old buffer is released."""
        if packed.extras is None or packed.table is not None:
            return packed
        lay = packed.layout
        if lay.extras_stride_bytes == 0:
            return packed
        if N.wrong_device(packed.buffer):
            with N.on_device_of(packed.buffer):
                return self.compact(packed, extras_words)
        lib = N.load()
        dev = packed.buffer.device
        stream = N.stream_ptr(dev)
        table = torch.empty(int(lib.smaq_extras_table_entries(lay.n)), dtype=torch.int32, device=dev)
        dense = torch.empty((int(extras_words) + 16) * 4, dtype=torch.uint8, device=dev)  # + the decoder's look-ahead
        N.check(lib.smaq_extras_compact(N.ptr(packed.buffer), packed.buffer.numel(), N.ptr(packed.extras), lay.n,
                                        lay.bits_main, lay.bits_outlier, N.ptr(table), N.ptr(dense), int(extras_words) * 4,
                                        None, stream), "smaq_extras_compact")
        packed.extras, packed.table = dense, table
        return packed

    @torch.no_grad()
    def decode(self, packed, all_positive: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if N.wrong_device(packed.buffer):
            with N.on_device_of(packed.buffer):
                return self.decode(packed, all_positive, out)
        lib = N.load()
        lay = packed.layout
        y = out if out is not None else torch.empty(packed.shape, dtype=torch.float32, device=packed.buffer.device)
        ext, table = packed.extras, packed.table   # read once: a compaction may swap them
        if ext is not None:
            N.check(
                lib.smaq_decode_split(N.ptr(packed.buffer), packed.buffer.numel(), N.ptr(ext), ext.numel(),
                                      None if table is None else N.ptr(table), lay.n, lay.bits_main, lay.bits_outlier,
                                      int(bool(all_positive)), N.ptr(y), N.stream_ptr(y.device)),
                "smaq_decode_split",
            )
            return y
        N.check(
            lib.smaq_decode(N.ptr(packed.buffer), packed.buffer.numel(), lay.n, lay.bits_main, lay.bits_outlier,
                            int(bool(all_positive)), N.ptr(y), N.stream_ptr(y.device)),
            "smaq_decode",
        )
        return y

    # -- --use_batch_norm (smart.py:136-149,174-179): off by default -----------------------------
    def _bn_affine_params(self, x, stats):
        """(gamma, beta, channels, inner) for ``smaq_roundtrip_bn``.  The reference broadcasts gamma / beta over the
        LAST axis of ``x.permute(0, 3, 2, 1)``, i.e. over axis 1 of a 4-D map: per channel of an NCHW tensor;
        with --bn_scalar_params both are replaced by their means (smart.py:140-142)."""
        gamma, beta = stats
        if x.dim() != 4:
            raise RuntimeError(f"--use_batch_norm expects a 4-D (N, C, H, W) feature map, got {tuple(x.shape)} "
                               "(the reference's permute(0, 3, 2, 1) raises on anything else)")
        gamma = gamma.detach().to(device=x.device, dtype=torch.float32)
        beta = beta.detach().to(device=x.device, dtype=torch.float32)
        if self.hparams.bn_scalar_params:
            return gamma.mean().reshape(1), beta.mean().reshape(1), 1, x.numel()
        channels = x.shape[1]
        if gamma.numel() != channels or beta.numel() != channels:
            raise RuntimeError(f"batch_norm_stats of {gamma.numel()} / {beta.numel()} entries for {channels} channels")
        return gamma.contiguous().view(-1), beta.contiguous().view(-1), channels, x.shape[2] * x.shape[3]

    # -- size accounting (smart.py:184-187) ------------------------------------------------------
    def _compressed_bits(self, flat, mean_std, params) -> float:
        lib = N.load()
        counter = torch.zeros(1, dtype=torch.int64, device=flat.device)
        N.check(
            lib.smaq_count_outliers(N.ptr(flat), flat.numel(), N.ptr(mean_std), C.byref(params), N.ptr(counter),
                                    N.stream_ptr(flat.device)),
            "smaq_count_outliers",
        )
        n_out = int(counter.item())
        hp = self.hparams
        return n_out * hp.num_bits_outlier + (flat.numel() - n_out) * hp.num_bits_main
