"""Feature-map / gradient-map hooks — reference smart_compress/util/pytorch/autograd.py:12-77.

``Compressor`` wraps an ``autograd.Function`` whose forward compresses what a layer produced
(tag ``forward_autograd``) and whose backward compresses the gradient flowing back into that
layer (tag ``backward_autograd``).  ``register_autograd_module`` re-binds ``forward`` on every
module the layer predicate accepts, sharing one ``Compressor``.  Backward calls arrive on
autograd's device worker thread; the codec picks up that thread's current stream itself.
"""
from argparse import Namespace

import torch
import torch.nn as nn
from torch.autograd import Function

from .quantization import is_valid_layer_type


def process_input(args):
    """Split a trailing ``{"batch_norm_stats": ...}`` dict off the positional arguments."""
    if len(args) >= 1 and type(args[-1]) == dict and "batch_norm_stats" in args[-1]:
        return args[:-1], args[-1]
    return args, {}


class Compressor(nn.Module):
    def __init__(self, compress_fn, forward=True, backward=True):
        super().__init__()
        do_forward, do_backward = forward, backward

        class CompressorAutoGradFn(Function):
            @staticmethod
            def forward(ctx, x: torch.Tensor, *args, **kwargs):
                if not do_forward:
                    return x
                positional, extra = process_input(args)
                return compress_fn(x, *positional, **kwargs, **extra, tag="forward_autograd")

            @staticmethod
            def backward(ctx, grad_output):
                if not do_backward:
                    return grad_output, None
                if not ctx.needs_input_grad[0]:
                    return None, None
                return compress_fn(grad_output, tag="backward_autograd"), None

        self.compress_fn = CompressorAutoGradFn.apply

    def forward(self, *args, **kwargs):
        return self.compress_fn(*args, **kwargs)


def register_autograd_module(model: nn.Module, compress_fn, hparams: Namespace):
    compressor = Compressor(compress_fn, forward=hparams.compress_forward, backward=hparams.compress_backward)
    pass_bn_stats = bool(getattr(hparams, "use_batch_norm", False))

    def patch(module: nn.Module):
        if not is_valid_layer_type(module):
            return
        inner = module.forward
        with_stats = pass_bn_stats and type(module) == nn.BatchNorm2d

        apply = compressor.compress_fn  # == compressor(...) minus nn.Module.__call__'s hook dispatch (2 us per layer)

        def new_forward(*args, **kwargs):
            result = inner(*args, **kwargs)
            if with_stats:
                return apply(result, dict(batch_norm_stats=(module.weight.detach(), module.bias.detach())))
            return apply(result)

        module.forward = new_forward

    return model.apply(patch)


class packed_saved_tensors(torch.autograd.graph.saved_tensors_hooks):
    """Keep what autograd saves for backward as PACKED SmaQ streams (6/8 bits per element with the default
    flags, one byte of capacity) instead of fp32 — the memory reduction the reference's README claims
    (README.md:25) but its fake quantisation never delivers, since it stores fp32 (smart.py:154-172).

    Not in the reference and it changes numerics (a saved activation is quantised once more, with its own
    statistics), so it is opt-in::

        with packed_saved_tensors(codec):
            loss = model(x).sum()
        loss.backward()

    Parameters, small tensors (< ``min_numel``) and anything that is not CUDA fp32 are saved as they are.
    Exact zeros survive: SmaQ's grid has no point at 0.0 in general, and a saved ReLU *output* whose zeros came back as
    small values of either sign would give the backward mask ``output > 0`` wrong for half of the dead units (round 1
    did).  The encoder is therefore asked to put zero ON the grid (``zero_on_grid``: the mean the codes are relative
    to moves by at most half a quantisation step, on the device, no synchronisation), so 0.0 decodes to exactly 0.0.
    ``keep=`` still leaves chosen tensors alone.
    Exact-size storage without a synchronisation (``exact_size=True``, the default): the stream is encoded into two
    allocations — header + planes (6 bits per element, exact) and the extras at capacity (2 bits per element) — and
    its header is copied to pinned host memory asynchronously; at a LATER pack / unpack call, once that copy has
    landed (``event.query()``), the extras are compacted to the used words (two small kernels) and the capacity
    buffer is released.  All but the last few saved tensors therefore occupy what the reference only accounts for
    (smart.py:184-187): ~6.4 bits per element instead of 32, 5.0x instead of round 1's capacity-sized 4.0x.
    Encoding and decoding run on the calling thread's current stream (the backward pass decodes on
    autograd's worker thread)."""

    def __init__(self, codec, min_numel: int = 1 << 16, keep=None, zero_on_grid: bool = True, exact_size: bool = True):
        import ctypes as C
        import threading
        from collections import deque

        from ... import _native as N

        lock = threading.Lock()
        pending = deque()          # (packed, pinned header slot, event), oldest first
        slots = torch.empty(1024, 128, dtype=torch.uint8).pin_memory() if exact_size else None
        # smaq_packed_header.extras_words is the uint64 at byte 72 of the slot
        words_view = slots.numpy().view("<u8")[:, 9] if exact_size else None
        counter = [0]
        self.compacted = 0

        def drain():
            # compact every stream whose header has arrived (oldest first; stop at the first that has not)
            while pending and pending[0][2].query():
                packed, index, _ = pending.popleft()
                codec.compact(packed, int(words_view[index]))
                self.compacted += 1

        def pack(t: torch.Tensor):
            if (isinstance(t, nn.Parameter) or not t.is_cuda or t.dtype != torch.float32 or t.numel() < min_numel
                    or (keep is not None and keep(t))):
                return t
            exact = exact_size and not torch.cuda.is_current_stream_capturing()  # (a graph cannot poll an event)
            packed = codec.encode(t.detach(), zero_on_grid=zero_on_grid, split=exact)
            if exact:
                with lock:
                    if len(pending) >= slots.shape[0]:   # every slot in flight: wait for the oldest (never in practice)
                        pending[0][2].synchronize()
                    drain()
                    index = counter[0] % slots.shape[0]
                    counter[0] += 1
                    slots[index].copy_(packed.buffer[:128], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                    pending.append((packed, index, ev))
            return packed

        def unpack(obj):
            if isinstance(obj, torch.Tensor):
                return obj
            if exact_size:
                with lock:
                    drain()
                    return codec.decode(obj)
            return codec.decode(obj)

        super().__init__(pack, unpack)
