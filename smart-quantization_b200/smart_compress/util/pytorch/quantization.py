"""Layer predicate and float_quantize wrapper — reference
smart_compress/util/pytorch/quantization.py:8-204.

``is_valid_layer_type`` decides which module outputs reach the codec: conv / linear / pool /
normalisation layers by default, plus every ``torch.nn`` container and activation module and
every module defined under ``smart_compress.models.pytorch`` (quantization.py:163-184).

``float_quantize`` is the reference's wrapper (quantization.py:187-204) around qtorch 0.2.0's
``float_quantize(x, exp, man, rounding="stochastic")``; here both the rounding and the
"+max becomes +inf" fix-up run inside one sm_100a kernel (``smaq_float_quantize``).
"""
from __future__ import annotations

import ctypes as C
import itertools
from argparse import ArgumentParser

import torch
from torch import nn

from ... import _native as N

CONV_LAYERS = [nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d,
               nn.Unfold, nn.Fold]
POOL_LAYERS = [nn.MaxPool1d, nn.MaxPool2d, nn.MaxPool3d, nn.MaxUnpool1d, nn.MaxUnpool2d, nn.MaxUnpool3d,
               nn.AvgPool1d, nn.AvgPool2d, nn.AvgPool3d, nn.FractionalMaxPool2d, nn.LPPool1d, nn.LPPool2d,
               nn.AdaptiveMaxPool1d, nn.AdaptiveMaxPool2d, nn.AdaptiveMaxPool3d,
               nn.AdaptiveAvgPool1d, nn.AdaptiveAvgPool2d, nn.AdaptiveAvgPool3d]
PAD_LAYERS = [nn.ReflectionPad1d, nn.ReflectionPad2d, nn.ReplicationPad1d, nn.ReplicationPad2d, nn.ZeroPad2d,
              nn.ConstantPad1d, nn.ConstantPad2d, nn.ConstantPad3d]
ACTIVATION_LAYERS = [nn.ELU, nn.Hardshrink, nn.Hardtanh, nn.LeakyReLU, nn.LogSigmoid, nn.PReLU, nn.ReLU, nn.ReLU6,
                     nn.RReLU, nn.SELU, nn.Sigmoid, nn.Softplus, nn.Softshrink, nn.Softsign, nn.Tanh,
                     nn.Tanhshrink, nn.Threshold, nn.Softmin, nn.Softmax, nn.Softmax2d, nn.LogSoftmax]
NORM_LAYERS = [nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d, nn.GroupNorm, nn.InstanceNorm1d, nn.InstanceNorm2d,
               nn.InstanceNorm3d, nn.LayerNorm, nn.LocalResponseNorm]
LINEAR_LAYERS = [nn.Linear, nn.Bilinear]
DROPOUT_LAYERS = [nn.Dropout, nn.Dropout2d, nn.Dropout3d, nn.AlphaDropout]
LOSS_LAYERS = [nn.L1Loss, nn.MSELoss, nn.CrossEntropyLoss, nn.NLLLoss, nn.PoissonNLLLoss, nn.KLDivLoss, nn.BCELoss,
               nn.BCEWithLogitsLoss, nn.MarginRankingLoss, nn.HingeEmbeddingLoss, nn.MultiLabelMarginLoss,
               nn.SmoothL1Loss, nn.SoftMarginLoss, nn.MultiLabelSoftMarginLoss, nn.MultiMarginLoss,
               nn.TripletMarginLoss]

LAYERS_TYPES = {
    "conv": CONV_LAYERS,
    "linear": LINEAR_LAYERS,
    "pool": POOL_LAYERS,
    "pad": PAD_LAYERS,
    "activation": ACTIVATION_LAYERS,
    "normalization": NORM_LAYERS,
    "dropout": DROPOUT_LAYERS,
    "loss": LOSS_LAYERS,
}
DEFAULT_LAYER_TYPES = ["conv", "linear", "pool", "normalization"]

# modules whose *qualified type name* contains one of these are always wrapped
_ALWAYS_WRAPPED = ("smart_compress.models.pytorch.", "torch.nn.modules.container.", "torch.nn.modules.activation.")


def is_valid_layer_type(module, layer_types=DEFAULT_LAYER_TYPES):
    accepted = []
    for layer_type in layer_types:
        assert layer_type in LAYERS_TYPES
        accepted += LAYERS_TYPES[layer_type]
    kind = type(module)
    if kind in accepted:
        return True
    name = str(kind)
    return any(marker in name for marker in _ALWAYS_WRAPPED)


def add_float_quantize_args(parent_parser: ArgumentParser):
    parser = ArgumentParser(parents=[parent_parser], add_help=False)
    parser.add_argument("--no_float_quantize_check_inf", action="store_false", dest="float_quantize_check_inf")
    return parser


_calls = itertools.count()


def make_floatq_params(exp: int, man: int, hparams, rounding: str = "stochastic") -> N.FloatqParams:
    p = N.FloatqParams()
    p.exp_bits, p.man_bits = exp, man
    p.rounding = 1 if rounding == "stochastic" else 0
    p.check_inf = int(bool(getattr(hparams, "float_quantize_check_inf", True)))
    p.max_exp_bias = 0  # qtorch 0.2.0 rule (oracle/floatq.py)
    p.seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    p.offset, p.offset_base = _next_stream()
    return p


def _next_stream():
    """(stream number, device counter or None): see _native.counted_step."""
    sc = N.active_counter()
    if sc is None:
        return next(_calls), None
    return sc.next(), sc.base_ptr


def float_quantize(x: torch.Tensor, exp: int, man: int, hparams, rand_bits: torch.Tensor = None):
    """quantization.py:187-204.  Returns a new tensor; fp16 tensors round-trip through fp32 when
    ``hparams.precision == 16`` exactly as there."""
    if x.is_cuda and N.wrong_device(x):  # launch on the tensor's own GPU (the reference's eager ops do)
        with N.on_device_of(x):
            return float_quantize(x, exp, man, hparams, rand_bits)
    lib = N.load()
    is_16_bit = getattr(hparams, "precision", 32) == 16
    src = x.float() if is_16_bit else x
    N.require_cuda_f32(src, "float_quantize")
    if not (rand_bits is None and N.is_dense(src)):   # dense layouts keep their strides (elementwise, order-free)
        src = src.contiguous()
    out = torch.empty_like(src)
    params = make_floatq_params(exp, man, hparams)
    rb = None
    if rand_bits is not None:
        rand_bits = rand_bits.to(device=src.device, dtype=torch.int32).contiguous()
        rb = N.ptr(rand_bits)
    N.check(lib.smaq_float_quantize(N.ptr(src), N.ptr(out), src.numel(), rb, C.byref(params),
                                    N.stream_ptr(src.device)), "smaq_float_quantize")
    return out.half() if is_16_bit else out


_multi_cache = {}   # (device, tensors' pointers and sizes) -> device array of smaq_tensor_desc
_multi_ws = {}      # device -> grow-only scratch
_graph_keep = []    # buffers a captured CUDA graph reads at replay


def _descriptors(batch, device):
    """Device array of smaq_tensor_desc for [(index, tensor, stream number)], cached by pointers and sizes."""
    key = (device, tuple((t.data_ptr(), t.numel(), sn) for _, t, sn in batch))
    descs = _multi_cache.get(key)
    if descs is None:
        host = (N.TensorDesc * len(batch))()
        for j, (_, t, sn) in enumerate(batch):
            host[j].x = host[j].y = t.data_ptr()
            host[j].n = t.numel()
            host[j].all_positive = 0
            host[j].stream = sn
        N.ensure_pinned_arena()
        capturing = torch.cuda.is_current_stream_capturing()
        raw = (N.pinned_arena_take(bytes(host)) if capturing
               else torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).pin_memory())
        descs = raw.to(device, non_blocking=True)
        if capturing:
            _graph_keep.append((raw, descs))   # the captured copy reads `raw` at every replay
        else:
            if len(_multi_cache) > 64:
                _multi_cache.clear()
            _multi_cache[key] = (descs, raw)
    else:
        descs = descs[0]
    return descs


def s2fp8_many(tensors, hparams, stats_out=None):
    """``[S2FP8(hparams)(t) for t in tensors]`` in three launches (``smaq_s2fp8_multi``) — the loops OptimLP runs over
    every parameter, gradient and state tensor with --compress s2fp8 (reference optimizer.py:69-127, s2fp8.py:31-48).
    Contiguous fp32 CUDA tensors are updated IN PLACE and returned as the same objects; ``stats_out`` (a dict,
    parity tests) receives {index: device float[2]} with each tensor's (mu, m).  Returns None for tensors it did not
    take (the caller falls back to the per-tensor call)."""
    lead = next((t for t in tensors if t.is_cuda), None)
    if lead is not None and N.wrong_device(lead):
        with N.on_device_of(lead):
            return s2fp8_many(tensors, hparams, stats_out)
    lib = N.load()
    results = [None] * len(tensors)
    if getattr(hparams, "precision", 32) == 16:
        return results
    batch = []
    first = None
    for i, t in enumerate(tensors):
        if not (t.is_cuda and t.dtype == torch.float32 and N.is_dense(t)) or t.numel() == 0:
            continue
        if batch and t.device != batch[0][1].device:
            continue
        no = _next_stream()[0]
        if first is None:
            first = no
        batch.append((i, t, no - first))
    if not batch:
        return results
    device = batch[0][1].device
    descs = _descriptors(batch, device)
    params = make_floatq_params(5, 2, hparams)
    params.offset = first
    total = sum(t.numel() for _, t, _ in batch)
    need = lib.smaq_s2fp8_multi_workspace_bytes(len(batch), total)
    key = (device, "s2")
    ws = _multi_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _multi_ws[key] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=device)
    mm = None
    if stats_out is not None:
        mm = torch.zeros(len(batch), 2, dtype=torch.float32, device=device)
        for j, (i, _, _) in enumerate(batch):
            stats_out[i] = mm[j]
    N.check(lib.smaq_s2fp8_multi(N.ptr(descs), len(batch), total, C.byref(params), N.ptr(ws), ws.numel(),
                                 None if mm is None else N.ptr(mm), N.stream_ptr(device)), "smaq_s2fp8_multi")
    for i, t, _ in batch:
        results[i] = t
    return results


def float_quantize_many(tensors, exp: int, man: int, hparams):
    """``[float_quantize(t, exp, man, hparams) for t in tensors]`` in two launches (``smaq_float_quantize_multi``):
    what OptimLP's loops over every parameter, gradient and state tensor (reference optimizer.py:69-127) cost with
    --compress fp8 | fp16 | bf16 was one launch per tensor.  Philox streams are numbered as the per-tensor loop
    numbers its calls, so every tensor gets the bits that loop would give it.  Contiguous fp32 CUDA tensors are
    updated IN PLACE and returned as the same objects (the reference re-binds ``.data`` to a fresh tensor; nothing
    else aliases optimizer tensors); anything else goes through ``float_quantize``."""
    lead = next((t for t in tensors if t.is_cuda), None)
    if lead is not None and N.wrong_device(lead):  # launch on the tensors' own GPU (before any stream number is drawn)
        with N.on_device_of(lead):
            return float_quantize_many(tensors, exp, man, hparams)
    lib = N.load()
    results = list(tensors)
    is_16_bit = getattr(hparams, "precision", 32) == 16
    batch = []
    first = None
    for i, t in enumerate(tensors):
        if is_16_bit or not (t.is_cuda and t.dtype == torch.float32 and N.is_dense(t)) or t.numel() == 0:
            results[i] = float_quantize(t, exp, man, hparams)   # draws its own stream number
            continue
        no = _next_stream()[0]
        if first is None:
            first = no
        batch.append((i, t, no - first))
    if not batch:
        return results
    device = batch[0][1].device
    if any(t.device != device for _, t, _ in batch):
        for i, t, _ in batch:
            results[i] = float_quantize(t, exp, man, hparams)
        return results
    descs = _descriptors(batch, device)
    params = make_floatq_params(exp, man, hparams)
    params.offset = first      # make_floatq_params drew one more number: harmless, the streams stay distinct
    need = lib.smaq_floatq_multi_workspace_bytes(len(batch))
    ws = _multi_ws.get(device)
    if ws is None or ws.numel() < need:
        ws = _multi_ws[device] = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=device)
    total = sum(t.numel() for _, t, _ in batch)
    N.check(lib.smaq_float_quantize_multi(N.ptr(descs), len(batch), total, C.byref(params), N.ptr(ws), ws.numel(),
                                          N.stream_ptr(device)), "smaq_float_quantize_multi")
    return results
