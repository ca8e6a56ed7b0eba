"""Optimizer wrapping and the alternative global forward hook —
reference smart_compress/util/pytorch/hooks.py:15-53."""
from argparse import Namespace
from functools import partial

import torch
import torch.nn as nn
from torch.nn.modules.module import register_module_forward_hook

from .optimizer import OptimLP
from .quantization import DEFAULT_LAYER_TYPES, is_valid_layer_type


def _wrap_fn(fn, **bound):
    def wrapped(*args, **kwargs):
        return fn(*args, **bound, **kwargs)

    many = getattr(fn, "compress_many", None)
    if many is not None:  # a codec that batches small tensors (SmartFP): OptimLP uses it per phase
        wrapped.compress_many = lambda tensors, kwargs_list=None: many(tensors, kwargs_list, **bound)
    return wrapped


def wrap_optimizer(optimizer, compress_fn, hparams: Namespace):
    """tags: optimizer_weight / optimizer_grad / optimizer_momentum (hooks.py:25-29)."""
    quantizers = {}
    if hparams.compress_weights:
        quantizers["weight_quant"] = _wrap_fn(compress_fn, tag="optimizer_weight")
    if hparams.compress_gradients:
        quantizers["grad_quant"] = _wrap_fn(compress_fn, tag="optimizer_grad")
    if hparams.compress_momentum_vectors:
        quantizers["momentum_quant"] = _wrap_fn(compress_fn, tag="optimizer_momentum")
    if not quantizers:
        return optimizer
    return OptimLP(optimizer, **quantizers)


def _register_forward_hook(compress_fn, layer_types=DEFAULT_LAYER_TYPES):
    def forward_hook(module: nn.Module, _inputs, output):
        if type(output) != torch.Tensor or not is_valid_layer_type(module, layer_types=layer_types):
            return None
        return compress_fn(output, tag="forward_hook")

    return register_module_forward_hook(forward_hook)


def register_global_hooks(compress_fn, hparams, layer_types=DEFAULT_LAYER_TYPES):
    handles = []
    if hparams.compress_forward:
        handles.append(_register_forward_hook(compress_fn, layer_types=layer_types))
    return handles
