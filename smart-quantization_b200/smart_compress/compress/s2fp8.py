"""S2FP8 plugin — reference smart_compress/compress/s2fp8.py:11-48 in two sm_100a kernels.

Pass 1 (``smaq_s2fp8_stats``): mu = mean(L), m = max(L) with L = log2|x| and L = 0 where x == 0
(the reference's ``torch.where(X_abs == 0, X_abs, log2(X_abs))``).  Pass 2 (``smaq_s2fp8_apply``):
alpha = 15/(m - mu), beta = -alpha*mu, y = sign(x) * (Q_e5m2(|x|^alpha * 2^beta) * 2^-beta)^(1/alpha).
The reference spends ~12 full-tensor passes on the same thing.  No host synchronisation.
"""
import ctypes as C
from argparse import ArgumentParser, Namespace

import torch

from .. import _native as N
from ..util.pytorch.quantization import add_float_quantize_args, make_floatq_params, s2fp8_many
from .base import CompressionAlgorithmBase, chain_parser


class S2FP8(CompressionAlgorithmBase):
    @staticmethod
    def add_argparse_args(parent_parser: ArgumentParser):
        return chain_parser(add_float_quantize_args(CompressionAlgorithmBase.add_argparse_args(parent_parser)))

    def __init__(self, hparams: Namespace):
        super().__init__(hparams)

    def statistics(self, flat: torch.Tensor) -> torch.Tensor:
        lib = N.load()
        out = torch.empty(2, dtype=torch.float32, device=flat.device)
        ws_bytes = lib.smaq_stats_workspace_bytes(flat.numel())
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=flat.device)
        N.check(lib.smaq_s2fp8_stats(N.ptr(flat), flat.numel(), N.ptr(out), N.ptr(ws), ws_bytes,
                                     N.stream_ptr(flat.device)), "smaq_s2fp8_stats")
        return out

    @torch.no_grad()
    def __call__(self, tensor: torch.Tensor, tag: str = None, **extra):
        if tensor.is_cuda and N.wrong_device(tensor):  # launch on the tensor's own GPU (the reference's eager ops do)
            with N.on_device_of(tensor):
                return self.__call__(tensor, tag, **extra)
        self.log_ratio(tag, tensor.numel(), 32, 8, overhead=64)
        is_16_bit = getattr(self.hparams, "precision", 32) == 16
        src = tensor.float() if is_16_bit else tensor
        N.require_cuda_f32(src, "S2FP8")
        lib = N.load()
        if not (extra.get("_rand_bits") is None and N.is_dense(src)):   # dense layouts keep their strides
            src = src.contiguous()
        flat = N.storage_order(src)
        out = torch.empty_like(src)
        if flat.numel() == 0:
            return out
        mu_max = extra.get("_mu_max")
        if mu_max is None:
            mu_max = self.statistics(flat)
        params = make_floatq_params(5, 2, self.hparams)
        rand_bits = extra.get("_rand_bits")
        rb = None
        if rand_bits is not None:
            rand_bits = rand_bits.to(device=src.device, dtype=torch.int32).contiguous()
            rb = N.ptr(rand_bits)
        N.check(lib.smaq_s2fp8_apply(N.ptr(flat), N.ptr(out), flat.numel(), N.ptr(mu_max), rb, C.byref(params),
                                     N.stream_ptr(src.device)), "smaq_s2fp8_apply")
        return out.half() if is_16_bit else out

    @torch.no_grad()
    def compress_many(self, tensors, kwargs_list=None, tag: str = None, stats_out=None):
        """The loops OptimLP runs over every parameter, gradient and state tensor (reference optimizer.py:69-127) in
        three launches (``smaq_s2fp8_multi``) instead of two per tensor; in place, same objects back."""
        tensors = list(tensors)
        for t in tensors:
            self.log_ratio(tag, t.numel(), 32, 8, overhead=64)
        done = s2fp8_many(tensors, self.hparams, stats_out)
        return [d if d is not None else self(t, tag=tag) for d, t in zip(done, tensors)]
