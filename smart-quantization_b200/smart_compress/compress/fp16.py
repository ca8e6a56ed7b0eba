"""FP16 emulation plugin (reference smart_compress/compress/fp16.py): every value is rounded to
e5m10 with stochastic rounding by ``float_quantize`` — one sm_100a kernel here."""
from argparse import ArgumentParser, Namespace

import torch

from ..util.pytorch.quantization import add_float_quantize_args, float_quantize, float_quantize_many
from .base import CompressionAlgorithmBase, chain_parser


class FP16(CompressionAlgorithmBase):
    EXP_BITS, MAN_BITS, STORED_BITS = 5, 10, 16

    @staticmethod
    def add_argparse_args(parent_parser: ArgumentParser):
        return chain_parser(add_float_quantize_args(CompressionAlgorithmBase.add_argparse_args(parent_parser)))

    def __init__(self, hparams: Namespace):
        super().__init__(hparams)

    @torch.no_grad()
    def __call__(self, tensor: torch.Tensor, tag: str = None, **extra):
        self.log_ratio(tag, tensor.numel(), 32, self.STORED_BITS)
        return float_quantize(tensor, exp=self.EXP_BITS, man=self.MAN_BITS, hparams=self.hparams,
                              rand_bits=extra.get("_rand_bits"))

    @torch.no_grad()
    def compress_many(self, tensors, kwargs_list=None, tag: str = None):
        """The optimizer-side loop (OptimLP) in two launches; kwargs (``all_positive``) do not matter to this codec."""
        if getattr(self.hparams, "measure_compression_ratio", False):
            for t in tensors:
                self.log_ratio(tag, t.numel(), 32, self.STORED_BITS)
        return float_quantize_many(tensors, self.EXP_BITS, self.MAN_BITS, self.hparams)
