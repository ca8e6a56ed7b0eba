"""FP32 "compression": the identity plugin (reference smart_compress/compress/fp32.py:10-23)."""
from argparse import ArgumentParser, Namespace

import torch

from .base import CompressionAlgorithmBase, chain_parser


class FP32(CompressionAlgorithmBase):
    @staticmethod
    def add_argparse_args(parent_parser: ArgumentParser):
        return chain_parser(CompressionAlgorithmBase.add_argparse_args(parent_parser))

    def __init__(self, hparams: Namespace):
        super().__init__(hparams)

    @torch.no_grad()
    def __call__(self, tensor: torch.Tensor, tag: str = None, **_):
        self.log_ratio(tag, tensor.numel(), 32, 32)
        return tensor
