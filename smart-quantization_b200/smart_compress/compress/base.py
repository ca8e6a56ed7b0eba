"""Plugin base type — the contract of reference smart_compress/compress/base.py:25-106.

Kept: static ``add_argparse_args(parent_parser)`` that chains parsers, ``Cls(hparams)``,
``update_hparams``, ``__call__(tensor, tag=None, **kwargs)``, the externally assigned ``log`` /
``log_custom`` callables, and the compression-ratio bookkeeping (``log_ratio`` / ``log_size``),
which is only evaluated under ``--measure_compression_ratio`` (base.py:79).
"""
from __future__ import annotations

from argparse import ArgumentParser, Namespace

import torch

_RATIO_KEYS = ("compression_ratio", "new_size", "orig_size")


@torch.no_grad()
def _sum_reduce(values):
    """Sizes are summed over a logging window, not averaged (reference base.py:8-18)."""
    if not isinstance(values, list):
        return torch.sum(values)
    if not values:
        return 0
    return torch.sum(torch.stack(values)) if torch.is_tensor(values[0]) else sum(values)


def chain_parser(parent: ArgumentParser) -> ArgumentParser:
    return ArgumentParser(parents=[parent], add_help=False)


class CompressionAlgorithmBase:
    log = None          # assigned by the training harness (reference util/train.py:209)
    log_custom = None   # ditto (:210); receives metrics for tags starting with "optimizer_"

    @staticmethod
    def add_argparse_args(parent_parser: ArgumentParser) -> ArgumentParser:
        parser = chain_parser(parent_parser)
        parser.add_argument("--measure_compression_ratio", action="store_true", dest="measure_compression_ratio")
        return parser

    def __init__(self, hparams: Namespace):
        super().__init__()
        self.hparams = hparams

    def update_hparams(self, hparams: Namespace):
        self.hparams = hparams

    # -- metrics ---------------------------------------------------------------------------
    def _emit(self, scalars: dict, custom: bool):
        if custom and self.log_custom is not None:
            self.log_custom(scalars)
            return
        for key, value in scalars.items():
            extra = dict(reduce_fx=_sum_reduce, tbptt_reduce_fx=_sum_reduce) if "size" in key else {}
            self.log(key, value, **extra)

    def log_ratio(self, tag, size, orig_bitcount, new_bitcount, overhead=0):
        return self.log_size(tag, size * orig_bitcount, size * new_bitcount, overhead=overhead)

    def log_size(self, tag, orig_size, new_size, overhead=0):
        if not getattr(self.hparams, "measure_compression_ratio", False):
            return
        orig = orig_size() if callable(orig_size) else orig_size
        new = (new_size() if callable(new_size) else new_size) + overhead
        assert hasattr(self, "log")
        values = dict(zip(_RATIO_KEYS, (orig / new, new, orig)))
        scalars = {}
        for key, value in values.items():
            scalars[key] = float(value)
            scalars[f"{key}_{tag}"] = float(value)
        self._emit(scalars, custom=str(tag).startswith("optimizer_"))

    def __call__(self, tensor: torch.Tensor, tag: str = None, **_):
        raise Exception("Not implemented")
