"""The compression flag surface and its wiring — reference smart_compress/util/train.py:94-213 without Lightning.

The reference builds its command line in two phases (``init_model_from_args``): phase 1 picks the classes
(``--compress {bf16,fp8,fp16,fp32,s2fp8,smart}`` -> ``compression_cls``, default ``fp32``; ``--no_compress``;
``--compression_hook_fn {autograd,global_hook}``; the five ``--no_compress_*`` switches; ``--compress_loss``,
train.py:118-163) with ``parse_known_args``; phase 2 lets the chosen class add its own flags
(``args.compression_cls.add_argparse_args(parser)``, :180) and parses again.  Then it instantiates the codec
(:197-199), injects the loggers (:209-210) and installs the hooks (:212-213).

Model / dataset / Trainer flags (``--model``, ``--dataset``, Lightning's ``Trainer.add_argparse_args``) are the
control plane and out of scope; ``parse_compression_args`` tolerates them (unknown flags are ignored, the way the
reference's phase 1 does) and reads Lightning's ``--precision`` because the codecs do (smart.py:82-84).
"""
from __future__ import annotations

import argparse
import inspect
from argparse import ArgumentParser, Namespace
from typing import List, Optional, Sequence, Union

from .pytorch.autograd import register_autograd_module
from .pytorch.hooks import register_global_hooks, wrap_optimizer

DATA_STRUCTURES = ("forward", "backward", "weights", "gradients", "momentum_vectors")


def mapping_action(mapping: dict):
    """``argparse_utils.mapping_action``: the option takes one of ``mapping``'s keys and stores the mapped value
    (a string default is mapped too — the reference relies on it: train.py:58 asserts ``compression_cls`` is a class)."""

    class MappingAction(argparse.Action):
        def __init__(self, option_strings, dest, default=None, **kwargs):
            kwargs.pop("choices", None)
            if isinstance(default, str):
                default = mapping[default]
            super().__init__(option_strings, dest, default=default, choices=list(mapping), **kwargs)

        def __call__(self, parser, namespace, values, option_string=None):
            setattr(namespace, self.dest, mapping[values])

    return MappingAction


def compression_classes() -> dict:
    from ..compress import ALGORITHMS

    return dict(ALGORITHMS)  # bf16, fp8, fp16, fp32, s2fp8, smart — train.py:119-126


def add_compression_args(parser: ArgumentParser) -> ArgumentParser:
    """Phase-1 flags, names / dests / defaults verbatim from reference util/train.py:118-163."""
    parser.add_argument("--no_compress", action="store_false", dest="compress")
    parser.add_argument("--compress", action=mapping_action(compression_classes()), default="fp32",
                        dest="compression_cls")
    parser.add_argument("--compression_hook_fn",
                        action=mapping_action(dict(autograd=register_autograd_module, global_hook=register_global_hooks)),
                        default="autograd")
    for what in DATA_STRUCTURES:
        parser.add_argument(f"--no_compress_{what}", action="store_false", dest=f"compress_{what}")
    parser.add_argument("--compress_loss", action="store_true", dest="compress_loss")
    return parser


def _add_arg_names(args: Namespace) -> Namespace:
    """train.py:52-71: classes / functions are mirrored as strings so hparams serialise."""
    for name, value in dict(vars(args)).items():
        if value is None:
            continue
        if name.endswith("_cls"):
            assert inspect.isclass(value), f"{name} is not a class"
            setattr(args, f"{name}_name", f"{value.__module__}.{value.__name__}")
        elif name.endswith("_fn"):
            assert inspect.isfunction(value), f"{name} is not a function"
            setattr(args, f"{name}_name", value.__name__)
    return args


def parse_compression_args(argv: Union[None, str, Sequence[str]] = None, parser: Optional[ArgumentParser] = None,
                           strict: bool = False) -> Namespace:
    """The reference's two-phase parse restricted to the compression surface.  ``parser`` may already hold the
    caller's own flags (model, data, trainer).  Flags nobody declared are ignored, as in the reference's phase 1
    (the model / dataset / Trainer flags live outside this package); ``strict=True`` makes the second phase reject
    them as the reference's does (train.py:184) — e.g. ``--num_bits_main`` without ``--compress smart``."""
    if isinstance(argv, str):
        argv = argv.split(" ")  # train.py:91-92
    parser = parser if parser is not None else ArgumentParser()
    declared = {s for a in parser._actions for s in a.option_strings}
    if "--compress" not in declared:
        parser = add_compression_args(parser)
    if "--precision" not in declared:  # Lightning's Trainer flag; the codecs read hparams.precision
        parser.add_argument("--precision", type=int, default=32)
    args, _ = parser.parse_known_args(argv)                      # phase 1, train.py:171
    parser = args.compression_cls.add_argparse_args(parser)      # phase 2, train.py:180
    args = parser.parse_args(argv) if strict else parser.parse_known_args(argv)[0]   # train.py:184
    return _add_arg_names(args)


def build_compression(args: Namespace, model=None, optimizer=None, log=None, log_custom=None):
    """train.py:197-213 (+ models/base.py:152-157 for the optimizer): instantiate the codec unless
    ``--no_compress``, inject the loggers, install the feature-map / gradient-map hooks with the chosen hook
    function, wrap the optimizer.  Returns ``(compression, model, optimizer)``; ``Globals.compression`` is set
    (train.py:216)."""
    from .globals import Globals

    compression = args.compression_cls(args) if args.compress else None   # train.py:197-199
    if compression is not None:
        if log is not None:
            compression.log = log                                         # train.py:209
        if log_custom is not None:
            compression.log_custom = log_custom                           # train.py:210
    if compression is not None and model is not None:
        if args.compression_hook_fn is register_global_hooks:
            # the reference calls hook_fn(model, compression, hparams) (train.py:213) although
            # register_global_hooks's signature is (compress_fn, hparams) and it returns handles, not a model
            # (broken as wired); here the global hook is installed and the model kept
            register_global_hooks(compression, args)
        else:
            model = args.compression_hook_fn(model, compression, args)    # train.py:212-213
    if compression is not None and optimizer is not None and (
            args.compress_weights or args.compress_gradients or args.compress_momentum_vectors):
        optimizer = wrap_optimizer(optimizer, compression, args)          # models/base.py:152-157
    Globals.compression = compression
    return compression, model, optimizer


def compress_loss(loss, compression, args):
    """models/base.py:114-115: ``loss.data = compression(loss.data, tag="loss")`` under ``--compress_loss``."""
    if compression is not None and getattr(args, "compress_loss", False):
        loss.data = compression(loss.data, tag="loss")
    return loss


def compression_argv(compress: str = "smart", only: Optional[List[str]] = None, extra: Sequence[str] = ()) -> List[str]:
    """Command line for ``--compress NAME`` on the listed data structures (all five when ``only`` is None)."""
    argv = ["--compress", compress]
    if only is not None:
        argv += [f"--no_compress_{w}" for w in DATA_STRUCTURES if w not in only]
    return argv + list(extra)
