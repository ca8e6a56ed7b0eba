"""Load the UNMODIFIED reference (`/root/reference`) in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it.  This module in particular only works where ``/root/reference``
exists (the build container) and is used by ``oracle/gen_golden.py`` to produce
the fixtures in ``tests/golden/`` and by the container-only pinning tests.  The
GPU box has no ``/root/reference``; nothing that runs there imports this file.

The reference imports three packages that are not installed here
(``pytorch_lightning``, ``argparse_utils``, ``qtorch``).  We register inert
stand-ins in ``sys.modules`` so that the reference's own files import
unchanged.  ``qtorch`` is *not* re-implemented
here: its stand-in raises, so FP8/S2FP8 cannot be pinned against the reference
(see oracle/floatq.py header: "parity unpinned").
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from argparse import ArgumentParser

REFERENCE_ROOT = os.environ.get("SMAQ_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "smart_compress", "compress", "smart.py"))


class _NullProfiler:
    """Stands in for pytorch_lightning's BaseProfiler (smart.py:119 calls .profile)."""

    @contextlib.contextmanager
    def profile(self, name):
        yield


class ParsedArgs(Exception):
    """Raised by the stand-in Trainer to hand the reference's fully parsed Namespace back to the caller."""

    def __init__(self, args):
        super().__init__("parsed")
        self.parsed = args


class _CapturingTrainer:
    """Stands in for pytorch_lightning.Trainer in util/train.py: contributes the one Trainer flag the codecs read
    (--precision) and stops ``init_model_from_args`` right after its two-phase parse (train.py:191)."""

    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--precision", type=int, default=32)
        parser.add_argument("--terminate_on_nan", action="store_true")
        return parser

    @staticmethod
    def from_argparse_args(args, **kwargs):
        raise ParsedArgs(args)


def _mapping_action(mapping):
    """argparse_utils.mapping_action: store mapping[value]; a string default is mapped as well (the reference
    asserts that ``compression_cls`` is a class even when --compress is absent, util/train.py:58)."""
    import argparse

    class MappingAction(argparse.Action):
        def __init__(self, option_strings, dest, default=None, **kwargs):
            kwargs.pop("choices", None)
            if isinstance(default, str):
                default = mapping[default]
            super().__init__(option_strings, dest, default=default, choices=list(mapping), **kwargs)

        def __call__(self, parser, namespace, values, option_string=None):
            setattr(namespace, self.dest, mapping[values])

    return MappingAction


def _install_stubs(float_quantize_impl=None):
    import torch.nn as nn

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = type("LightningModule", (nn.Module,), {})
    prof = types.ModuleType("pytorch_lightning.profiler")
    base = types.ModuleType("pytorch_lightning.profiler.base")
    base.BaseProfiler = _NullProfiler
    pl.LightningDataModule = type("LightningDataModule", (), {"__init__": lambda self, *a, **k: None})
    pl.Trainer = _CapturingTrainer
    loggers = types.ModuleType("pytorch_lightning.loggers")
    tt = types.ModuleType("pytorch_lightning.loggers.test_tube")
    tt.TestTubeLogger = lambda *a, **k: None
    plugins = types.ModuleType("pytorch_lightning.plugins")
    ttype = types.ModuleType("pytorch_lightning.plugins.training_type")
    ttype.DDPPlugin = type("DDPPlugin", (), {})
    au = types.ModuleType("argparse_utils")
    aum = types.ModuleType("argparse_utils.mapping")
    au.mapping_action = aum.mapping_action = _mapping_action
    qt = types.ModuleType("qtorch")
    qtq = types.ModuleType("qtorch.quant")
    qtf = types.ModuleType("qtorch.quant.quant_function")

    def _no_qtorch(*a, **k):
        raise RuntimeError("qtorch 0.2.0 is not available in this image (parity unpinned)")

    qtf.float_quantize = float_quantize_impl or _no_qtorch
    for name, mod in {
        "pytorch_lightning": pl,
        "pytorch_lightning.profiler": prof,
        "pytorch_lightning.profiler.base": base,
        "pytorch_lightning.loggers": loggers,
        "pytorch_lightning.loggers.test_tube": tt,
        "pytorch_lightning.plugins": plugins,
        "pytorch_lightning.plugins.training_type": ttype,
        "argparse_utils": au,
        "argparse_utils.mapping": aum,
        "qtorch": qt,
        "qtorch.quant": qtq,
        "qtorch.quant.quant_function": qtf,
    }.items():
        sys.modules.setdefault(name, mod)


@contextlib.contextmanager
def reference_modules(float_quantize_impl=None):
    """Context in which ``import smart_compress...`` resolves to the reference.

    On exit every ``smart_compress*`` module is evicted from ``sys.modules`` so
    the repo's own host-side mirror (same module paths, by design) can be
    imported afterwards in the same process.
    """
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    saved = {k: v for k, v in sys.modules.items() if k == "smart_compress" or k.startswith("smart_compress.")}
    for k in saved:
        del sys.modules[k]
    _install_stubs(float_quantize_impl)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "smart_compress" or k.startswith("smart_compress.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def load_reference_smartfp(argv=(), precision=32):
    """Return an instance of the reference's own SmartFP built from its own flags."""
    with reference_modules():
        from smart_compress.compress.smart import SmartFP
        from smart_compress.util.globals import Globals

        Globals.profiler = _NullProfiler()
        args = SmartFP.add_argparse_args(ArgumentParser()).parse_args(list(argv))
        args.precision = precision
        return SmartFP(args)


def reference_parse_args(argv):
    """Run the reference's own ``init_model_from_args`` (smart_compress/util/train.py:74-184) up to the end of
    its two-phase parse and return (Namespace, {name: value}) of what it parsed.  Classes and functions are
    reported by name (the reference's objects do not outlive the import context)."""
    import importlib

    with reference_modules():
        for optional in ("torchmetrics", "torchmetrics.functional", "datasets"):
            try:
                importlib.import_module(optional)
            except Exception:
                stub = types.ModuleType(optional)
                stub.load_dataset = stub.load_metric = lambda *a, **k: None
                sys.modules[optional] = stub
        from smart_compress.util.train import init_model_from_args

        try:
            init_model_from_args(list(argv))
        except ParsedArgs as e:
            args = e.parsed
        else:  # pragma: no cover
            raise RuntimeError("the stand-in Trainer did not stop init_model_from_args")
        flat = {}
        for k, v in vars(args).items():
            flat[k] = getattr(v, "__name__", v) if (isinstance(v, type) or callable(v)) else v
        return args, flat
