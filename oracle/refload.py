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


def _install_stubs(float_quantize_impl=None):
    import torch.nn as nn

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = type("LightningModule", (nn.Module,), {})
    prof = types.ModuleType("pytorch_lightning.profiler")
    base = types.ModuleType("pytorch_lightning.profiler.base")
    base.BaseProfiler = _NullProfiler
    au = types.ModuleType("argparse_utils")
    aum = types.ModuleType("argparse_utils.mapping")
    au.mapping_action = aum.mapping_action = lambda *a, **k: None
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
