"""The C-ABI library builds, loads, and exports every symbol include/smaq_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "smaq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|size_t|int64_t|char\s*\*|const char\s*\*)\s*\*?\s*(smaq_\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("smaq_stats_full", "smaq_stats_sampled", "smaq_roundtrip", "smaq_encode", "smaq_decode",
                 "smaq_float_quantize", "smaq_s2fp8_stats", "smaq_s2fp8_apply", "smaq_roundtrip_multi"):
        assert must in names, must


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as entry

    entry.build()
    from smart_compress import _native

    lib = _native.load()
    raw = C.CDLL(_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} declared in smaq_b200.h but not exported"
        assert name in _native.EXPORTED_SYMBOLS, f"{name} has no ctypes signature"
    assert lib.smaq_b200_abi_version() == _native.ABI_VERSION == 4
    assert set(_native.EXPORTED_SYMBOLS) <= set(declared_functions())


def test_struct_layouts_match_the_header():
    """ctypes mirrors of the ABI structs must have the C sizes (checked against a compiled probe)."""
    import subprocess
    import tempfile

    from smart_compress import _native as N

    src = r'''
#include <stdio.h>
#include "smaq_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(smaq_codec_params), sizeof(smaq_floatq_params), sizeof(smaq_tensor_desc),
         sizeof(smaq_packed_layout), sizeof(smaq_packed_header));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    got = [C.sizeof(t) for t in (N.CodecParams, N.FloatqParams, N.TensorDesc, N.PackedLayout, N.PackedHeader)]
    assert got == sizes


def test_cpu_tensors_are_refused_not_emulated():
    import pytest
    import torch
    from argparse import ArgumentParser

    from smart_compress._native import NativeLibraryError
    from smart_compress.compress import FP8, S2FP8, SmartFP

    for cls in (SmartFP, FP8, S2FP8):
        args = cls.add_argparse_args(ArgumentParser()).parse_args([])
        args.precision = 32
        with pytest.raises(NativeLibraryError):
            cls(args)(torch.randn(64))


def test_batched_entry_points_refuse_cpu_tensors_too():
    """compress_many (what OptimLP calls per phase) must fail as loudly as the per-tensor call: no CPU emulation."""
    import pytest
    import torch
    from argparse import ArgumentParser

    from smart_compress._native import NativeLibraryError
    from smart_compress.compress import BF16, FP8, FP16, SmartFP

    for cls in (SmartFP, FP8, FP16, BF16):
        args = cls.add_argparse_args(ArgumentParser()).parse_args([])
        args.precision = 32
        with pytest.raises(NativeLibraryError):
            cls(args).compress_many([torch.randn(64), torch.randn(100)], None, tag="optimizer_grad")
