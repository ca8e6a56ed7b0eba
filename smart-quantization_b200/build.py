"""Build libsmaq_b200.so (sm_100a) in-tree with nvcc.  No torch headers: the library is a plain C ABI.

    python smart-quantization_b200/build.py [--force] [--verbose]

The .so lands next to the Python host package (smart_compress/_lib/) so it travels with the repo
snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "smart_compress", "_lib")
# SMAQ_LIB_NAME: development builds next to the product library (A/B experiments; see SMAQ_B200_LIB in _native.py)
LIB = os.path.join(OUT_DIR, os.environ.get("SMAQ_LIB_NAME", "libsmaq_b200.so"))
STAMP = LIB + ".stamp"

SOURCES = ["capi.cu", "smaq_stats.cu", "smaq_roundtrip.cu", "smaq_pack.cu", "float_quantize.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # parity: IEEE division and square root, denormals kept
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC",
]


# Per-file FMA contraction.  The SmaQ element arithmetic must round after every operation, and
# ptxas was observed to fuse mul.rn.f32x2 + add.rn.f32x2 into FFMA2 under the default -fmad=true,
# so those translation units are built with -fmad=false (explicit fma.rn is unaffected).  The
# float-emulation / statistics units keep the default: libdevice's powf / log2f must be compiled
# the way torch compiles them for S2FP8 to agree bit for bit with torch's CUDA operators.
NO_FMAD = {"smaq_roundtrip.cu", "smaq_pack.cu"}


def extra_flags(source: str):
    flags = []
    if os.path.basename(source) in NO_FMAD:
        flags.append("-fmad=false")
    if os.environ.get("SMAQ_DEV") == "1":  # development builds: only the default 6/8-bit packed kernels
        flags.append("-DSMAQ_PACK_MINIMAL")
    flags += os.environ.get("SMAQ_NVCC_EXTRA", "").split()  # experiments, e.g. -DSMAQ_PHILOX_ROUNDS=7
    return flags


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("SMAQ_DEV", "0").encode())
    h.update(os.environ.get("SMAQ_NVCC_EXTRA", "").encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "smaq_b200.h"))
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == digest:
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUT_DIR, os.path.basename(LIB) + "." + os.path.basename(s)[:-3] + ".o")
        cmd = [nvcc_path(), *NVCC_FLAGS, *extra_flags(s), "-c", s, "-o", o]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    link = [nvcc_path(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]  # cudart is linked statically (nvcc default)
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    for o in objs:
        os.remove(o)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
