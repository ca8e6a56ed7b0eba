#!/usr/bin/env python
"""Device time of S2FP8's two passes (statistics, apply) on a B200, CUDA events, per input distribution.

    python tools/s2_bench.py [--log2n 30] [--reps 20]
    SMAQ_B200_LIB=.../libsmaq_s2exact.so python tools/s2_bench.py        # another build of the library (A/B)

Prints one JSON line per distribution: ms per launch, algorithmic GB/s (statistics 4 B/element, apply 8 B/element)
and the fraction of MEASURED_PEAKS.json's HBM copy peak.  The input is larger than L2 at the default size; smaller
sizes flush L2 between launches.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]
import torch  # noqa: E402

from bench import load_peaks, make_input, make_plugin  # noqa: E402
from smart_compress import _native as N  # noqa: E402
from smart_compress.util.pytorch.quantization import make_floatq_params  # noqa: E402


def distributions(n, dev):
    g = torch.Generator(device=dev).manual_seed(1)
    yield "bench_input", make_input(n, dev)
    yield "gradient_3e-6", torch.randn(n, generator=g, device=dev) * 3e-6
    x = torch.randn(n, generator=g, device=dev)
    yield "relu_half_zero", x.clamp_(min=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=30)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = N.load()
    n = 1 << a.log2n
    peak, peak_src = load_peaks()
    fp = make_plugin()
    p8 = make_floatq_params(5, 2, fp.hparams)
    y = torch.empty(n, dtype=torch.float32, device=dev)
    mm = torch.empty(2, dtype=torch.float32, device=dev)
    sws_b = lib.smaq_stats_workspace_bytes(n)
    sws = torch.empty(sws_b, dtype=torch.uint8, device=dev)
    st = N.stream_ptr(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if 8 * n < (512 << 20) else None
    for name, x in distributions(n, dev):
        def stats():
            N.check(lib.smaq_s2fp8_stats(x.data_ptr(), n, mm.data_ptr(), sws.data_ptr(), sws_b, st), "stats")

        def apply():
            N.check(lib.smaq_s2fp8_apply(x.data_ptr(), y.data_ptr(), n, mm.data_ptr(), None, C.byref(p8), st), "apply")

        out = {"input": name, "log2n": a.log2n, "lib": os.path.basename(os.environ.get("SMAQ_B200_LIB", "libsmaq_b200.so"))}
        def fp8():
            N.check(lib.smaq_float_quantize(x.data_ptr(), y.data_ptr(), n, None, C.byref(p8), st), "fp8")

        for label, fn, bytes_per in (("stats", stats, 4), ("apply", apply, 8), ("fp8", fp8, 8)):
            stats()
            for _ in range(3):
                fn()
            tot = 0.0
            for _ in range(a.reps):
                if flush is not None:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                tot += e0.elapsed_time(e1)
            ms = tot / a.reps
            gbs = bytes_per * n / ms / 1e6
            out[label] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
        both = out["stats"]["ms"] + out["apply"]["ms"]
        out["s2fp8"] = {"ms": round(both, 4), "GBps": round(12 * n / both / 1e6, 1), "frac": round(12 * n / both / 1e6 / peak, 4)}
        out["peak"] = {"GBps": peak, "source": peak_src}
        print(json.dumps(out), flush=True)
        del x


if __name__ == "__main__":
    main()
