#!/usr/bin/env python
"""Per-launch device time of the statistics / round-trip kernels and the fused entry the plugin calls, at the
tensor sizes training produces (2^14 .. 2^26 elements): 50 launches queued between two CUDA events, so the
figure is what a stream of hook calls pays per call on the GPU (launch gaps included, host cost excluded).

    python tools/midsize_bench.py [--min 14] [--max 26]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]
import torch  # noqa: E402

from bench import make_input, make_plugin  # noqa: E402
from smart_compress import _native as N  # noqa: E402
from smart_compress.compress.packed import packed_layout  # noqa: E402
from smart_compress.util.pytorch.quantization import make_floatq_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--min", type=int, default=14)
ap.add_argument("--max", type=int, default=26)
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel CUPTI durations")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = N.load()
fp = make_plugin()
st = N.stream_ptr(dev)


def queued(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def kernel_us(fn, reps):
    """GPU-side duration per kernel (CUPTI via torch.profiler): {kernel name prefix: mean us}."""
    from collections import defaultdict

    from torch.profiler import ProfilerActivity, profile

    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    tot, cnt = defaultdict(float), defaultdict(int)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name.split("<")[0].replace("void ", "").replace("smaq::", "")
            tot[k] += e.device_time
            cnt[k] += 1
    return {k: tot[k] / cnt[k] for k in tot}


print(f"{'log2n':>5} {'MB':>8} {'stats us':>9} {'rt us':>9} {'fused us':>9} {'fused GB/s (12 B/el)':>20}   GPU-side kernel durations (us)")
for log2n in range(a.min, a.max + 1):
    n = 1 << log2n
    x = make_input(n, dev)
    y = torch.empty_like(x)
    ms = torch.empty(2, dtype=torch.float32, device=dev)
    params = fp._params(all_positive=False)
    sws_b = lib.smaq_stats_workspace_bytes(n)
    sws = torch.empty(sws_b, dtype=torch.uint8, device=dev)
    cws_b = lib.smaq_compress_workspace_bytes(n)
    cws = torch.empty(cws_b, dtype=torch.uint8, device=dev)
    N.check(lib.smaq_compress_workspace_init(cws.data_ptr(), cws_b, st), "init")
    t_s = queued(lambda: lib.smaq_stats_full(x.data_ptr(), n, 1, ms.data_ptr(), sws.data_ptr(), sws_b, st), a.reps)
    t_r = queued(lambda: lib.smaq_roundtrip(x.data_ptr(), y.data_ptr(), n, ms.data_ptr(), None, C.byref(params), st), a.reps)
    t_f = queued(lambda: lib.smaq_compress(x.data_ptr(), y.data_ptr(), n, None, C.byref(params), cws.data_ptr(), cws_b, st), a.reps)
    lay = packed_layout(n, 6, 8)
    packed = torch.empty(lay.total_capacity_bytes, dtype=torch.uint8, device=dev)
    pws = torch.zeros(lay.workspace_bytes, dtype=torch.uint8, device=dev)  # smaq_encode_workspace_init: zero once
    p8 = make_floatq_params(5, 2, fp.hparams)

    def everything():
        lib.smaq_compress(x.data_ptr(), y.data_ptr(), n, None, C.byref(params), cws.data_ptr(), cws_b, st)
        lib.smaq_encode(x.data_ptr(), n, ms.data_ptr(), None, C.byref(params), packed.data_ptr(), packed.numel(),
                        pws.data_ptr(), pws.numel(), st)
        lib.smaq_decode(packed.data_ptr(), packed.numel(), n, 6, 8, 0, y.data_ptr(), st)
        lib.smaq_float_quantize(x.data_ptr(), y.data_ptr(), n, None, C.byref(p8), st)

    ku = {} if a.no_kernels else kernel_us(everything, 20)
    print(f"{log2n:>5} {4 * n / 2**20:>8.2f} {t_s:>9.2f} {t_r:>9.2f} {t_f:>9.2f} {12.0 * n / t_f / 1e3:>20.0f}   " +
          "  ".join(f"{k} {v:.2f}" for k, v in sorted(ku.items())))
