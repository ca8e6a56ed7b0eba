#!/usr/bin/env python
"""What the round trip finds in L2 when it runs right behind the statistics kernel on the same tensor (a hook call).

    ncu --cache-control none --clock-control none --profile-from-start off \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum \
        python tools/l2_evidence.py --log2n 22

`--cache-control none`: ncu must not flush the caches between the two kernels, or the question is void.  The tensor
is written by a torch kernel first (as a layer's output would be), then smaq_compress (statistics + round trip as a
dependent launch) runs `reps` times; the last repetition is profiled."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]
import torch  # noqa: E402

from bench import make_input, make_plugin  # noqa: E402
from smart_compress import _native as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=22)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--flush", action="store_true", help="write a 512 MB buffer between the producer and the hook call (cold L2)")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = N.load()
n = 1 << a.log2n
src = make_input(n, dev)
x = torch.empty_like(src)
y = torch.empty_like(src)
fp = make_plugin()
params = fp._params(all_positive=False)
st = N.stream_ptr(dev)
need = lib.smaq_compress_workspace_bytes(n)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
N.check(lib.smaq_compress_workspace_init(ws.data_ptr(), ws.numel(), st), "init")
junk = torch.empty(128 << 20, dtype=torch.float32, device=dev)
for rep in range(a.reps):
    x.copy_(src).mul_(1.0)                      # the producing layer: x was just written
    if a.flush:
        junk.fill_(float(rep))
    if rep == a.reps - 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    N.check(lib.smaq_compress(x.data_ptr(), y.data_ptr(), n, None, C.byref(params), ws.data_ptr(), ws.numel(), st), "compress")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(y[:8].sum()))
