#!/usr/bin/env python
"""Host-side cost of one plugin call (the training hooks make hundreds per step): back-to-back calls on a
tensor small enough that the GPU never limits.  usage: python tools/call_overhead.py [numel]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "smart-quantization_b200")]
import torch
from bench import make_plugin
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
fp = make_plugin()
x = torch.randn(n, device="cuda")
for _ in range(50): fp(x, tag="t")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): fp(x, tag="t")
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"numel {n}: host {1e6*(t1-t0)/2000:.1f} us/call, with drain {1e6*(t2-t0)/2000:.1f} us/call")
