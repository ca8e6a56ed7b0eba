#!/bin/bash
# FP8 emulation kernel: CTAs per resident slot 4 (default) / 2 / 1 at mid sizes
L=$PWD/smart-quantization_b200/smart_compress/_lib
for n in 20 22 24 26 30; do
  for lib in libsmaq_b200.so libsmaq_fqw2.so libsmaq_fqw1.so; do
    SMAQ_B200_LIB=$L/$lib timeout 300 python tools/s2_bench.py --log2n $n --reps 30 2>/dev/null | grep "^{" | head -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['lib'], d['log2n'], 'fp8', d['fp8']['ms'], d['fp8']['frac'], 'apply', d['apply']['ms'])"
  done
done
