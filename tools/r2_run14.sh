#!/bin/bash
# S2FP8 apply at the sizes the hooks produce: table prologue cut to the distinct entries; CTAs per resident slot 4 / 2 / 1
L=$PWD/smart-quantization_b200/smart_compress/_lib
timeout 600 python -m pytest tests/test_gpu_floatq.py -m gpu -q --timeout=600 -x 2>&1 | tail -2
for n in 20 22 24 26 30; do
  for lib in libsmaq_b200.so libsmaq_s2w2.so libsmaq_s2w1.so; do
    SMAQ_B200_LIB=$L/$lib timeout 300 python tools/s2_bench.py --log2n $n --reps 30 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['lib'], d['log2n'], d['input'][:8], 'stats', d['stats']['ms'], 'apply', d['apply']['ms'], d['apply']['frac'])" | head -2
  done
done
