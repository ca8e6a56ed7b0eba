#!/bin/bash
# S2FP8 screened apply path: parity tests, then A/B timing against the exact-only build and a 3-CTA/SM build
L=smart-quantization_b200/smart_compress/_lib
timeout 900 python -m pytest tests/test_gpu_floatq.py -m gpu -x -q --timeout=600 -s 2>&1 | grep -v "^$" | tail -12
for lib in libsmaq_b200.so libsmaq_s2exact.so libsmaq_fq3.so libsmaq_b200.so; do
  SMAQ_B200_LIB=$PWD/$L/$lib timeout 300 python tools/s2_bench.py --log2n 30 2>&1 | grep "^{" | cut -c1-400
done
SMAQ_B200_LIB=$PWD/$L/libsmaq_b200.so timeout 300 python tools/s2_bench.py --log2n 24 2>&1 | grep "^{" | cut -c1-400
SMAQ_B200_LIB=$PWD/$L/libsmaq_s2exact.so timeout 300 python tools/s2_bench.py --log2n 24 2>&1 | grep "^{" | cut -c1-400
