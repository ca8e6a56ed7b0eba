#!/bin/bash
# round trip: CTAs per SM of the multi-wave grid (SMAQ_RT_WAVES, default 8 = 2.67 waves of 3 resident CTAs) 6 / 9 / 12
L=$PWD/smart-quantization_b200/smart_compress/_lib
{
for lib in libsmaq_b200.so libsmaq_rtw6.so libsmaq_rtw9.so libsmaq_rtw12.so libsmaq_b200.so; do
  echo "== $lib"
  SMAQ_B200_LIB=$L/$lib timeout 300 python tools/midsize_bench.py --min 24 --max 29 --no-kernels 2>&1 | cut -c1-75 | grep -v -i warn
  SMAQ_B200_LIB=$L/$lib timeout 300 python bench.py --no-train --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for r in d['sweep']['smaq'][3:]: print(r['log2n'], 'rt', r['roundtrip']['ms'], r['roundtrip']['frac'], 'compress', r['compress']['ms'], r['compress']['frac'])
print('pure normal rt', d['sweep']['pure_normal_input']['roundtrip'])"
done
} > gpurun_out/run23.log 2>&1
cat gpurun_out/run23.log
