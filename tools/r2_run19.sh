#!/bin/bash
# encoder at 3 resident CTAs per SM (80 registers, -DSMAQ_ENC_CTAS=3) against 2 (125 registers): parity + timings
L=$PWD/smart-quantization_b200/smart_compress/_lib
{
for lib in libsmaq_b200.so libsmaq_enc3.so; do
  echo "== $lib"
  SMAQ_B200_LIB=$L/$lib timeout 600 python -m pytest tests/test_gpu_pack.py -m gpu -q --timeout=300 -x 2>&1 | tail -2
  for rep in 1 2; do
  SMAQ_B200_LIB=$L/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-train > gpurun_out/enc_ab_$lib.json 2> gpurun_out/enc_ab.err; tail -2 gpurun_out/enc_ab.err
  python - "$lib" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/enc_ab_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("step", d["value"], d["kernels"]["encode"])
for r in d["sweep"]["smaq"]: print(r["log2n"], "encode", r["encode"], "decode", r["decode"]["ms"])
print("pure normal", d["sweep"]["pure_normal_input"]["encode"])
PY
  done
done
} > gpurun_out/run19.log 2>&1
tail -60 gpurun_out/run19.log
