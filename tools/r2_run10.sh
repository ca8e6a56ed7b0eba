#!/bin/bash
python -m pytest tests/test_gpu_pack.py tests/test_gpu_hooks.py tests/test_gpu_smaq.py tests/test_gpu_edge_cases.py -m gpu -q --timeout=900 2>&1 | tail -8
for L in "" deep8 deep64; do
  if [ -n "$L" ]; then export SMAQ_B200_LIB=$PWD/smart-quantization_b200/smart_compress/_lib/libsmaq_$L.so; fi
  python bench.py --steps 10 --warmup 3 --no-train --no-e2e --no-cpu > gpurun_out/ab_$L.json 2> /dev/null
  python - "$L" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1] or "default", {k: v["ms"] for k, v in d["kernels"].items()}, [(r["log2n"], r["encode"]["ms"]) for r in d["sweep"]["smaq"]])
PY
done
unset SMAQ_B200_LIB
python tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 20 --warmup 5 --packed-activations 2>&1 | tail -1 | cut -c1-330
