#!/bin/bash
# round-2 GPU session 3: graph tests, config-3 / BERT eager vs graph, kernel A/B of the offset-base change
python -m pytest tests/test_gpu_hooks.py -m gpu -q --timeout=600 2>&1 | tail -15
for m in "resnet18 --batch 256 --image 32" "bert-base --batch 32 --seq 128"; do
  for g in "" "--cuda-graph"; do
    timeout 600 python tools/train_bench.py --model $m --compress smart --steps 50 --warmup 10 $g 2>&1 | tail -1 | cut -c1-330
  done
  timeout 600 python tools/train_bench.py --model $m --compress fp32 --steps 50 --warmup 10 2>&1 | tail -1 | cut -c1-200
  timeout 600 python tools/train_bench.py --model $m --compress fp32 --steps 50 --warmup 10 --cuda-graph 2>&1 | tail -1 | cut -c1-200
done
timeout 600 python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress fp8 --steps 50 --warmup 10 --cuda-graph 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 5 --warmup 3 --no-train --no-e2e --no-cpu > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench3.json").read().strip().splitlines()[-1])
for r in d["sweep"]["smaq"]:
    print("  ", r["log2n"], {k: (v["ms"], v["frac"]) for k, v in r.items() if k != "log2n"})
PY
