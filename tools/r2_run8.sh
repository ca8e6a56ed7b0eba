#!/bin/bash
# 8-GPU session: compressed all-reduce check + DDP training with and without it, then bench.py as the driver runs it
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29540 tools/allreduce_check.py > gpurun_out/ar8.log 2>&1; grep "^{" gpurun_out/ar8.log | cut -c1-3000; grep -A3 "Error" gpurun_out/ar8.log | head -8
for AR in "" "--compress-allreduce p2p"; do
  timeout 300 $T --master-port 29542 tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 50 --warmup 10 $AR > gpurun_out/ar_train8.log 2>&1
  grep "^{" gpurun_out/ar_train8.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step', 'loss', 'compress_allreduce', 'allreduce_stats')})"
  grep -B2 -A6 "Error" gpurun_out/ar_train8.log | head -20
done
timeout 900 $T --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
grep -A3 "Error" gpurun_out/r2_bench_n8.err | head
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["n_gpus"], d["e2e"]); print(json.dumps(d["train"])[:1700])
PY
