#!/bin/bash
# 2-GPU sanity: compressed all-reduce check, DDP with the compressed all-reduce, bench.py under torchrun
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $T --master-port 29540 tools/allreduce_check.py > gpurun_out/ar.log 2>&1; grep "^{" gpurun_out/ar.log | cut -c1-200; grep -A3 "Error" gpurun_out/ar.log | head -8
for AR in "" "--compress-allreduce p2p" "--compress-allreduce nccl"; do
  timeout 200 $T --master-port 29542 tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 30 --warmup 8 $AR > gpurun_out/ar_train.log 2>&1
  grep "^{" gpurun_out/ar_train.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step', 'loss', 'compress_allreduce', 'allreduce_stats')})"
  grep -B2 -A6 "Error" gpurun_out/ar_train.log | head -20
done
timeout 200 $T --master-port 29543 tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 10 --warmup 5 --profile 2>&1 | grep "profile\]" | head -20
timeout 600 $T --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
grep -A3 "Error" gpurun_out/r2_bench_n2.err | head
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["n_gpus"], d["e2e"]["value"]); print(json.dumps(d["train"])[:1500])
PY
