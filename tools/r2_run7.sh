#!/bin/bash
# 2-GPU: DDP with the compressed all-reduce; new GPU tests (S2FP8 batched, float fields)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for AR in "" "--compress-allreduce p2p" "--compress-allreduce nccl"; do
  timeout 200 $T --master-port 29542 tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 30 --warmup 8 $AR > gpurun_out/ar_train.log 2>&1
  grep "^{" gpurun_out/ar_train.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('value', 'ms_per_step', 'loss', 'compress_allreduce', 'allreduce_stats')})"
  grep -B2 -A6 "Error" gpurun_out/ar_train.log | head -20
done
timeout 200 $T --master-port 29543 tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 10 --warmup 5 --profile 2>&1 | grep "profile\]" | head -8
timeout 600 python -m pytest tests/test_gpu_floatq.py tests/test_allreduce_host.py -m gpu -q --timeout=600 2>&1 | tail -8
for g in "" "--cuda-graph"; do timeout 300 python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress s2fp8 --steps 50 --warmup 10 $g 2>&1 | tail -1 | cut -c1-150; done
