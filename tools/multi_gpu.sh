#!/bin/bash
# GPU-box script (gpurun --gpus N): bench.py and the DDP training steps under torchrun on N GPUs of one node.
#   bash tools/multi_gpu.sh <N> [all|resnet|bench]        Output: gpurun_out/multi_gpu_<N>.txt
N=${1:-8}
WHAT=${2:-all}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
if [ "$WHAT" = all ] || [ "$WHAT" = bench ]; then
echo "# bench.py --gpus $N"
$TR bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1
fi
if [ "$WHAT" != bench ]; then
echo "# tools/train_bench.py resnet34 224x224 batch 32/GPU, DDP, $N GPUs"
$TR tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 10 --warmup 4 2>/dev/null | grep value
fi
if [ "$WHAT" = all ]; then
echo "# tools/train_bench.py bert-base seq 128 batch 32/GPU, DDP, $N GPUs"
$TR tools/train_bench.py --model bert-base --batch 32 --compress smart --steps 10 --warmup 4 2>/dev/null | grep value
fi
} > gpurun_out/multi_gpu_$N.txt 2>&1
cat gpurun_out/multi_gpu_$N.txt | cut -c1-400
