#!/bin/bash
# GPU-box script: A/B of the programmatic dependent launches (SMAQ_DEPENDENT_LAUNCH=0: ordinary launches; 1, the
# default: statistics -> round trip, encode pass 1 -> pass 2 and S2FP8 statistics -> apply as dependent launches) —
# parity tests both ways, per-call device time at the training sizes, and training steps (alternating, after one
# discarded warm-up run: the first training run on a fresh box is ~10 % slow).
# Output: gpurun_out/compress_ab.txt
{
for m in 0 1; do SMAQ_DEPENDENT_LAUNCH=$m python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_pack.py tests/test_gpu_floatq.py -m gpu -x -q --timeout=300 2>&1 | tail -1; done
for m in 0 1; do echo MODE $m; SMAQ_DEPENDENT_LAUNCH=$m python tools/midsize_bench.py --min 16 --max 28 --no-kernels 2>&1 | cut -c1-75 | grep -v -i warn; done
python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress smart > /dev/null 2>&1
for args in "--model resnet18 --batch 256 --image 32 --compress smart" "--model resnet18 --batch 256 --image 32 --compress s2fp8" \
            "--model resnet34 --batch 32 --image 224 --compress smart" "--model resnet34 --batch 32 --image 224 --compress smart --packed-activations" \
            "--model bert-base --batch 32 --compress smart"; do
  for m in 0 1 0 1; do echo "MODE $m $args"; SMAQ_DEPENDENT_LAUNCH=$m python tools/train_bench.py $args 2>&1 | grep -E "value" | cut -c1-140; done
done
} > gpurun_out/compress_ab.txt 2>&1
tail -70 gpurun_out/compress_ab.txt
