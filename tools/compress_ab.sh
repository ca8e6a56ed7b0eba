#!/bin/bash
# GPU-box script: A/B of smaq_compress's launch form (SMAQ_COMPRESS_MODE=0: two ordinary launches, 1 (default): the
# round trip as a programmatic dependent launch behind the statistics kernel) — parity tests, per-call device time
# at the training sizes, and the ResNet training steps.
# Output: gpurun_out/compress_ab.txt
{
for m in 0 1; do SMAQ_COMPRESS_MODE=$m python -m pytest tests/test_gpu_edge_cases.py -m gpu -x -q --timeout=300 2>&1 | tail -1; done
for m in 0 1; do echo MODE $m; SMAQ_COMPRESS_MODE=$m python tools/midsize_bench.py --min 16 --max 28 --no-kernels 2>&1 | cut -c1-75 | grep -v -i warn; done
for m in 0 1 0 1; do echo MODE $m; SMAQ_COMPRESS_MODE=$m python tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart 2>&1 | grep -E "value" | cut -c1-140; done
for m in 0 1 0 1; do echo MODE $m; SMAQ_COMPRESS_MODE=$m python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress smart 2>&1 | grep -E "value" | cut -c1-140; done
} > gpurun_out/compress_ab.txt 2>&1
tail -70 gpurun_out/compress_ab.txt
