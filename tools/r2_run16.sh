#!/bin/bash
# A/B: the first kernel of a codec call as a programmatic dependent of the tensor's producer
# (SMAQ_DEPENDENT_LAUNCH=2, new) against dependent launches inside a call only (=1)
{
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -x 2>&1 | tail -3
for m in 1 2; do echo MODE $m; SMAQ_DEPENDENT_LAUNCH=$m python tools/midsize_bench.py --min 16 --max 26 --no-kernels 2>&1 | cut -c1-75 | grep -v -i warn; done
python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress smart > /dev/null 2>&1
for args in "--model resnet18 --batch 256 --image 32 --compress smart --cuda-graph" "--model resnet18 --batch 256 --image 32 --compress smart" \
            "--model resnet34 --batch 32 --image 224 --compress smart" "--model resnet18 --batch 256 --image 32 --compress fp8 --cuda-graph" \
            "--model bert-base --batch 32 --compress smart"; do
  for m in 1 2 1 2; do echo "MODE $m $args"; SMAQ_DEPENDENT_LAUNCH=$m timeout 300 python tools/train_bench.py $args 2>&1 | grep -E "value" | cut -c1-140; done
done
} > gpurun_out/run16.log 2>&1
tail -60 gpurun_out/run16.log
