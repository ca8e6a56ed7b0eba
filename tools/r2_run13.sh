#!/bin/bash
# S2FP8 through the hooks: the screened build and the exact-only build must give the SAME loss (bit-identical codec
# output); throughput with a learning rate that keeps the run finite
L=$PWD/smart-quantization_b200/smart_compress/_lib
for lib in libsmaq_b200.so libsmaq_s2exact.so; do
  echo "== $lib, lr 0.1, 20 steps"
  SMAQ_B200_LIB=$L/$lib timeout 300 python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress s2fp8 --steps 15 --warmup 5 2>&1 | grep "^{" | grep -o '"value": [^,]*\|"loss": [^,]*' | paste - -
done
for lib in libsmaq_b200.so libsmaq_s2exact.so libsmaq_b200.so libsmaq_s2exact.so; do
  for extra in "" "--cuda-graph"; do
    echo "== $lib lr 0.002 $extra"
    SMAQ_B200_LIB=$L/$lib timeout 300 python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress s2fp8 --lr 0.002 $extra 2>&1 | grep "^{" | grep -o '"value": [^,]*\|"loss": [^,]*' | paste - -
  done
done
