#!/bin/bash
# S2FP8 screened apply path: full float-quantize parity file, ncu capture of the two S2FP8 kernels, training steps
timeout 900 python -m pytest tests/test_gpu_floatq.py -m gpu -q --timeout=600 -s 2>&1 | grep -E "lg2 abs|passed|failed|Error|assert" | head -20
timeout 300 python tools/run_kernels.py --log2n 30 --reps 2 --only s2fp8 > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"_kernel" -c 4 -o gpurun_out/prof_r2_s2 -f \
    python tools/run_kernels.py --log2n 30 --reps 2 --only s2fp8 > gpurun_out/ncu_full_s2.log 2>&1
tail -2 gpurun_out/ncu_full_s2.log
for extra in "" "--cuda-graph"; do
  timeout 600 python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress s2fp8 $extra 2>&1 | grep "^{" | cut -c1-300
done
