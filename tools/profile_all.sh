#!/bin/bash
# GPU-box script: default bench, per-launch timing list of the same command, and one --set full capture of
# every kernel of the hot path at 2^30 elements.  Outputs under gpurun_out/ (copied into profiles/ by hand).
set -u
R=${1:-r1}
if [ -z "${SKIP_BENCH:-}" ]; then
python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err || exit 1
tail -c 600 gpurun_out/bench_${R}.err
fi
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sweep > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sweep > gpurun_out/ncu_launches.log 2>&1
python tools/run_kernels.py --log2n 30 --reps 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"_kernel" -c 12 -o gpurun_out/prof_${R}_all -f \
    python tools/run_kernels.py --log2n ${LOG2N:-30} --reps 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
