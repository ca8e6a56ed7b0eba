#!/bin/bash
# round-2 GPU session 5: the one-launch compress (statistics + round trip behind a ticket barrier): tests first, each
# under its own timeout (a barrier that never completes traps after ~1 s; the timeout is the second belt), then A/B
set -u
timeout 600 python -m pytest tests/test_gpu_smaq.py tests/test_gpu_edge_cases.py tests/test_gpu_hooks.py -m gpu -q --timeout=300 -x 2>&1 | tail -6
echo "--- smaq_compress per call, one launch (default) vs two launches (SMAQ_FUSED_MAX_LOG2N=0)"
timeout 300 python tools/midsize_bench.py --min 16 --max 26 --no-kernels 2>&1 | tail -12
SMAQ_FUSED_MAX_LOG2N=0 timeout 300 python tools/midsize_bench.py --min 16 --max 26 --no-kernels 2>&1 | tail -12
echo "--- training, one launch vs two"
for F in 24 0; do
  SMAQ_FUSED_MAX_LOG2N=$F timeout 600 python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress smart --steps 50 --warmup 10 --cuda-graph 2>&1 | tail -1 | cut -c1-140
  SMAQ_FUSED_MAX_LOG2N=$F timeout 600 python tools/train_bench.py --model resnet18 --batch 256 --image 32 --compress smart --steps 50 --warmup 10 2>&1 | tail -1 | cut -c1-140
  SMAQ_FUSED_MAX_LOG2N=$F timeout 600 python tools/train_bench.py --model resnet34 --batch 32 --image 224 --compress smart --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-140
  SMAQ_FUSED_MAX_LOG2N=$F timeout 600 python tools/train_bench.py --model bert-base --batch 32 --seq 128 --compress smart --steps 30 --warmup 5 --cuda-graph 2>&1 | tail -1 | cut -c1-140
done
echo "--- L2 behaviour of a hook call, application replay (every pass sees the natural cache state)"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for L in 22 24; do
  for F in 24 0; do
    SMAQ_FUSED_MAX_LOG2N=$F ncu --replay-mode application --cache-control none --clock-control none --profile-from-start off --metrics $M --csv \
        --log-file gpurun_out/l2app_${L}_fused${F}.csv python tools/l2_evidence.py --log2n $L > gpurun_out/l2_ncu.log 2>&1
  done
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/l2app_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    h = rows[0]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = {}
    for r in rows[1:]:
        out.setdefault(r[ki].split("(")[0].replace("void ", "")[:36], {})[r[mi]] = r[vi] + " " + r[ui]
    print(f)
    for k, m in out.items():
        print("   ", k, m)
PY
