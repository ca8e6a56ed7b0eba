#!/bin/bash
# round-2 GPU session 4: ncu evidence (launch list, one full capture per kernel at 2^30, L2 behaviour of a hook call)
SKIP_BENCH=1 bash tools/profile_all.sh r2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
for L in 20 22 24 26; do
  for F in "" "--flush"; do
    python tools/l2_evidence.py --log2n $L $F > /dev/null 2>&1 &&
    ncu --cache-control none --clock-control none --profile-from-start off --metrics $M --csv \
        --log-file gpurun_out/l2_${L}${F:+_flush}.csv python tools/l2_evidence.py --log2n $L $F > gpurun_out/l2_ncu.log 2>&1
  done
done
ls -la gpurun_out/ | tail -15
python -m pytest tests/test_gpu_pack.py tests/test_gpu_hooks.py -m gpu -q --timeout=600 -k "zero_on_grid or packed_saved or graph or counted" 2>&1 | tail -5
