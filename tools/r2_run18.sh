#!/bin/bash
# N = 2 check of the final code: bench.py (kernel bench + DDP training block) and its reference arm under torchrun,
# the compressed all-reduce against the oracle
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
{
SECONDS=0; $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "bench rc $?"
echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2b.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_numa_binding"))
t = d["train"]; print("train", t["img_per_s"], t["plain_img_per_s"], t["reference_eager_cuda_img_per_s"], t.get("clocks"))
PY
} > gpurun_out/run18.log 2>&1
tail -40 gpurun_out/run18.log
