#!/bin/bash
# N = 8 check of the final bench.py (kernel bench + DDP training block) under torchrun
NG=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
{
SECONDS=0; timeout 900 $TR bench.py --gpus $NG > gpurun_out/r2_bench_n${NG}c.json 2> gpurun_out/r2_bench_n${NG}c.err; echo "N=$NG rc $? wall ${SECONDS}s"
python - $NG <<'PY'
import json, sys
f = f"n{sys.argv[1]}c"
d = json.loads(open(f"gpurun_out/r2_bench_{f}.json").read().strip().splitlines()[-1])
t = d["train"]
print(f, "value", d["value"], "clocks", d["clocks"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_numa_binding"))
print("  train", t["img_per_s"], t["plain_img_per_s"], t["reference_eager_cuda_img_per_s"], t.get("profile", {}).get("codec_kernels_ms_per_step"), t.get("profile", {}).get("nccl_kernels_ms_per_step"), t.get("clocks"))
print("  c5", {k: v for k, v in (t.get("config5_bert_base") or {}).items() if "seq_per" in k})
PY
} > gpurun_out/run21.log 2>&1
tail -12 gpurun_out/run21.log
