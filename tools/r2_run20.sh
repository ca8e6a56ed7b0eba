#!/bin/bash
# bench.py after the sampler / profile-last changes: N = 1 and N = 2 on one box
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
{
SECONDS=0; timeout 900 python bench.py > gpurun_out/r2_bench_n1c.json 2> gpurun_out/r2_bench_n1c.err; echo "N=1 rc $? wall ${SECONDS}s"
SECONDS=0; timeout 900 $TR bench.py --gpus 2 > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; echo "N=2 rc $? wall ${SECONDS}s"
python - <<'PY'
import json
for f in ("n1c", "n2c"):
    d = json.loads(open(f"gpurun_out/r2_bench_{f}.json").read().strip().splitlines()[-1])
    t = d["train"]
    print(f, "value", d["value"], "clocks", d["clocks"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("host_numa_binding"))
    print("  train", t["img_per_s"], t["plain_img_per_s"], t["reference_eager_cuda_img_per_s"], t.get("profile", {}).get("codec_kernels_ms_per_step"), t.get("profile", {}).get("nccl_kernels_ms_per_step"))
    print("  c3", {k: v for k, v in (t.get("config3_resnet18_cifar") or {}).items() if "img" in k})
    print("  c5", {k: v for k, v in (t.get("config5_bert_base") or {}).items() if "seq_per" in k})
PY
} > gpurun_out/run20.log 2>&1
tail -30 gpurun_out/run20.log
