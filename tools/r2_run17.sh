#!/bin/bash
# why BERT-base loses 2 % with SMAQ_DEPENDENT_LAUNCH=2: per-kernel times under torch.profiler, both modes;
# and the e2e leg with the pinned buffers bound to the GPU's NUMA node at N = 1
{
for m in 1 2; do echo "MODE $m"; SMAQ_DEPENDENT_LAUNCH=$m timeout 300 python tools/train_bench.py --model bert-base --batch 32 --compress smart --steps 20 --warmup 5 --profile 2>&1 | grep -E "profile\]|value" | cut -c1-170; done
timeout 300 python bench.py --no-sweep --no-cpu --no-train > gpurun_out/e2e_numa.json 2> gpurun_out/e2e_numa.err; tail -2 gpurun_out/e2e_numa.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/e2e_numa.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_numa_binding"])
PY
nvidia-smi topo -m | head -14
} > gpurun_out/run17.log 2>&1
tail -80 gpurun_out/run17.log
