#!/bin/bash
# 2-GPU box: the whole GPU suite (the cuda:1 / two-rank tests are not skipped here) and the sweep's new column
{
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -4
timeout 300 python bench.py --no-train --no-cpu --no-e2e > gpurun_out/sweep_cmp.json 2> gpurun_out/sweep_cmp.err; echo "bench rc $?"; tail -2 gpurun_out/sweep_cmp.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/sweep_cmp.json").read().strip().splitlines()[-1])
for r in d["sweep"]["smaq"]: print(r["log2n"], "stats", r["stats"]["ms"], "rt", r["roundtrip"]["ms"], "compress", r["compress"])
PY
} > gpurun_out/run22.log 2>&1
tail -14 gpurun_out/run22.log
