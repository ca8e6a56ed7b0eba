#!/bin/bash
# full GPU suite + smoke + default bench on one GPU (what the driver runs at round end)
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_bench_final.err | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["kernels"], d["roofline"], d["encode_decode_gbs"], d["encode_decode_frac_of_measured_peak"], d["encode_decode_frac_of_nominal_8000"])
print(d["e2e"]); print(d["cpu_baseline"]); print(d["clocks"], d["gpu_launches"])
print(json.dumps(d["train"])[:2500])
PY
