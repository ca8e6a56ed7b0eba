#!/bin/bash
# What the driver runs at round end, on one box: the whole GPU suite, smoke(), bench.py and its reference arm.
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc $? wall ${SECONDS}s"; tail -c 400 gpurun_out/r2_bench_final.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("kernels", d["kernels"]); print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic")})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step")}, "cpu", d["cpu_baseline"]["value"])
print("float", d["sweep"]["float_emulation"])
print("train", json.dumps(d.get("train"))[:1500])
r = json.loads(open("gpurun_out/r2_bench_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["unit"], r.get("cpu_baseline"))
PY
