#!/bin/bash
# quick GPU check: roundtrip/pack parity tests (default widths) + kernel timings at 2^28
timeout 300 python -m pytest tests/test_gpu_smaq.py tests/test_gpu_pack.py -m gpu -q --timeout=120 -x -k "not other_bit_widths and not golden_inputs_through and not full_size" 2>&1 | tail -3
timeout 200 python bench.py --log2n 28 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/quick.json 2> gpurun_out/quick.err; tail -2 gpurun_out/quick.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/quick.json").read().strip().splitlines()[-1])
print("step", d["value"], d["kernels"])
r=d["sweep"]["smaq"][-1]; print(r["log2n"], {k:(v["ms"],v["gbs"],v["frac"]) for k,v in r.items() if k!="log2n"})
print(d["sweep"]["float_emulation"])
PY
