#!/bin/bash
# round-2 GPU session: default bench (with the training block), A/B of the round trip's rounding form, full-size tests
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -3 gpurun_out/r2_bench2.err
SMAQ_B200_LIB=$PWD/smart-quantization_b200/smart_compress/_lib/libsmaq_rt0.so python bench.py --steps 5 --warmup 3 --no-train --no-e2e --no-cpu > gpurun_out/r2_bench2_rt0.json 2> gpurun_out/r2_bench2_rt0.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench2.json", "gpurun_out/r2_bench2_rt0.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d["value"], {k: v["ms"] for k, v in d["kernels"].items()})
    for r in d["sweep"]["smaq"]:
        print("  ", r["log2n"], {k: (v["ms"], v["frac"]) for k, v in r.items() if k != "log2n"})
    print("  ", d["sweep"].get("float_emulation")); print("  ", d["sweep"].get("pure_normal_input")); print("  ", d["sweep"].get("use_sample_stats"))
    if "train" in d: print("  train", json.dumps(d["train"]))
    if "e2e" in d: print("  e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
python -m pytest tests/test_gpu_pack.py -m gpu -q --timeout=900 -k "full_size or 2p26" 2>&1 | tail -5
