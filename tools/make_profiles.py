#!/usr/bin/env python
"""Turn one `ncu --set full` report of tools/run_kernels.py into the committed summaries under profiles/.

    python tools/make_profiles.py gpurun_out/prof_r1s3_all.ncu-rep gpurun_out/launches_r1s3.csv r1 30
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, launches, tag, log2n = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
n = 1 << log2n
here = os.path.dirname(os.path.abspath(__file__))
out_dir = os.path.join(os.path.dirname(here), "profiles")
os.makedirs(out_dir, exist_ok=True)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
names = {"stats_kernel<0": "stats", "roundtrip_kernel": "roundtrip", "encode_kernel": "encode",
         "decode_kernel": "decode", "floatq_kernel<0": "fp8", "stats_kernel<2": "s2fp8_stats", "floatq_kernel<1": "s2fp8_apply"}
traffic = {}
seen = set()
for idx, r in enumerate(rows[2:]):
    kname = r[h.index("Kernel Name")]
    short = next((v for k, v in names.items() if k in kname), None)
    if short is None or short in seen:
        continue
    seen.add(short)
    g = lambda k: float(r[h.index(k)])  # noqa: E731
    unit = lambda k: rows[1][h.index(k)]  # noqa: E731
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = g("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
    wr = g("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    tus = g("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[unit("gpu__time_duration.sum")]
    traffic[short] = {"kernel": kname[:80], "log2n": log2n, "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                      "dram_bytes": int(rd + wr), "ncu_time_us": round(tus, 1),
                      "registers": int(g("launch__registers_per_thread")),
                      "issue_active_pct": round(g("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                      "dram_throughput_pct": round(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 1)}
    base = kname.split("(")[0].replace("void ", "").strip().split("<")[0]
    # ncu matches kernels by base name: pick this launch by its index among the launches of that name
    nth = sum(1 for rr in rows[2:2 + idx] if rr[h.index("Kernel Name")].split("(")[0].replace("void ", "").strip().split("<")[0] == base)
    with open(os.path.join(out_dir, f"{tag}_{short}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch, N = 2^{log2n} elements; source: {os.path.basename(rep)} "
                f"(launch {idx}: {kname[:60]})\n")
        rpt = subprocess.run([sys.executable, os.path.join(here, "ncu_report.py"), rep, "^" + base + "$", str(n), str(nth)],
                             capture_output=True, text=True).stdout
        f.write(rpt)
with open(os.path.join(out_dir, f"{tag}_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
# launch list: every launch of two bench steps with its device time
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hh = rows[0]
ki, vi, ui = hh.index("Kernel Name"), hh.index("Metric Value"), hh.index("Metric Unit")
with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sweep\n")
    f.write("# (cold-cache, serialised: compare SHARES, not absolutes)\n")
    tot = {}
    for r in rows[1:]:
        if "smaq" in r[ki]:
            f.write(f"{r[ki][:70]:72s} {r[vi]:>10s} {r[ui]}\n")
            key = r[ki].split("(")[0][:50]
            tot[key] = tot.get(key, 0.0) + float(r[vi])
    s = sum(tot.values())
    f.write("# share of the smaq:: launches\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"#   {k:52s} {100 * v / s:5.1f} %\n")
print(json.dumps(traffic, indent=1)[:1500])
