#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep: headline counters, stall breakdown, opcode mix per element.

    python tools/ncu_report.py gpurun_out/prof.ncu-rep <kernel-regex> <elements-per-launch> [launch-index]
"""
import csv
import io
import subprocess
import sys

rep, kre, n_elem = sys.argv[1], sys.argv[2], float(sys.argv[3])
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
base = ["ncu", "-i", rep, "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip, "--launch-count", "1"]
raw = subprocess.run(base + ["--page", "raw"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, r = rows[0], rows[1], rows[2]
get = lambda k: r[h.index(k)] if k in h else "n/a"  # noqa: E731
print("kernel:", get("Kernel Name")[:100])
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"):
    if k in h:
        print(f"  {k:78s} {get(k)} {u[h.index(k)]}")
inst = float(get("smsp__inst_executed.sum"))
print(f"  warp-instructions per element x32: {inst * 32 / n_elem:.1f}")
print("  stalls (warps per issue-active cycle):")
st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
      for i, k in enumerate(h) if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for k, v in sorted(st, key=lambda t: -t[1])[:9]:
    print(f"    {k:28s} {v:.3f}")
src = subprocess.run(base + ["--page", "source"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
subprocess.run([sys.executable, __file__.replace("ncu_report.py", "ncu_opcode_mix.py"), "/tmp/_src.csv", str(n_elem)])
