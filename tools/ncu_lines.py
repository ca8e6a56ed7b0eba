#!/usr/bin/env python
"""Executed warp-instructions per SOURCE LINE of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).

ncu lists an inlined instruction under every frame of its inline stack; each SASS address is counted ONCE here,
under its innermost project line (callees are defined before their callers in these files, so: a project header
line if there is one, else the smallest line number of the .cu file).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-regex> <elements-per-launch> [top]
"""
import collections
import csv
import io
import os
import subprocess
import sys

rep, kre, n_elem = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--csv", "--kernel-name", f"regex:{kre}", "--launch-count", "1", "--page", "source",
                      "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
text = {}
cands = collections.defaultdict(list)  # address -> [(file, line)]
execd, samples, sass = {}, {}, {}
fname, hdr, line = "", None, None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        ai, si, ex_i, sm_i = r.index("Address"), r.index("Address") + 1, r.index("Instructions Executed"), r.index("# Samples")
        hdr = True
        continue
    if hdr is None or len(r) <= ex_i:
        continue
    if r[0].strip():
        line = int(r[0])
        text[(fname, line)] = r[1].strip()[:96]
    if r[ex_i].isdigit() and r[ai]:
        cands[r[ai]].append((fname, line))
        execd[r[ai]] = int(r[ex_i])
        samples[r[ai]] = int(r[sm_i]) if r[sm_i].isdigit() else 0
        sass[r[ai]] = r[si]


def rank(c):
    f, l = c
    proj_hdr = f.endswith(".cuh")
    proj = f.endswith(".cu") or proj_hdr
    return (0 if proj_hdr else 1 if proj else 2, l)


per_line, per_line_s = collections.Counter(), collections.Counter()
for a, cs in cands.items():
    k = min(cs, key=rank)
    per_line[k] += execd[a]
    per_line_s[k] += samples[a]
tot, tots = sum(per_line.values()), max(1, sum(per_line_s.values()))
print(f"total warp-instr {tot}  per element x32: {tot * 32 / n_elem:.1f}")
for key, c in per_line.most_common(top):
    print(f"{c * 32 / n_elem:6.2f}/elem {100 * per_line_s[key] / tots:5.1f}%smp  {key[0]}:{key[1]:>4}  {text.get(key, '')}")
