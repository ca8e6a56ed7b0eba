#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp-instructions per element
and the stall samples attributed to each opcode.  usage: ncu_opcode_mix.py src.csv <elements>"""
import collections
import csv
import re
import sys

path, n_elem = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
samples = collections.Counter()
total = 0
tot_samples = 0
sections = 0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        sections += 1
        continue
    if sections or len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        continue
    src = r[ix["Source"]].strip()
    src = re.sub(r"^@!?U?P\d+\s+", "", src)
    op = src.split()[0].split(".")[0] if src else "?"
    ex = int(r[ix["Instructions Executed"]] or 0)
    sm = int(r[ix["# Samples"]] or 0)
    ops[op] += ex
    samples[op] += sm
    total += ex
    tot_samples += sm
print(f"total warp-instr {total}  per element (x32/n) {total * 32 / n_elem:.1f}   samples {tot_samples}")
for op, c in ops.most_common(32):
    print(f"  {op:10s} {c * 32 / n_elem:7.2f} /elem   {100 * c / total:5.1f}% of instrs   {100 * samples[op] / max(1, tot_samples):5.1f}% of stall samples")
