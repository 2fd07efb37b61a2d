#!/usr/bin/env python
"""Top stall hot spots of an ncu report (SASS view): python profiles/ncu_hotspots.py report.ncu-rep [kernel-index] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report may hold several kernels: split at "Kernel Name" rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr = b["rows"][0]
data = b["rows"][1:]
si = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(int(r[si]) for r in data if r[si].isdigit())
print(b["name"], "total samples", tot)
agg = {}
for r in data:
    if not r[si].isdigit():
        continue
    op = r[1].split()[0] if r[1].split() else "?"
    if op.startswith("@"):
        op = r[1].split()[1]
    agg[op.split(".")[0]] = agg.get(op.split(".")[0], 0) + int(r[si])
print("by opcode:", sorted(agg.items(), key=lambda kv: -kv[1])[:12])
for r in sorted((r for r in data if r[si].isdigit()), key=lambda r: -int(r[si]))[:top]:
    reasons = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print("%6d %5.1f%%  %-70s %s" % (int(r[si]), 100.0 * int(r[si]) / tot, r[1].strip()[:70], reasons))
