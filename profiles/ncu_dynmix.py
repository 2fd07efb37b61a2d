#!/usr/bin/env python
"""Dynamic instruction mix of one kernel of an ncu report (source page): executed warp instructions and stall samples per
opcode, per tile if a tile count is given:  python profiles/ncu_dynmix.py report.ncu-rep [kernel-index] [tiles]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = rows[start[k] + 1:(start[k + 1] if k + 1 < len(start) else len(rows))]
tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr = blk[0]
ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = {}
for r in blk[1:]:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    w = r[1].split()
    if not w:
        continue
    op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
    a = agg.setdefault(op, [0, 0])
    a[0] += int(r[ie])
    a[1] += int(r[si]) if r[si].isdigit() else 0
tot = sum(a[0] for a in agg.values())
ts = sum(a[1] for a in agg.values())
print(rows[start[k]][1], "executed %.0f (%.1f per tile), samples %d" % (tot, tot / tiles, ts))
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print("%-10s %12d  %8.1f per tile  %5.1f%% of instr  %5.1f%% of samples" % (op, a[0], a[0] / tiles, 100.0 * a[0] / tot, 100.0 * a[1] / max(ts, 1)))
