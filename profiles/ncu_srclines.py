#!/usr/bin/env python
"""Instructions executed and stall samples per CUDA source line, from the source correlation stored in the report itself
(captured with --import-source on; no matching .so needed): python profiles/ncu_srclines.py report.ncu-rep [kernel-index] [top] [units]
units = number of work units (tiles) of the launch: instruction counts are then printed per unit."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
by = 3 if "--by-samples" in sys.argv else 4
sys.argv = [a for a in sys.argv if a != "--by-samples"]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# blocks: "File Path", "Function Name", header, lines ...; one sequence of files per profiled launch
launches = []; cur = None; fpath = None
for i, r in enumerate(rows):
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name":
        fn = r[1]
        if cur is None or (cur["seen"].get(fpath)):  # the same file again: next launch
            cur = {"fn": fn, "seen": {}, "lines": []}; launches.append(cur)
        cur["seen"][fpath] = True; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0].isdigit():
        cur["lines"].append((fpath.split("/")[-1], int(r[0]), r[1], r))
L = launches[which]
isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
num = lambda x: int(x) if x.isdigit() else 0
agg = [(f, ln, src, num(r[isamp]), num(r[iex])) for f, ln, src, r in L["lines"]]
ts = sum(a[3] for a in agg); ti = sum(a[4] for a in agg)
print(f"launch {which}: {L['fn'][:70]}  samples {ts}  warp instructions {ti} ({ti/units:.1f} per unit)")
for f, ln, src, s, n in sorted(agg, key=lambda a: -a[by])[:top]:
    print(f"{n/units:9.1f} {100*n/ti:5.1f}% instr  {100*s/ts:5.1f}% samples  {f}:{ln}  {src.strip()[:90]}")
