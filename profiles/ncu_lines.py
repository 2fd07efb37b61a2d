#!/usr/bin/env python
"""Stall samples per CUDA source line: joins the SASS view of an ncu report with nvdisasm line info of the kernel.
usage: python profiles/ncu_lines.py report.ncu-rep lib.so kernel-substring [top]"""
import csv, re, subprocess, sys, tempfile, os, glob
rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"][0]
hdr = rows[start + 1]; data = rows[start + 2:]
si = hdr.index("# Samples")
sass = [(r[1].strip(), int(r[si]) if r[si].isdigit() else 0) for r in data if len(r) > si]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    lines = None
    for cubin in glob.glob(td + "/*.cubin"):
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        # split per function
        blocks = re.split(r"\n\s*\.section\s+\.text\.", dis)
        for b in blocks:
            if kname in b.split("\n", 1)[0]:
                lines = b.splitlines()
                break
        if lines: break
cur = None; seq = []
for ln in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if m: seq.append((m.group(1).strip(), cur))
# align by order (ncu SASS rows and nvdisasm rows are the same instruction sequence)
agg = {}; n = min(len(seq), len(sass))
mismatch = sum(1 for i in range(n) if seq[i][0].split()[0].strip("@!P0123456789 ").split(".")[0] not in sass[i][0])
for i in range(n):
    agg[seq[i][1]] = agg.get(seq[i][1], 0) + sass[i][1]
tot = sum(agg.values())
print(f"instructions ncu {len(sass)} nvdisasm {len(seq)} (opcode mismatches {mismatch}), samples {tot}")
src = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        cand = glob.glob(f"/root/repo/**/{f}", recursive=True)
        src[f] = open(cand[0]).read().splitlines() if cand else []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print(f"{v:6d} {100*v/tot:5.1f}%  {f}:{l}  {text}")
