#!/usr/bin/env python
"""Dynamic warp-instruction counts per CUDA source line (same join as ncu_lines.py).
usage: python profiles/ncu_instr.py report.ncu-rep lib.so kernel-substring units [top]"""
import csv, re, subprocess, sys, tempfile, os, glob
rep, lib, kname, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"][0]
hdr = rows[start + 1]; data = rows[start + 2:]
ii = hdr.index("Instructions Executed")
sass = [(r[1].strip(), int(r[ii]) if r[ii].isdigit() else 0) for r in data if len(r) > ii]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    lines = None
    for cubin in glob.glob(td + "/*.cubin"):
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        for b in re.split(r"\n\s*\.section\s+\.text\.", dis):
            if kname in b.split("\n", 1)[0]:
                lines = b.splitlines(); break
        if lines: break
cur = None; seq = []
for ln in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if m: seq.append((m.group(1).strip(), cur))
agg = {}; ops = {}
for (ins, loc), (txt, cnt) in zip(seq, sass):
    agg[loc] = agg.get(loc, 0) + cnt
    op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
    ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + cnt
tot = sum(agg.values())
print(f"warp-instructions per unit: {tot/units:.0f}")
print("by opcode per unit:", ", ".join(f"{k} {v/units:.0f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]))
src = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        cand = glob.glob(f"/root/repo/**/{f}", recursive=True)
        src[f] = open(cand[0]).read().splitlines() if cand else []
    text = src[f][l - 1].strip()[:95] if 0 < l <= len(src[f]) else ""
    print(f"{v/units:7.1f}  {f}:{l}  {text}")
