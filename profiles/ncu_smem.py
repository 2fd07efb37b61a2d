#!/usr/bin/env python
"""Shared-memory wavefronts (total / excessive = bank-conflict replays) per CUDA source line.
usage: python profiles/ncu_smem.py report.ncu-rep lib.so kernel-substring units [top]"""
import csv, re, subprocess, sys, tempfile, os, glob
rep, lib, kname, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"][0]
hdr = rows[start + 1]; data = rows[start + 2:]
iw, ie = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
num = lambda x: int(x) if x.isdigit() else 0
sass = [(num(r[iw]), num(r[ie])) for r in data if len(r) > max(iw, ie)]
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
    if re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln): seq.append(cur)
agg = {}
for loc, (w, e) in zip(seq, sass):
    a = agg.setdefault(loc, [0, 0]); a[0] += w; a[1] += e
tw, te = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"shared wavefronts per unit: {tw/units:.0f}, of which excessive (bank conflicts): {te/units:.0f}")
src = {}
for (f, l), (w, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        cand = glob.glob(f"/root/repo/**/{f}", recursive=True)
        src[f] = open(cand[0]).read().splitlines() if cand else []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print(f"{w/units:7.1f} {e/units:7.1f}  {f}:{l}  {text}")
