#!/usr/bin/env python
"""DRAM traffic of the stage kernel from an `ncu --set full` capture of consecutive stage launches (one RK4 step):
python profiles/ncu_traffic.py report.ncu-rep <elements of the profiled mesh> <kernel name as dgb_kernel_name() reports it>
Prints a JSON record {kernel: {dram_bytes_per_element, launches: [...], source}} to merge into profiles/r01_traffic.json."""
import csv
import json
import subprocess
import sys

rep, K, name = sys.argv[1], float(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def col(metric):
    return [i for i, h in enumerate(hdr) if h == metric or h.endswith("." + metric)][0]


ir, iw, it = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for r in rows[2:]:
    rd = float(r[ir]) * scale[units[ir]]
    wr = float(r[iw]) * scale[units[iw]]
    launches.append({"dram_read_bytes": rd, "dram_write_bytes": wr, "duration_us": float(r[it]) / (1e3 if units[it] == "ns" else 1.0)})
avg = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches) / len(launches)
print(json.dumps({name: {"dram_bytes_per_element": avg / K, "profiled_elements": K, "launches": launches,
                         "source": f"ncu --set full, {len(launches)} consecutive stage launches (one RK4 step) on a {int(K)}-element mesh: {rep.split('/')[-1]}"}}, indent=1))
