#!/usr/bin/env python
"""Order / dimension sweep (BASELINE config 2): python profiles/order_sweep.py [out.json]   (under gpurun, 1 GPU)
Runs bench.py on tetrahedra of order 1..6 and on a refined square of triangles of order 1..6, every mesh sized so that each
state array exceeds the 126 MB L2, with the automatic kernel choice and, where another kernel exists, that one beside it.
One row per run: kernel, ms per stage, DOF-updates/s, fraction of the HBM roof, fraction of the FP64 roof."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
out_path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "r02_order_sweep.json"
# optional: --only dim:order,dim:order  re-measures those rows and merges them into an existing file (the other rows are kept)
ONLY = None
if "--only" in sys.argv:
    ONLY = {tuple(int(x) for x in item.split(":")) for item in sys.argv[sys.argv.index("--only") + 1].split(",")}
# (dim, order, cells, [kernel ids])   kernel 0 = automatic
RUNS = [(3, 1, 56, [0, 1, 6]), (3, 2, 48, [0, 1, 7]), (3, 3, 48, [0, 3, 1]), (3, 4, 62, [0, 3]), (3, 5, 40, [0, 1]), (3, 6, 36, [0, 1]),
        (2, 1, 850, [0, 1, 6]), (2, 2, 600, [0, 1, 6]), (2, 3, 480, [0, 1, 7]), (2, 4, 400, [0, 1]), (2, 5, 340, [0, 1]), (2, 6, 300, [0, 1])]
rows = []
if ONLY is not None:  # gpurun_out/ does not travel to the GPU box: fall back to the committed table
    base = out_path if out_path.exists() else ROOT / "profiles" / "r02_order_sweep.json"
    rows = [r for r in json.loads(base.read_text())["rows"] if (r["dim"], r["order"]) not in ONLY]
for dim, order, cells, kernels in RUNS:
    if ONLY is not None and (dim, order) not in ONLY:
        continue
    for k in kernels:
        cmd = [sys.executable, str(ROOT / "bench.py"), "--dim", str(dim), "--order", str(order), "--cells", str(cells), "--steps", "5", "--warmup", "3",
               "--no-cpu-baseline", "--kernel", str(k)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            rf = d["roofline"]
            row = {"dim": dim, "order": order, "cells": cells, "requested_kernel": k, "kernel": d["kernel"], "workload": d["config"]["workload"],
                   "stage_ms": rf["stage_kernel_ms"], "dof_updates_per_s": d["value"], "hbm_frac": rf["frac"], "hbm_gbs": rf["achieved"],
                   "fp64_frac": rf["fp64"]["frac"], "fp64_tflops": rf["fp64"]["achieved"], "alg_bytes_per_launch": rf["alg_bytes_per_launch"],
                   "alg_flops_per_launch": rf["fp64"]["alg_flops_per_launch"], "finite": d["finite"], "clocks": d["clocks"]}
        except Exception as e:
            row = {"dim": dim, "order": order, "cells": cells, "requested_kernel": k, "error": str(e), "stderr": r.stderr[-400:]}
        rows.append(row)
        print(json.dumps({k2: row.get(k2) for k2 in ("dim", "order", "kernel", "stage_ms", "dof_updates_per_s", "hbm_frac", "fp64_frac", "error")}), flush=True)
        rows_sorted = sorted(rows, key=lambda r: (-r["dim"], r["order"], r["requested_kernel"]))
        out_path.write_text(json.dumps({"peaks": {"hbm_gbs": "MEASURED_PEAKS.json", "fp64_tflops": 37.1}, "rows": rows_sorted}, indent=1))
