#!/bin/bash
# Round 2, step n (under gpurun, 1 GPU): stage_bb2 on triangles (orders 1..6) and on tetrahedra of order 1 — parity, then the
# order sweep rows these kernels change, beside the generic kernel.
mkdir -p gpurun_out
python -m pytest tests/test_zz_bb_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/r02n_tests.log
cat gpurun_out/r02n_tests.log
run() {  # dim order cells kernel
  python bench.py --dim $1 --order $2 --cells $3 --kernel $4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_d$1p$2k$4.json 2> gpurun_out/r02n_d$1p$2k$4.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02n_d$1p$2k$4.json")); r=d["roofline"]
    print("dim $1 p$2", d["kernel"], "stage ms %.3f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("dim $1 p$2 kernel $4", "failed", e)
PY
}
run 3 1 56 6
run 3 4 62 0
for cfg in "1 850" "2 600" "3 480" "4 400" "5 340" "6 300"; do set -- $cfg; run 2 $1 $2 6; done
