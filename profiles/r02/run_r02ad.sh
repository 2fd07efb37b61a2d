#!/bin/bash
# Round 2, step ad (under gpurun, 1 GPU): occupancy targets of stage_bb2 on triangles of order 4 / 5 / 6: 12 / 12 / 8 (default) vs 16 / 14 / 10 (variant wa)
mkdir -p gpurun_out
run() {  # tag dim order cells kernel
  python bench.py --dim $2 --order $3 --cells $4 --kernel $5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02ad_$1_d$2p$3k$5.json 2> gpurun_out/r02ad_$1_d$2p$3k$5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02ad_$1_d$2p$3k$5.json")); r=d["roofline"]
    print("$1 dim $2 p$3", d["kernel"], "stage ms %.4f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("$1 dim $2 p$3 kernel $5", "failed", e)
PY
}
for v in base wa; do
  if [ "$v" = base ]; then unset DGB_LIB; else export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so; fi
  run $v 2 3 480 6; run $v 2 4 400 6; run $v 2 5 340 6; run $v 2 6 300 6
done
