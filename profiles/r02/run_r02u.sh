#!/bin/bash
# Round 2, step u (under gpurun, 1 GPU): one vs two trace buffers where not yet measured; ncu --set full of the current order-4 kernel
mkdir -p gpurun_out
run() {  # tag dim order cells kernel
  python bench.py --dim $2 --order $3 --cells $4 --kernel $5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_$1_d$2p$3k$5.json 2> gpurun_out/r02u_$1_d$2p$3k$5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02u_$1_d$2p$3k$5.json")); r=d["roofline"]
    print("$1 dim $2 p$3", d["kernel"], "stage ms %.4f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("$1 dim $2 p$3 kernel $5", "failed", e)
PY
}
for v in tb1 tb2; do
  export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so
  run $v 2 3 480 6; run $v 2 5 340 6; run $v 2 2 600 6; run $v 3 1 56 6
done
unset DGB_LIB
bash profiles/run_full.sh r02u stageBB2
