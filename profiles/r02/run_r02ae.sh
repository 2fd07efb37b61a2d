#!/bin/bash
# Round 2, step ae (under gpurun, 1 GPU): stage_bbe — padded rows filled by cooperative cp.async (conflict-free, no copy-engine requests) instead of one bulk copy
# per tile into unpadded rows (conflicts), for Np = 6 / 10 (variant coop6) and for every Np (variant coopall)
mkdir -p gpurun_out
run() {  # tag dim order cells kernel
  timeout 120 python bench.py --dim $2 --order $3 --cells $4 --kernel $5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02ae_$1_d$2p$3k$5.json 2> gpurun_out/r02ae_$1_d$2p$3k$5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02ae_$1_d$2p$3k$5.json")); r=d["roofline"]
    print("$1 dim $2 p$3", d["kernel"], "stage ms %.4f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("$1 dim $2 p$3 kernel $5", "failed", e)
PY
}
unset DGB_LIB; run base 2 1 850 7; run base 2 2 600 7
export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_coop6.so; run coop6 2 2 600 7; run coop6 2 3 480 7
export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_coopall.so; run coopall 2 1 850 7
DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_coopall.so timeout 120 python -m pytest tests/test_zz_bb_gpu.py -x -q -k "triangles_and_order_1" 2>&1 | tail -2
