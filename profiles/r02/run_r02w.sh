#!/bin/bash
# Round 2, step w (under gpurun, 1 GPU): stage_bbe at Np = 10 (triangles of order 3, tetrahedra of order 2): padded rows vs one bulk copy per tile
mkdir -p gpurun_out
python -m pytest tests/test_zz_bb_gpu.py -x -q -k "triangles_and_order_1" 2>&1 | tail -4 > gpurun_out/r02w_tests.log
cat gpurun_out/r02w_tests.log
run() {  # tag dim order cells kernel
  python bench.py --dim $2 --order $3 --cells $4 --kernel $5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_$1_d$2p$3k$5.json 2> gpurun_out/r02w_$1_d$2p$3k$5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02w_$1_d$2p$3k$5.json")); r=d["roofline"]
    print("$1 dim $2 p$3", d["kernel"], "stage ms %.4f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("$1 dim $2 p$3 kernel $5", "failed", e)
PY
}
for v in base bbec10; do
  if [ "$v" = base ]; then unset DGB_LIB; else export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so; fi
  run $v 3 2 48 7; run $v 2 3 480 7
done
