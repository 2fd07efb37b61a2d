#!/bin/bash
# Round 2, step t (under gpurun, 1 GPU): stage_bb2 with acc in global memory (variant accg: u and the next stage input never waited for) vs the default build
mkdir -p gpurun_out
DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_accg.so python -m pytest tests/test_zz_bb_gpu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4 > gpurun_out/r02t_tests_accg.log
cat gpurun_out/r02t_tests_accg.log
run() {  # tag dim order cells kernel
  python bench.py --dim $2 --order $3 --cells $4 --kernel $5 --steps 8 --warmup 3 --no-cpu-baseline $6 $7 $8 $9 > gpurun_out/r02t_$1_d$2p$3k$5.json 2> gpurun_out/r02t_$1_d$2p$3k$5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02t_$1_d$2p$3k$5.json")); r=d["roofline"]
    print("$1 dim $2 p$3", d["kernel"], "stage ms %.3f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("$1 dim $2 p$3 kernel $5", "failed", e)
PY
}
for v in base accg; do
  if [ "$v" = base ]; then unset DGB_LIB; else export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so; fi
  run $v 3 4 62 6; run $v 3 3 48 6; run $v 3 5 40 6; run $v 3 2 48 6; run $v 2 4 400 6; run $v 2 3 480 6
done
