#!/bin/bash
# Round 2, step o (under gpurun, 1 GPU): element-per-thread Bernstein kernel (stage_bbe, kernel 7) — parity, then the low-order
# rows of the sweep beside stage_bb2 / the generic kernel; A/B of the u-tile L2 prefetch in stage_bb2 at order 4.
mkdir -p gpurun_out
python -m pytest tests/test_zz_bb_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/r02o_tests.log
cat gpurun_out/r02o_tests.log
run() {  # tag dim order cells kernel
  python bench.py --dim $2 --order $3 --cells $4 --kernel $5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_$1_d$2p$3k$5.json 2> gpurun_out/r02o_$1_d$2p$3k$5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02o_$1_d$2p$3k$5.json")); r=d["roofline"]
    print("$1 dim $2 p$3", d["kernel"], "stage ms %.3f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("$1 dim $2 p$3 kernel $5", "failed", e)
PY
}
for v in base bbehi; do
  if [ "$v" = base ]; then unset DGB_LIB; else export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so; fi
  run $v 2 1 850 7; run $v 2 2 600 7; run $v 2 3 480 7; run $v 3 1 56 7
done
unset DGB_LIB
run base 3 4 62 6
export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_nopf.so
run nopf 3 4 62 6
unset DGB_LIB
run base2 3 4 62 6
