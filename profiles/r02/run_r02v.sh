#!/bin/bash
# Round 2, step v (under gpurun, 1 GPU): stage_bb2 with the two-pass epilogue and no u reload in the first stage: parity, bench
mkdir -p gpurun_out
python -m pytest tests/test_zz_bb_gpu.py tests/test_gpu_parity.py tests/test_golden_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r02v_tests.log
cat gpurun_out/r02v_tests.log
run() {  # tag dim order cells kernel
  python bench.py --dim $2 --order $3 --cells $4 --kernel $5 --steps 8 --warmup 3 --no-cpu-baseline $6 $7 $8 $9 > gpurun_out/r02v_$1_d$2p$3k$5.json 2> gpurun_out/r02v_$1_d$2p$3k$5.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02v_$1_d$2p$3k$5.json")); r=d["roofline"]
    print("$1 dim $2 p$3", d["kernel"], "stage ms %.4f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("$1 dim $2 p$3 kernel $5", "failed", e)
PY
}
run base 3 4 62 6; run base 3 3 48 6; run base 3 5 40 6; run base 3 2 48 6; run base 2 4 400 6; run base 2 3 480 6; run base 2 6 300 6
run flow 3 4 62 6 --v0 30 10 5
