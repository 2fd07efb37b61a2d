#!/bin/bash
# Usage (under gpurun, 1 GPU): bash profiles/run_unverified.sh        (about 6 minutes)
# First hardware run of what was written after round 1's GPU budget was spent:
#   1. the gated parity tests of the Bernstein-Bezier kernels and of the curved-element kernel (their sources already pass on
#      the CPU through the CUDA emulation of oracle/cuda_emu.h);
#   2. config 5 with the two Bernstein schedules (--kernel 4, 5) at 32 / 16 / 8 elements per CTA beside the shipped
#      warp-specialised kernel (--kernel 3).
# Writes gpurun_out/{bb_tests,curved_tests}.log and gpurun_out/{bb,bbseq,ws}_bench.json. The 2-GPU pieces (direct halo
# exchange, partitioned Bernstein runs) are in profiles/run_scale_exchange.sh.
mkdir -p gpurun_out
DGB_TEST_BB=1 timeout 900 python -m pytest tests/test_zz_bb_gpu.py -x -q 2>&1 | tee gpurun_out/bb_tests.log | tail -15
DGB_TEST_CURVED=1 timeout 900 python -m pytest tests/test_curved.py -x -q -m gpu 2>&1 | tee gpurun_out/curved_tests.log | tail -15
DGB_TEST_CLI=1 timeout 600 python -m pytest tests/test_zz_cli_gpu.py -x -q 2>&1 | tee gpurun_out/cli_tests.log | tail -8
for KT in 4:32 5:32 4:8 5:8 5:16 3:0; do
  K=${KT%%:*}; T=${KT##*:}
  name=$([ $K = 4 ] && echo bb$T || ([ $K = 5 ] && echo bbseq$T || echo ws))
  timeout 900 python bench.py --no-cpu-baseline --kernel $K --bb-tile $T > gpurun_out/${name}_bench.json 2> gpurun_out/${name}_bench.err
  python - gpurun_out/${name}_bench.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(d["kernel"], "value %.2f G/s" % (d["value"] / 1e9), "stage %.3f ms" % r["stage_kernel_ms"], "hbm frac %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
