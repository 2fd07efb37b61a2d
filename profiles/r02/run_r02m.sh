#!/bin/bash
# Round 2, step m (under gpurun, 1 GPU): parity of the Bernstein kernels after the trace-gather remap, then bench lines at
# orders 2..5 and one ncu --set full capture of the order-4 kernel.
mkdir -p gpurun_out
python -m pytest tests/test_zz_bb_gpu.py tests/test_gpu_parity.py tests/test_golden_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r02m_tests.log
cat gpurun_out/r02m_tests.log
for cfg in "4 62" "3 48" "5 40" "2 48"; do
  set -- $cfg
  python bench.py --order $1 --cells $2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_p$1.json 2> gpurun_out/r02m_p$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02m_p$1.json")); r=d["roofline"]
    print("p$1", d["kernel"], "stage ms %.3f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "hbm %.3f" % r["frac"], "finite", d["finite"])
except Exception as e:
    print("p$1", "failed", e)
PY
done
python bench.py --order 4 --cells 62 --v0 30 10 5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_p4_flow.json 2> gpurun_out/r02m_p4_flow.err
python -c "
import json; d=json.load(open('gpurun_out/r02m_p4_flow.json')); print('p4 flow', d['kernel'], d['roofline']['stage_kernel_ms'])"
bash profiles/run_full.sh r02m stageBB2
