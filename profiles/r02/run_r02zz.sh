#!/bin/bash
# Round 2, final check of the final build on one GPU (under gpurun): GPU suite, smoke, bench lines, launch list + ncu of the headline kernel, the sweep rows that changed
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02zz_gpu_suite.log; cat gpurun_out/r02zz_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02zz_smoke.log 2>&1; tail -2 gpurun_out/r02zz_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02zz_bench.json 2> gpurun_out/r02zz_bench.err; cut -c1-330 gpurun_out/r02zz_bench.json
python bench.py --steps 10 --warmup 3 --v0 30 10 5 --no-cpu-baseline > gpurun_out/r02zz_bench_flow.json 2> gpurun_out/r02zz_bench_flow.err
bash profiles/run_profile.sh r02zz stageBB2 > gpurun_out/r02zz_profile.log 2>&1
python profiles/order_sweep.py gpurun_out/r02_order_sweep_merge.json --only 3:3,3:4,2:3,2:4,2:5,2:6 > gpurun_out/r02zz_order_sweep.log 2>&1; grep -c dim gpurun_out/r02zz_order_sweep.log
