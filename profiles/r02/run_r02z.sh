#!/bin/bash
# Round 2, final single-GPU evidence run (under gpurun): GPU suite, smoke, bench lines (default, mean flow, reference arm), launch list + ncu --set full
# of the headline kernel, ncu --set full of stage_bbe, order / dimension sweep, config 1 timing.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02z_gpu_suite.log; cat gpurun_out/r02z_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; cut -c1-400 gpurun_out/r02z_bench.json
python bench.py --steps 10 --warmup 3 --v0 30 10 5 --no-cpu-baseline > gpurun_out/r02z_bench_flow.json 2> gpurun_out/r02z_bench_flow.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err; cut -c1-300 gpurun_out/r02z_bench_reference.json
bash profiles/run_profile.sh r02z stageBB2 > gpurun_out/r02z_profile.log 2>&1
bash profiles/run_full.sh r02z_bbe stageBBE --dim 2 --order 1 --cells 400 >> gpurun_out/r02z_profile.log 2>&1
python profiles/order_sweep.py gpurun_out/r02z_order_sweep.json > gpurun_out/r02z_order_sweep.log 2>&1; tail -30 gpurun_out/r02z_order_sweep.log | cut -c1-200
python profiles/graph_time.py 0 2>&1 | grep "config 1" > gpurun_out/r02z_config1.txt; cat gpurun_out/r02z_config1.txt
