#!/bin/bash
# Round 2, step q (under gpurun, 1 GPU): the whole GPU suite with the new automatic kernel choices, low-order bench rows, config 1 timing per kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02q_tests.log
cat gpurun_out/r02q_tests.log
bash profiles/r02/run_r02p.sh base 2>&1 | sed 's/^/auto-or-7: /'
python profiles/graph_time.py 1 6 7 2>&1 | grep "config 1" > gpurun_out/r02q_graph_time.txt
cat gpurun_out/r02q_graph_time.txt
