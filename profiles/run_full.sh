#!/bin/bash
# Usage (under gpurun): bash profiles/run_full.sh <tag> <kernel-regex> [bench args...]
# One `ncu --set full` capture of FOUR consecutive stage launches (the RK1..RK4 stages of the second time step) on a
# 32^3 x 6 = 196 608 element mesh, so that the replays stay short.
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 4 -c 4 -f -o gpurun_out/${TAG}_full \
    python bench.py --cells 32 --steps 3 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_full_bench.log 2>&1
tail -2 gpurun_out/${TAG}_full_bench.log | cut -c1-300
