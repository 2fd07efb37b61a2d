#!/bin/bash
# Usage (under gpurun): bash profiles/run_full.sh <tag> <kernel-regex> [bench args...]
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 4 -c 1 -f -o gpurun_out/${TAG}_full \
    python bench.py --cells 32 --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_full_bench.log 2>&1
tail -2 gpurun_out/${TAG}_full_bench.log | cut -c1-300
