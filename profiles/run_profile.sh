#!/bin/bash
# Usage (under gpurun): bash profiles/run_profile.sh <tag> [kernel-regex]
# Produces gpurun_out/<tag>_launches.csv (every launch with its device time) and gpurun_out/<tag>_full.ncu-rep
# (one --set full capture of the stage kernel on a smaller mesh so that the replays stay short).
TAG=${1:-r01}
KRE=${2:-stage}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --cells 40 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 4 -c 2 -f -o gpurun_out/${TAG}_full \
    python bench.py --cells 32 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_full_bench.log 2>&1
ls -la gpurun_out
