#!/bin/bash
# Usage (under gpurun): bash profiles/run_profile.sh <tag> [kernel-regex] [bench args, e.g. --kernel 5 --bb-tile 8]
# Produces gpurun_out/<tag>_launches.csv (every kernel launch of a short bench.py run on the headline workload with its
# device time: the kernel's share of the step) and gpurun_out/<tag>_full.ncu-rep (profiles/run_full.sh).
TAG=${1:-r01}
KRE=${2:-stage}
shift; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/${TAG}_launches_bench.log 2>&1
bash "$(dirname "$0")/run_full.sh" ${TAG} ${KRE} "$@"
ls -la gpurun_out
