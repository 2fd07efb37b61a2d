#!/bin/bash
# Usage (under gpurun): [KERNEL=4] bash profiles/ab_bench.sh <cells> <variant-name|base> ...   -> one line per variant (KERNEL: bench.py --kernel)
CELLS=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset DGB_LIB; else export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so; fi
  timeout 300 python bench.py --cells $CELLS --steps 8 --warmup 3 --no-cpu-baseline --kernel ${KERNEL:-0} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -2 gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$v.json")); r=d["roofline"]
    print("$v", d["kernel"], "stage ms %.3f" % r["stage_kernel_ms"], "G/s %.1f" % (d["value"]/1e9), "finite", d["finite"])
except Exception as e:
    print("$v", "failed", e)
PY
done
