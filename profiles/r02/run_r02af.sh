#!/bin/bash
# Round 2, step af (under gpurun, 1 GPU): occupancy target of stage_bb2<3,3>: 12 (default) vs 10 / 14 warps per SM
mkdir -p gpurun_out
for v in base p3w10 p3w14; do
  if [ "$v" = base ]; then unset DGB_LIB; else export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so; fi
  timeout 120 python bench.py --order 3 --cells 48 --kernel 6 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02af_$v.json 2> gpurun_out/r02af_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02af_$v.json')); r=d['roofline']; print('$v', d['kernel'], 'stage ms %.4f' % r['stage_kernel_ms'], 'hbm %.3f' % r['frac'])"
done
