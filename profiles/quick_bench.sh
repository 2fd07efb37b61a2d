#!/bin/bash
# Usage (under gpurun): bash profiles/quick_bench.sh [bench args]  -> tiled-kernel tests + one-line bench summary
python -m pytest tests/test_gpu_parity.py -x -q -k "tiled or rk4 or golden" 2>&1 | tail -3
python bench.py --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print(d['kernel'], 'value %.2f G/s'%(d['value']/1e9), 'e2e %.2f G/s'%(d['e2e']['value']/1e9), 'stage ms %.3f'%r['stage_kernel_ms'], 'hbm frac %.3f fp64 frac %.3f'%(r['frac'], r['fp64']['frac']), d['clocks'])"
