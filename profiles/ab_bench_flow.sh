for v in "$@"; do
  if [ "$v" = base ]; then unset DGB_LIB; else export DGB_LIB=$PWD/dgfem-acoustic_b200/lib/variants/libdgb_$v.so; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --v0 30 10 0 --no-cpu-baseline > gpurun_out/abf_$v.json 2> gpurun_out/abf_$v.err || tail -2 gpurun_out/abf_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/abf_$v.json')); r=d['roofline']; print('$v', d['kernel'], 'stage ms %.3f' % r['stage_kernel_ms'], 'G/s %.1f' % (d['value']/1e9), d['finite'])"
done
