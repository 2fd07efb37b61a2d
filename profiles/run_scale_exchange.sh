#!/bin/bash
# Usage (under `gpurun --gpus N`): bash profiles/run_scale_exchange.sh N [bench args]
# A/B of the halo exchange at N GPUs on config 5: NCCL send/recv (exchange 0) vs direct peer-to-peer stores (exchange 1),
# plus the 2-rank parity cases of the direct exchange that have not run on hardware yet. Writes gpurun_out/scale_n<N>_x<e>.json.
N=${1:-2}; shift
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  DGB_TEST_P2P=1 DGB_TEST_BB=1 timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "direct_exchange or bernstein" 2>&1 | tail -5
fi
for X in 0 1; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $((29600 + X)) \
    bench.py --gpus "$N" --steps 10 --warmup 3 --exchange "$X" "$@" > "gpurun_out/scale_n${N}_x${X}.json" 2> "gpurun_out/scale_n${N}_x${X}.err"
  python - "gpurun_out/scale_n${N}_x${X}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f G/s" % (d["value"] / 1e9), "ms/step %.3f" % d["ms_per_step"], "stage kernel %.3f ms" % d["roofline"]["stage_kernel_ms"], d["config"].get("exchange"))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
