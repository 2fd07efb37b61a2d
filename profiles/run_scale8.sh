#!/bin/bash
# Usage (under `gpurun --gpus 8`): bash profiles/run_scale8.sh <tag> [exchange modes, default "2 0"]
# Config 5 on 8 GPUs for each exchange mode (2 fused into the stage kernel, 1 three launches, 0 NCCL); writes gpurun_out/<tag>_n8_x<e>.json
TAG=${1:-r02}; shift
MODES=${*:-2 0}
mkdir -p gpurun_out
for X in $MODES; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800 + X)) \
    bench.py --gpus 8 --steps 20 --warmup 5 --exchange "$X" > "gpurun_out/${TAG}_n8_x${X}.json" 2> "gpurun_out/${TAG}_n8_x${X}.err"
  python - "gpurun_out/${TAG}_n8_x${X}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"].get("exchange"), "value %.1f G/s" % (d["value"] / 1e9), "e2e %.1f G/s" % (d["e2e"]["value"] / 1e9), "ms/step %.3f" % d["ms_per_step"],
          "stage kernel %.3f ms" % d["roofline"]["stage_kernel_ms"], "parity", d.get("parity", {}).get("rel_l2_vs_single"))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
