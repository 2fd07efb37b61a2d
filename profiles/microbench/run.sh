#!/bin/bash
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 250 > gpurun_out/mb_clocks.csv &
SMI=$!
./profiles/microbench/fp64_peaks.bin | tee gpurun_out/fp64_peaks.txt
python - <<'PY' | tee gpurun_out/dgemm.txt
import torch, time
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    for _ in range(3): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cuBLAS DGEMM n={n}: {best:.3f} ms {2*n**3/best*1e-9:.2f} TFLOP/s")
PY
kill $SMI
sort gpurun_out/mb_clocks.csv | uniq -c | sort -rn | head -8
