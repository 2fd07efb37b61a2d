#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report: python profiles/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_active.avg", "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio"]
for r in rows[2:]:
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w or h.endswith("." + w):
                print(f"{h} [{units[i]}] = {r[i]}")
                break
    st = sorted(((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                 for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and r[i]), reverse=True)[:6]
    print("stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st))
    print("---")
