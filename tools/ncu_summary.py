#!/usr/bin/env python
"""Markdown table of the judged metrics of one .ncu-rep (first profiled launch).  usage: ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.avg", "lts__t_sectors.max",
        "lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
r = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, u, v = r[0], r[1], r[2]
print("kernel:", v[h.index("Kernel Name")])
print("| metric | value |\n|---|---|")
for k in KEYS:
    if k in h:
        print("| %s | %s %s |" % (k, v[h.index(k)], u[h.index(k)]))
for i, k in enumerate(h):
    if "tensor" in k and k not in KEYS and v[i] not in ("0", "", "n/a") and "peak_sustained" not in k.split(".")[-1] \
            and ".per_second" not in k and not k.startswith("device__"):
        print("| %s | %s %s |" % (k, v[i], u[i]))
