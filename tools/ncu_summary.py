#!/usr/bin/env python
"""Print the headline raw metrics of every kernel in an ncu report.  Usage: tools/ncu_summary.py rep"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "sm__ctas_launched.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:80])
    for w in WANT:
        if w in h:
            print(f"  {w:72s} {r[h.index(w)]:>16s} {units[h.index(w)]}")
