#!/usr/bin/env python
"""Markdown summary of the key metrics of every kernel in an `ncu --set full` report.

  ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python profiles/tools/ncu_raw_summary.py raw.csv
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    print("## %s\n" % r[ik].replace("<unnamed>::", ""))
    print("| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in hdr:
            print("| %s | %s | %s |" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    print()
