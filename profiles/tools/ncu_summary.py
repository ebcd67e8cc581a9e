#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page CSV on stdin or path argument): the metrics the roofline uses.

    ncu -i X.ncu-rep --page raw --csv | python profiles/tools/ncu_summary.py
"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("Kernel:", r[hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            print("  %-82s %-14s %s" % (k, units[hdr.index(k)], r[hdr.index(k)]))
    st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print("  stall %-76s %.3f" % (h[len(STALL):-len("_per_issue_active.ratio")], v))
