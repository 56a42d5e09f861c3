#!/usr/bin/env python3
"""Selected columns of an `ncu --page raw --csv` dump (one row per captured launch), as committed under profiles/.
usage: tools/ncu_raw_summary.py <raw.csv> > summary.csv"""
import csv, sys
COLS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
idx = [names.index(c) for c in COLS if c in names]
w = csv.writer(sys.stdout)
w.writerow([names[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in rows[hdr + 2:]:
    w.writerow([r[i] for i in idx])
