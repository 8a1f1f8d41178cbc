#!/usr/bin/env python
"""Condense `ncu --page raw --csv` dumps into a small per-kernel summary table (profiles/*.md)."""
import csv, sys, glob, os
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
out = []
for f in sorted(glob.glob(sys.argv[1])):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        out.append("### %s  (%s)\n" % (r[hdr.index('Kernel Name')][:90], os.path.basename(f)))
        for k in KEYS:
            if k in hdr:
                out.append("- `%s` = %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        out.append("")
print("\n".join(out))
