#!/usr/bin/env python
"""print the metrics we track from an .ncu-rep (raw page csv on stdin)"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'sass__inst_executed_global_loads', 'sass__inst_executed_shared_loads', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed', 'sm__cycles_elapsed.max']
stalls = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(f"{w:80s} {r[hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stalls), reverse=True)[:7]
    print("top stalls:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st))
    print()
