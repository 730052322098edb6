#!/usr/bin/env python
"""Print the handful of ncu metrics the roofline discussion needs from `ncu --page raw --csv` output.
usage: ncu -i prof.ncu-rep --page raw --csv | python scripts/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size']
seen = set()
for r in rows[2:]:
    name = r[idx['Kernel Name']][:60]
    if name in seen:
        continue
    seen.add(name)
    print('---', name)
    for w in want:
        if w in idx:
            print(f"  {w} [{units[idx[w]]}] {r[idx[w]]}")
    sel = [h for h in hdr if 'average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    for v, h in sorted(((float(r[idx[h]]), h) for h in sel), reverse=True)[:6]:
        print(f"  stall {v:6.3f} {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
