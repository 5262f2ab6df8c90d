#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary committed under profiles/."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keep = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
idx = [hdr.index(k) for k in keep if k in hdr]
w = csv.writer(open(sys.argv[2], 'w'))
w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx])
tn = hdr.index('gpu__time_duration.sum'); kn = hdr.index('Kernel Name'); rd = hdr.index('dram__bytes_read.sum'); wr = hdr.index('dram__bytes_write.sum')
for r in rows[2:]:
    print('%-44s %8.1f us  dram %7.1f MB' % (r[kn][:44], float(r[tn]) * (1000 if units[tn] == 'ms' else 1), float(r[rd]) + float(r[wr])))
