#!/usr/bin/env python
"""Markdown per-kernel table from an `ncu -i X.ncu-rep --page raw --csv` dump (the summary committed under profiles/).
usage: profile_summary.py <raw.csv> <width> <height> [peak_gbs]"""
import csv, sys, json, os
rows = list(csv.reader(open(sys.argv[1])))
W, H = int(sys.argv[2]), int(sys.argv[3])
peak = float(sys.argv[4]) if len(sys.argv) > 4 else 6546.9
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return d
print("| kernel | grid x block | us | DRAM read+write MB | DRAM GB/s | % of measured peak | warp instr (M) | instr / px | issue active % | warps active % | regs | shared-memory pipe % | XU % | L2 hit % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
tot_us = tot_inst = 0.0
for r in rows[2:]:
    name = r[ix['Kernel Name']].replace('asvgf::', '').replace('(bool)', '').replace('(int)', '')
    name = name.split('(')[0].replace('void ', '')
    us = g(r, 'gpu__time_duration.sum') * (1000.0 if units[ix['gpu__time_duration.sum']] == 'ms' else 1.0)
    mb = (g(r, 'dram__bytes_read.sum') + g(r, 'dram__bytes_write.sum'))
    for k in ('dram__bytes_read.sum',):
        if units[ix[k]].lower().startswith('gbyte'): mb *= 1000.0
        if units[ix[k]].lower().startswith('kbyte'): mb /= 1000.0
    inst = g(r, 'smsp__inst_executed.sum')
    px = W * H if 'gradient' not in name else ((W + 2) // 3) * ((H + 2) // 3) * 9
    gbs = mb / 1e3 / (us * 1e-6) if us > 0 else 0.0
    tot_us += us; tot_inst += inst
    print("| `%s` | %s x %s | %.1f | %.1f | %.0f | %.1f | %.1f | %.0f | %.1f | %.1f | %d | %.1f | %.1f | %.1f |" % (
        name, r[ix['Grid Size']], r[ix['Block Size']], us, mb, gbs, 100.0 * gbs / peak, inst / 1e6, inst * 32.0 / px,
        g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        int(g(r, 'launch__registers_per_thread')), g(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
        g(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'), g(r, 'lts__t_sector_hit_rate.pct')))
print()
print("Frame total under ncu: %.1f us, %.0f M warp instructions = %.0f instructions per pixel; at 148 SMs x 4 schedulers x 1.965 GHz "
      "that is %.3f ms at 100 %% issue efficiency." % (tot_us, tot_inst / 1e6, tot_inst * 32.0 / (W * H), tot_inst / (148 * 4 * 1.965e9) * 1e3))
