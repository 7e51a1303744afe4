#!/usr/bin/env python
"""One-screen summary of an ncu report (raw page): duration, DRAM bytes, occupancy, pipe utilisation, stall reasons.
usage: python scripts/ncu_summary.py report.ncu-rep [title]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
print("# %s" % title)
print("# source: ncu --set full --clock-control none --import-source on ; one block per captured launch")
for r in rows[2:]:
    print("--- launch %s  %s" % (r[0], r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''))
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print("%-86s %-16s %s" % (k, units[i], r[i]))
    st = sorted(((float(r[hdr.index(h)]), h) for h in stall), reverse=True)
    for v, h in st[:8]:
        print("%-86s %-16s %.3f" % (h.replace('smsp__average_warps_issue_', '').replace('_per_issue_active.ratio', ''), 'warps/issue', v))
