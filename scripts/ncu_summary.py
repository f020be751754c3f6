#!/usr/bin/env python
"""Key metrics per kernel launch of an .ncu-rep (ncu --page raw --csv)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_adu.sum']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx:
            print(f'{w:66s} {r[idx[w]][:70]} {units[idx[w]]}')
