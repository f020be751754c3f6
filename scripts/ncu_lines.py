#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
stall samples and executed warp-instructions.  usage: ncu_lines.py dump.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
cur_file = ''
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        i_samp = hdr.index('Warp Stall Sampling (All Samples)')
        i_inst = hdr.index('Instructions Executed') if 'Instructions Executed' in hdr else None
        continue
    if hdr is None or not r or not r[0] or not r[0].isdigit():
        continue
    try:
        s = int(r[i_samp])
    except Exception:
        s = 0
    try:
        ins = int(r[i_inst]) if i_inst is not None else 0
    except Exception:
        ins = 0
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0])
    a[0] += s
    a[1] += ins
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot}, total warp-instructions {toti}")
for (f, ln, src), (s, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot:5.1f}% samp {100*ins/toti:5.1f}% inst  {f}:{ln:<5d} {src}")
