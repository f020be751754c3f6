# per-launch times of the sort kernels (ncu serialises and runs cold: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"radix|DeviceScan|Onesweep|Histogram" -c 40 --csv --log-file gpurun_out/sort_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 8000000 > gpurun_out/sort_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/sort_ncu.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]
agg=collections.defaultdict(list)
for r in rows[hdr+1:]:
    d=dict(zip(H,r))
    agg[d['Kernel Name'][:70]].append(float(d['Metric Value'].replace(',','')))
for k,v in agg.items(): print(f"{sum(v)/len(v)/1000:9.1f} us x{len(v)}  {k}")
PY
