#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --batch 125000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 8000000 --min-seconds 0.01"
SFM_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_125k.csv $B > /dev/null 2>&1
SFM_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"bkt_pull" --launch-skip 4 -c 1 -o gpurun_out/r2_pull_125k $B > gpurun_out/r2_ncu_small.log 2>&1
ls -la gpurun_out/r2_pull_125k.ncu-rep
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_r2_125k.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    d=dict(zip(H,r))
    agg.setdefault(d['Kernel Name'][:70],[]).append(float(d['Metric Value'].replace(',','')))
for k,v in agg.items(): print(f"{sum(v)/len(v)/1000:9.1f} us x{len(v)}  {k}")
PY
