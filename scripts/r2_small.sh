#!/bin/bash
mkdir -p gpurun_out
for b in 64000 256000; do
python bench.py --batch $b --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-partition --rows 20000000 > gpurun_out/r2_small_$b.json 2> gpurun_out/r2_small_$b.err
python - $b <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r2_small_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "step", d["ms_per_step"], d["roofline"]["phase_ms"], "frac", d["roofline"]["frac"], "launches/step", d["gpu_launches"]/d["steps_timed"])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_64k.csv python bench.py --batch 64000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 8000000 --min-seconds 0.01 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_r2_64k.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    d=dict(zip(H,r))
    agg.setdefault(d['Kernel Name'][:70],[]).append(float(d['Metric Value'].replace(',','')))
for k,v in agg.items(): print(f"{sum(v)/len(v)/1000:9.1f} us x{len(v)}  {k}")
PY
