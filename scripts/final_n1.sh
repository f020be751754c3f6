#!/bin/bash
# round-2 final single-GPU evidence: bench lines, reference arm, launch list, full ncu capture, smoke, tests
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 8000000 --min-seconds 0.01"
SFM_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv $B > /dev/null 2>&1
SFM_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"bkt_pull|bkt_scatter|fm_forward_onehot16|bkt_count|bkt_offsets" --launch-skip 15 -c 5 -o gpurun_out/r2_final_hot $B > gpurun_out/r2_final_ncu.log 2>&1
ls -la gpurun_out/r2_final_hot.ncu-rep
python __graft_entry__.py smoke > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_final_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2_final_tests.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "csr", d["e2e_csr"]["value"], "part", d["partition_sampler"]["value"], "predict", d["predict"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
r=json.loads(open("gpurun_out/bench_r2_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
