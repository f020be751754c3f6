#!/bin/bash
# round 2: plain bench (phase times), launch list, one full capture of the three hot kernels
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 8000000 --min-seconds 0.01"
$B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || { tail -5 gpurun_out/r2_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bkt_pull|bkt_scatter|fm_forward_onehot16" --launch-skip 9 -c 3 -o gpurun_out/r2_hot $B > gpurun_out/r2_ncu_f.log 2>&1
ls -la gpurun_out/r2_hot.ncu-rep
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-partition > gpurun_out/r2_e2e.json 2> gpurun_out/r2_e2e.err; echo "e2e rc=$?"; tail -2 gpurun_out/r2_e2e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_e2e.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], "csr", d["e2e_csr"]["value"], "pack", d["e2e"]["host_packer"])
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/r2_gpu_tests.log
