#!/bin/bash
# round-2 closing check #4 (final defaults): GPU tests incl. the shard-offset sampler tests, the full bench line
# and the reference arm, one ncu --set full capture of a step's kernels (sampler kernels included)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2f5_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2f5_tests.log
timeout 200 python bench.py > gpurun_out/r2f5_bench_n1.json 2> gpurun_out/r2f5_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f5_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "part", d.get("partition_sampler",{}).get("value"), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "launches", d.get("gpu_launches"))
PY
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --min-seconds 0.01"
SFM_GRAPH=0 timeout 240 ncu --set full --clock-control none --import-source on -k regex:"bkt_pull|bkt_scatter|fm_forward_onehot16_kernel<1>|bkt_count|bkt_offsets|select_mask|select_write|scan_tiles" --launch-skip 18 -c 9 -o gpurun_out/r2_final2_hot $B > gpurun_out/r2f5_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_final2_hot.ncu-rep
