#!/bin/bash
# 2-GPU check of the sparse peer-memory exchange: tests, then the strong-scaling bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "n_gpus or bad_index" > gpurun_out/r2_multi2_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2_multi2_tests.log
for sp in 1 0; do
SFM_P2P_SPARSE=$sp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-partition --weak-batch 0 > gpurun_out/r2_n2_sparse$sp.json 2> gpurun_out/r2_n2_sparse$sp.err; echo "bench sparse=$sp rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_n2_sparse$sp.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], d.get("parity_n"))
except Exception as e:
    print("parse failed", e)
PY
done
tail -3 gpurun_out/r2_n2_sparse1.err
