#!/bin/bash
mkdir -p gpurun_out
export SFM_P2P_TIMEOUT_S=30
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_n2c_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r2_n2c_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29586"
for b in 1000000 128000; do
SFM_P2P_TRACE=1 timeout 300 $TR bench.py --gpus 2 --batch $b --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-partition --weak-batch 0 > gpurun_out/r2_n2c_$b.json 2> gpurun_out/r2_n2c_$b.err; echo "bench $b rc=$?"
python - $b <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/r2_n2c_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], d["config"]["global_batch"], "step", round(d["ms_per_step"],4), "value", round(d["value"]/1e6,1), "M/s parity", (d.get("parity_n") or {}).get("ok"))
except Exception as e:
    print("parse failed", e)
PY
grep "p2p trace" gpurun_out/r2_n2c_$b.err | head -2
done
