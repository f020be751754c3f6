#!/bin/bash
# round-2 closing check #5: 2 GPUs, short strong-scaling line with parity_n (bit-sliced sampler on shards whose
# offsets are not multiples of 64: 8,000,001 rows)
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e --no-partition --no-cpu-baseline --rows 8000001 --weak-batch 0 > gpurun_out/r2f6_n2.json 2> gpurun_out/r2f6_n2.err; echo "n2 rc=$?"
tail -3 gpurun_out/r2f6_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f6_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "n", d["n_gpus"], "parity", json.dumps(d.get("parity_n"))[:900])
PY
