#!/bin/bash
mkdir -p gpurun_out
export SFM_P2P_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29584"
timeout 600 $TR bench.py --gpus 4 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"
for b in 64000 256000; do
timeout 300 $TR bench.py --gpus 4 --batch $b --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-partition --no-parity --weak-batch 0 > gpurun_out/r2_c5_n4_$b.json 2> gpurun_out/r2_c5_n4_$b.err; echo "c5 $b rc=$?"
done
python - <<'PY'
import json
for f in ("r2_bench_n4","r2_c5_n4_64000","r2_c5_n4_256000"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["config"]["global_batch"], "step", round(d["ms_per_step"],4), "value", round(d["value"]/1e6,1), "M/s", "e2e", (d.get("e2e") or {}).get("value"), "weak", (d.get("weak_scaling") or {}).get("value"), "part", (d.get("partition_sampler") or {}).get("value"), "parity", (d.get("parity_n") or {}).get("ok"))
    except Exception as e:
        print(f, "parse failed", e)
PY
tail -3 gpurun_out/r2_bench_n4.err
