#!/bin/bash
mkdir -p gpurun_out
export SFM_P2P_TIMEOUT_S=30
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "n_gpus_match or bad_index" > gpurun_out/r2_n2b_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r2_n2b_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582"
run() { # name batch env...
  env "${@:3}" timeout 300 $TR bench.py --gpus 2 --batch $2 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-partition --no-parity --weak-batch 0 --rows 20000000 > gpurun_out/n2_$1.json 2> gpurun_out/n2_$1.err; 
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/n2_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:14s} global {d['config']['global_batch']} step {d['ms_per_step']:.4f} ms {d['value']/1e6:.0f} M/s", d["roofline"]["phase_ms"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run g128k 128000 SFM_GRAPH=1
run g128k_nograph 128000 SFM_GRAPH=0
run g128k_dense 128000 SFM_P2P_SPARSE=0
run g128k_nccl 128000 SFM_P2P=0
run g1m 1000000 SFM_GRAPH=1
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-partition > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "full n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print("n2 full: value", d["value"], "step", d["ms_per_step"], "e2e", d["e2e"]["value"], "csr", d["e2e_csr"]["value"], "weak", d["weak_scaling"], "parity", d["parity_n"]["ok"])
PY
