# usage: bash scripts/exp_p2p.sh NGPUS   -- A/B of the gradient exchange: NCCL all-reduce + update
# kernel (SFM_P2P=0) against the fused sum + update kernel over NVLink peer memory (SFM_P2P=1)
N=${1:-2}
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "part", d["partition_sampler"]["value"], d["partition_sampler"]["ms_per_step"], d["partition_sampler"]["roofline_step_frac"], d.get("comm_mode"))'
for q in ${MODES:-0 1}; do echo "== SFM_P2P=$q"; SFM_P2P=$q SFM_P2P_TIMEOUT_S=20 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$q bench.py --gpus $N --steps ${STEPS:-100} --warmup 5 --no-e2e --no-cpu-baseline 2>gpurun_out/p2p_err_$q.log | tee gpurun_out/p2p_n${N}_$q.json | python -c "$P"; tail -3 gpurun_out/p2p_err_$q.log; done
