# usage: bash scripts/exp_slices.sh NGPUS
N=${1:-2}
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "part", d["partition_sampler"]["value"], d["partition_sampler"]["ms_per_step"], d["partition_sampler"]["roofline_step_frac"])'
for q in ${SLICES:-1 2 4}; do echo "== SFM_AR_SLICES=$q"; SFM_AR_SLICES=$q python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$q bench.py --gpus $N --steps 50 --warmup 5 --no-e2e 2>/dev/null | python -c "$P"; done
