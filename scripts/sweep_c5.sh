# BASELINE config 5: Criteo-shaped config at per-GPU batch sizes 64k..1M.  usage: bash scripts/sweep_c5.sh NGPUS "64000 128000 ..."
N=${1:-1}
SIZES=${2:-"64000 128000 256000 512000 1000000"}
P='import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({"n_gpus": d["n_gpus"], "batch_per_gpu": d["config"]["batch_per_gpu"], "global_batch": d["config"]["global_batch"], "samples_per_s": d["value"], "ms_per_step": d["ms_per_step"], "step_roofline_frac": d["roofline"]["step"]["frac"], "partition_samples_per_s": d["partition_sampler"]["value"], "partition_ms_per_step": d["partition_sampler"]["ms_per_step"], "partition_roofline_frac": d["partition_sampler"]["roofline_step_frac"]}))'
for b in $SIZES; do
  if [ "$N" = "1" ]; then
    python bench.py --batch $b --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "$P"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --gpus $N --batch $b --steps 100 --warmup 5 --no-e2e 2>/dev/null | tail -1 | python -c "$P"
  fi
done
