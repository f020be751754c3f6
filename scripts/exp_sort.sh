# A/B of the transposition sort: library (SFM_SORT=cub) against sfm_radix.cu (default)
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"])'
for q in ${MODES:-cub own}; do echo "== SFM_SORT=$q"; SFM_SORT=$q timeout 300 python bench.py --steps ${STEPS:-60} --warmup 5 --no-e2e --no-cpu-baseline --no-partition 2>gpurun_out/sort_err_$q.log | tee gpurun_out/sort_$q.json | python -c "$P"; tail -3 gpurun_out/sort_err_$q.log; done
