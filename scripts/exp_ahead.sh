B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-partition"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["loss_first_last"])'
for a in 0 1; do echo "== SFM_SORT_AHEAD=$a"; SFM_SORT_AHEAD=$a $B 2>/dev/null | tail -1 | python -c "$P"; done
