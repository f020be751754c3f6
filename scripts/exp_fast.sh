B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], "part", d["partition_sampler"]["ms_per_step"], "pred", d["predict"]["value"])'
echo "== generic"; SFM_NO_FASTPATH=1 $B 2>/dev/null | tail -1 | python -c "$P"
echo "== fast path"; $B 2>/dev/null | tail -1 | python -c "$P"
