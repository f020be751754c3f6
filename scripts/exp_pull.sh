B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"])'
for lib in libsparkfm_b200.so libsparkfm_b200_u8.so; do for mb in 0 16 32; do echo "== $lib block_mb=$mb"; SFM_LIB=$PWD/sparkfm_b200/$lib SFM_PULL_BLOCK_MB=$mb $B 2>&1 | tail -1 | python -c "$P"; done; done
