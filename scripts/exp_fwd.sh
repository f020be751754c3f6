B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], d["predict"]["value"])'
for lib in libsparkfm_b200.so libsparkfm_b200_lb5.so libsparkfm_b200_lb6.so; do echo "== $lib"; SFM_LIB=$PWD/sparkfm_b200/$lib $B 2>&1 | tail -1 | python -c "$P"; done
