# A/B two builds of the library: bash scripts/exp_ab.sh libA.so libB.so
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms"], "part", d["partition_sampler"]["ms_per_step"])'
for lib in "$@"; do echo "== $lib"; SFM_LIB=$PWD/sparkfm_b200/$lib $B 2>/dev/null | tail -1 | python -c "$P"; done
