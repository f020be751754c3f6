#!/bin/bash
# round-2 closing check #2: GPU tests with the new defaults (stream priorities, plan-ahead), A/B of plan-ahead,
# small-batch points, full bench line
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2f3_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2f3_tests.log
run() { # name batch env...
  env "${@:3}" timeout 150 python bench.py --batch $2 --steps 100 --warmup 5 --no-cpu-baseline --no-partition --no-e2e > gpurun_out/r2f3_$1.json 2> gpurun_out/r2f3_$1.err || tail -3 gpurun_out/r2f3_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2f3_{sys.argv[1]}.json"))
p=d["roofline"]["phase_ms"]
print(f"{sys.argv[1]:12s} step {d['ms_per_step']:.4f} ms  {d['value']/1e6:.1f} M/s frac {d['roofline']['frac']:.3f} fwd {p['ms_forward']:.3f} sort {p['ms_sort']:.3f} reduce {p['ms_reduce']:.3f} loss {d['loss_first_last'][1]:.6f}")
PY
}
run plan 1000000
run noplan 1000000 SFM_PLAN_AHEAD=0
run plan2 1000000
run noplan2 1000000 SFM_PLAN_AHEAD=0
run plan64k 64000
run noplan64k 64000 SFM_PLAN_AHEAD=0
run noprio64k 64000 SFM_STREAM_PRIO=0
run plan256k 256000
timeout 240 python bench.py > gpurun_out/r2f3_bench_n1.json 2> gpurun_out/r2f3_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f3_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "part", d.get("partition_sampler",{}).get("value"), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
