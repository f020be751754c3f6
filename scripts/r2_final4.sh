#!/bin/bash
# round-2 closing check #3: bit-sliced Bernoulli sampler -- GPU tests (device sampler vs host twins in every
# training test), bench points, plan-ahead A/B once more, full bench line, ncu launch list
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2f4_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2f4_tests.log
run() { # name batch env...
  env "${@:3}" timeout 150 python bench.py --batch $2 --steps 100 --warmup 5 --no-cpu-baseline --no-partition --no-e2e > gpurun_out/r2f4_$1.json 2> gpurun_out/r2f4_$1.err || tail -3 gpurun_out/r2f4_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2f4_{sys.argv[1]}.json"))
p=d["roofline"]["phase_ms"]
print(f"{sys.argv[1]:12s} step {d['ms_per_step']:.4f} ms  {d['value']/1e6:.1f} M/s frac {d['roofline']['frac']:.3f} fwd {p['ms_forward']:.3f} sort {p['ms_sort']:.3f} reduce {p['ms_reduce']:.3f} loss {d['loss_first_last'][1]:.6f}")
PY
}
run bs1m 1000000
run bs1m_b 1000000
run bs1m_plan 1000000 SFM_PLAN_AHEAD=1
run bs1m_noprio 1000000 SFM_STREAM_PRIO=0
run bs64k 64000
run bs256k 256000
timeout 240 python bench.py > gpurun_out/r2f4_bench_n1.json 2> gpurun_out/r2f4_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f4_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "part", d.get("partition_sampler",{}).get("value"), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --min-seconds 0.01"
SFM_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_final2.csv $B > /dev/null 2>&1; echo "ncu list rc=$?"
