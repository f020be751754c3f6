#!/bin/bash
mkdir -p gpurun_out
run() { # name env...
  env "${@:2}" python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-partition --rows 20000000 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err || tail -3 gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
p=d["roofline"]["phase_ms"]
print(f"{sys.argv[1]:10s} step {d['ms_per_step']:.4f} ms  fwd {p['ms_forward']:.3f} sort {p['ms_sort']:.3f} reduce {p['ms_reduce']:.3f}  {d['value']/1e6:.0f} M/s loss {d['loss_first_last'][1]:.6f}")
PY
}
run base SFM_FWD_COUNT=0
run fwdcount SFM_FWD_COUNT=1
SFM_FWD_COUNT=1 timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_ab9_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_ab9_tests.log
