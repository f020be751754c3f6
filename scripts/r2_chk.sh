#!/bin/bash
mkdir -p gpurun_out
run() { # name batch env...
  env "${@:3}" python bench.py --batch $2 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-partition --rows 20000000 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err || tail -3 gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
p=d["roofline"]["phase_ms"]
print(f"{sys.argv[1]:16s} step {d['ms_per_step']:.4f} ms  {d['value']/1e6:.0f} M/s  frac {d['roofline']['frac']:.3f} fwd {p['ms_forward']:.3f} sort {p['ms_sort']:.3f} reduce {p['ms_reduce']:.3f} loss {d['loss_first_last'][1]:.6f}")
PY
}
run b64k 64000
run b125k 125000
run b256k 256000
run b1m 1000000
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_chk_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2_chk_tests.log
