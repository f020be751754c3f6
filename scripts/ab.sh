#!/bin/bash
# ab.sh NAME... : short bench (phase times) of the default build and of each variant build
run() {
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-partition $EXTRA > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err || tail -3 gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
p=d["roofline"]["phase_ms"]
print(f"{sys.argv[1]:10s} step {d['ms_per_step']:.3f} ms  fwd {p['ms_forward']:.3f} sort {p['ms_sort']:.3f} reduce {p['ms_reduce']:.3f}  {d['value']/1e6:.0f} M/s loss {d['loss_first_last'][1]:.6f}")
PY
}
run default
for v in "$@"; do SFM_LIB=$PWD/sparkfm_b200/variants/libsparkfm_b200_$v.so run $v; done
