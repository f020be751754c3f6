#!/bin/bash
# round-2 closing check on one GPU: GPU tests (incl. the full-size config 3 run), A/B of the own sampler
# kernels against the previous library-select build and of the stream-priority knob, smoke
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2f2_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2f2_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f2_smoke.log
run() { # name env...
  env "${@:2}" timeout 150 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-partition --e2e-steps 20 > gpurun_out/r2f2_$1.json 2> gpurun_out/r2f2_$1.err || tail -3 gpurun_out/r2f2_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/r2f2_{sys.argv[1]}.json"))
p=d["roofline"]["phase_ms"]
print(f"{sys.argv[1]:10s} step {d['ms_per_step']:.4f} ms  {d['value']/1e6:.1f} M/s frac {d['roofline']['frac']:.3f} fwd {p['ms_forward']:.3f} sort {p['ms_sort']:.3f} reduce {p['ms_reduce']:.3f} e2e {d['e2e']['value']/1e6:.1f} loss {d['loss_first_last'][1]:.6f}")
PY
}
run own
run prio SFM_STREAM_PRIO=1
run cubsel SFM_LIB=$PWD/sparkfm_b200/variants/libsparkfm_b200_cubsel.so
run prio2 SFM_STREAM_PRIO=1
run own2
