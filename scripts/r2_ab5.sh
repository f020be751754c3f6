#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  env "${@:2}" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 20000000 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err || tail -3 gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
p=d["roofline"]["phase_ms"]
print(f"{sys.argv[1]:10s} step {d['ms_per_step']:.3f} ms  fwd {p['ms_forward']:.3f} sort {p['ms_sort']:.3f} reduce {p['ms_reduce']:.3f}  {d['value']/1e6:.0f} M/s loss {d['loss_first_last'][1]:.6f}")
PY
}
run tile SFM_SCATTER=tile
run warp SFM_SCATTER=warp
SFM_SCATTER=warp timeout 600 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_ab5_tests.log 2>&1; echo "tests warp rc=$?"; tail -3 gpurun_out/r2_ab5_tests.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum --clock-control none -k regex:"bkt_w" -c 6 --csv --log-file gpurun_out/r2_wscatter.csv env SFM_SCATTER=warp python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 8000000 --min-seconds 0.01 > /dev/null 2>&1
tail -8 gpurun_out/r2_wscatter.csv
