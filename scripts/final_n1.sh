# round-end evidence on one GPU: tests, smoke, the default bench line, the ncu launch list
set -x
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-partition --rows 8000000 > gpurun_out/ncu_final.log 2>&1
tail -3 gpurun_out/final_gpu_tests.log gpurun_out/final_smoke.log; head -c 600 gpurun_out/bench_n1.json
