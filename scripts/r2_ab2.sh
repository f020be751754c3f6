#!/bin/bash
mkdir -p gpurun_out
EXTRA="--rows 20000000" bash scripts/ab.sh ipt16 ipt16m t1024 mp 2>&1 | tee gpurun_out/r2_ab2.txt
timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py tests/test_gpu_als.py -x -q > gpurun_out/r2_ab2_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_ab2_tests.log
for v in ipt16; do SFM_LIB=$PWD/sparkfm_b200/variants/libsparkfm_b200_$v.so timeout 600 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_ab2_tests_$v.log 2>&1; echo "tests $v rc=$?"; tail -3 gpurun_out/r2_ab2_tests_$v.log; done
