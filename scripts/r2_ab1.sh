#!/bin/bash
# A/B of the pull variants (phase times, 20 steps each) + the sort/parity tests on the default build
mkdir -p gpurun_out
EXTRA="--rows 20000000" bash scripts/ab.sh match u8 2>&1 | tee gpurun_out/r2_ab1.txt
timeout 900 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_ab1_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_ab1_tests.log
