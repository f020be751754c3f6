#!/bin/bash
mkdir -p gpurun_out
EXTRA="--rows 20000000" bash scripts/ab.sh c384mp c384 mp c384mpi12 2>&1 | tee gpurun_out/r2_ab4.txt
SFM_LIB=$PWD/sparkfm_b200/variants/libsparkfm_b200_c384mp.so timeout 600 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_ab4_tests.log 2>&1; echo "tests c384mp rc=$?"; tail -3 gpurun_out/r2_ab4_tests.log
