#!/bin/bash
mkdir -p gpurun_out
EXTRA="--rows 20000000" bash scripts/ab.sh sr2 sr2i16 sr1i16 sr0i16 2>&1 | tee gpurun_out/r2_ab6.txt
SFM_LIB=$PWD/sparkfm_b200/variants/libsparkfm_b200_sr2i16.so timeout 600 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_ab6_tests.log 2>&1; echo "tests sr2i16 rc=$?"; tail -3 gpurun_out/r2_ab6_tests.log
