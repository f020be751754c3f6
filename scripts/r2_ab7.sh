#!/bin/bash
mkdir -p gpurun_out
EXTRA="--rows 20000000" bash scripts/ab.sh pr2 2>&1 | tee gpurun_out/r2_ab7.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_ab7_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2_ab7_tests.log
