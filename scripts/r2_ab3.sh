#!/bin/bash
mkdir -p gpurun_out
EXTRA="--rows 20000000" bash scripts/ab.sh p256 p256i16 it4 it4mp p256mp 2>&1 | tee gpurun_out/r2_ab3.txt
