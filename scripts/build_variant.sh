#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. -DFLAG2=.." : A/B build of libsparkfm_b200 with extra nvcc defines
# -> sparkfm_b200/variants/libsparkfm_b200_NAME.so (select with SFM_LIB=...)
set -e
cd "$(dirname "$0")/../sparkfm_b200/csrc"
NAME=$1; DEFS=$2
mkdir -p build_$NAME ../variants
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -ccbin /usr/bin/g++ $DEFS"
for f in sfm_kernels sfm_sort sfm_radix sfm_scan sfm_bucket sfm_api sfm_shard sfm_p2p sfm_als; do
  if [ build/$f.o -nt $f.cu ] && [ "$f" != "sfm_bucket" ] && [ "$f" != "sfm_kernels" ] && [ -z "$ALL" ]; then cp build/$f.o build_$NAME/$f.o; else $NVCC $FLAGS -c $f.cu -o build_$NAME/$f.o & fi
done
for f in sfm_host sfm_nccl; do cp build/$f.o build_$NAME/$f.o; done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libsparkfm_b200_$NAME.so build_$NAME/*.o -ldl -lpthread
echo built ../variants/libsparkfm_b200_$NAME.so
