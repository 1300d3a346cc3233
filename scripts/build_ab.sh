#!/bin/bash
# build A/B variants of libsvgt.so with extra -D flags: bash scripts/build_ab.sh name "-DFOO=1 ..."
set -e
cd "$(dirname "$0")/../svtyper_b200/csrc"
mkdir -p ../ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared $2 -o ../ab/libsvgt_$1.so svgt_kernels.cu svgt_compact.cu svgt_api.cu
echo built ../ab/libsvgt_$1.so
