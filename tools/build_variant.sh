#!/bin/bash
# tools/build_variant.sh NAME [-D...]: builds build/libyavo_NAME.so with extra nvcc flags (A/B tuning runs: YAVO_LIB_PATH=build/libyavo_NAME.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -shared -Xcompiler -fPIC "$@" \
    -o build/libyavo_$name.so ya_vo_b200/csrc/yavo_capi.cu
echo built build/libyavo_$name.so
