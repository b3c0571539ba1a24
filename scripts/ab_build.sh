#!/bin/bash
# build one variant of the library for an A/B run on the GPU box: scripts/ab_build.sh NAME [-DFLAG ...]
N=$1; shift
mkdir -p build/ab
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -shared \
    "$@" flacarray_b200/csrc/fa_api.cu -o build/ab/$N.so
