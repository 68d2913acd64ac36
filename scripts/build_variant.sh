#!/bin/bash
# Developer helper: build an experiment variant of libljgpu.so with extra -D flags, selected at run time with
# LJGPU_LIBRARY=<path>.   usage: scripts/build_variant.sh <name> [-DLJ_TILE_REC32=0 ...]
set -e
cd "$(dirname "$0")/../lennard-jones-live-simulator_b200"
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
    -Xcompiler -fPIC,-ffp-contract=off -ccbin g++ "$@" -c csrc/lj_sim.cu -o build/variants/lj_sim_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libljgpu_$name.so \
    build/variants/lj_sim_$name.o build/lj_sort.o build/lj_host.o -ccbin g++
echo build/variants/libljgpu_$name.so
