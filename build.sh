#!/bin/bash
# Builds libxgc.so (CUDA path, sm_100a) in-tree and the oracle.  Same commands as __graft_entry__.build().
# Two translation units: xgc_lib.cu (strict arithmetic contract: --fmad=false) and xgc_fast.cu (fast native-draw path: FMA on).
set -e
cd "$(dirname "$0")"
mkdir -p build
ARCH="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Ixgc_msm_b200/csrc"
nvcc $ARCH --fmad=false -c xgc_msm_b200/csrc/xgc_lib.cu -o build/xgc_lib.o
nvcc $ARCH -c xgc_msm_b200/csrc/xgc_fast.cu -o build/xgc_fast.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o xgc_msm_b200/libxgc.so build/xgc_lib.o build/xgc_fast.o
make -C oracle
