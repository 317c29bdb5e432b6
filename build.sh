#!/bin/bash
# Builds libxgc.so (CUDA path, sm_100a) in-tree and the oracle.  Same commands as __graft_entry__.build().
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC -shared \
     -Iinclude -Ixgc_msm_b200/csrc -o xgc_msm_b200/libxgc.so xgc_msm_b200/csrc/xgc_lib.cu
make -C oracle
