#!/bin/bash
# usage: tools/exp_build2.sh NAME "-DFLAG=..."   -> gpurun_exp/libnagp_NAME.so
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    --expt-relaxed-constexpr -Xcompiler -fPIC -shared $2 -o gpurun_exp/libnagp_$1.so nowcastautogp_b200/csrc/*.cu
