#!/bin/bash
# usage: tools/exp_build.sh NAME "-DFLAG=... -DFLAG2=..."   -> gpurun_exp/libnagp_NAME.so (experiment builds; run with NAGP_LIB=...)
cd "$(dirname "$0")/.."
mkdir -p gpurun_exp
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    --expt-relaxed-constexpr -Xcompiler -fPIC -shared $2 -o gpurun_exp/libnagp_$1.so nowcastautogp_b200/csrc/*.cu -ldl
