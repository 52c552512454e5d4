#!/bin/bash
# Timing-attribution builds of libnagp (deliberately wrong kernels; never shipped): see NAGP_EXP in
# csrc/nagp_fused_v2.cu. Usage: tools/exp_build.sh 1 2 3 4
cd "$(dirname "$0")/.."
for e in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
    --expt-relaxed-constexpr -Xcompiler -fPIC -shared -DNAGP_EXP=$e \
    -o gpurun_exp/libnagp_exp$e.so nowcastautogp_b200/csrc/*.cu &
done
wait
