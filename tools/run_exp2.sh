echo "== prefetch"; NAGP_LIB=gpurun_exp/libnagp_pf.so timeout 40 python bench.py --steps 10 --warmup 3 --only-value
