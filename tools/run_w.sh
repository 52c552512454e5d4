for w in 4 6; do
NAGP_LIB=gpurun_exp/libnagp_w$w.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tile" 2>&1 | tail -1
NAGP_LIB=gpurun_exp/libnagp_w$w.so timeout 300 python bench.py --only-value --steps 5 --warmup 3 2>&1 | tail -1
done
timeout 300 python bench.py --only-value --steps 5 --warmup 3 2>&1 | tail -1
