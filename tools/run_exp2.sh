NAGP_LIB=gpurun_exp/libnagp_p7.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
echo "== panel 7"; NAGP_LIB=gpurun_exp/libnagp_p7.so timeout 120 python bench.py --steps 5 --warmup 3 --only-value
