tools/chol8_bench2_1; tools/chol8_bench2_0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_grad.py tests/test_gpu_edges.py -x -q 2>&1 | tail -4
for v in base layout; do echo "== $v"; NAGP_LIB=gpurun_exp/libnagp_$v.so timeout 120 python bench.py --steps 5 --warmup 3 --only-value; done
echo "== new"; timeout 120 python bench.py --steps 5 --warmup 3 --only-value
