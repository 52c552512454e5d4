timeout 60 python bench.py --steps 10 --warmup 3 --only-value
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_grad.py -x -q 2>&1 | tail -3
