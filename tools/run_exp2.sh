timeout 40 python bench.py --steps 10 --warmup 3 --only-value
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q 2>&1 | tail -2
