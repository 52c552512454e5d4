timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 120 python bench.py --steps 5 --warmup 3 --only-value
timeout 120 python bench.py --steps 5 --warmup 3 --only-value
