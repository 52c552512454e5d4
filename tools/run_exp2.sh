timeout 40 python bench.py --steps 10 --warmup 3 --only-value
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
