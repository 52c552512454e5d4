for K in 1 9 50 1000; do timeout 15 python tools/hang_probe.py $K 2>&1 | tail -1; done
timeout 40 python bench.py --steps 10 --warmup 3 --only-value
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
