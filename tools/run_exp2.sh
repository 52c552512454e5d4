# scratch runner for one-off GPU experiments (overwritten per experiment)
timeout 40 python bench.py --steps 10 --warmup 3 --only-value
