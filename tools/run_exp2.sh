for K in 1 9 50 1000; do timeout 15 python tools/hang_probe.py $K 2>&1 | tail -1; echo "K=$K"; done
timeout 40 python bench.py --steps 10 --warmup 3 --only-value
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_grad.py -x -q 2>&1 | tail -3
