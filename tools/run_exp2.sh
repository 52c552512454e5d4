for i in 1 2; do
python bench.py --steps 5 --warmup 3 --only-value
done
python -m pytest tests/test_gpu_grad.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
