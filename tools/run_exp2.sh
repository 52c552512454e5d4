timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_grad.py -x -q 2>&1 | tail -3
echo "== panel fused"; timeout 120 python bench.py --steps 5 --warmup 3 --only-value
NAGP_LIB=gpurun_exp/libnagp_exp9.so timeout 200 python tools/dbg_timeline_panel.py 1000 2>&1 | tail -24 | cut -c1-200
