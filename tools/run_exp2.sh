timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "failures_in_later or not_positive" 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['clocks'], d['gpu_launches'])"
