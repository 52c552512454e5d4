timeout 800 python -m pytest tests/test_gpu_large.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps(d['logml_microbench'],indent=None)[:600]); print(json.dumps(d['append_microbench'])[300:])"
