timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
NAGP_VARIANT=3 timeout 300 python bench.py --only-value --steps 5 --warmup 3 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('v2 step ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'fast', d['fast_path']['ms_per_step'])
print('logml512 ms', d['logml_microbench']['ms_per_step'], 'factor2048 ms', d['append_microbench']['factor_ms'], 'append ms', d['append_microbench']['append_ms'])"
