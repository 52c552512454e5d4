timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/round_bench.json 2> gpurun_out/round_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/round_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'ach',d['roofline']['achieved'],'kms',d['roofline']['kernel_ms'])
print('hmc', d['hmc_gradient_microbench']['ms_per_step'], 'fast', d['fast_path']['ms_per_step'], d['clocks'])"
