python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -3 gpurun_out/bench_now.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_now.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
for k in ('hmc_gradient_microbench','summary_microbench','logml_microbench','append_microbench'):
    print(k, json.dumps(d.get(k))[:700])
PY
