python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/round_ncu.log 2>&1; tail -2 gpurun_out/round_ncu.log | cut -c1-200
python tools/grad_probe.py 100 > gpurun_out/grad_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grad_tile -c 1 -s 3 -f -o gpurun_out/r01_grad_tile python tools/grad_probe.py 100 > gpurun_out/ncu_grad.log 2>&1
tail -3 gpurun_out/grad_plain.log; tail -2 gpurun_out/ncu_grad.log
ncu --set full --clock-control none --import-source on -k regex:fused_v2 -c 1 -s 4 -f -o gpurun_out/r01_v2e_fused python bench.py --steps 2 --warmup 3 --only-value > gpurun_out/ncu_v2e.log 2>&1; tail -2 gpurun_out/ncu_v2e.log
