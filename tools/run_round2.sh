# end-of-round check on one B200: GPU tests, smoke, both bench arms, ncu launch list of the bench command
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/round_tests.log 2>&1; tail -3 gpurun_out/round_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/round_smoke.log 2>&1; tail -2 gpurun_out/round_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/round_bench.json 2> gpurun_out/round_bench.err; tail -c 300 gpurun_out/round_bench.json; tail -3 gpurun_out/round_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/round_bench_ref.json 2>&1; tail -c 300 gpurun_out/round_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/round_ncu.log 2>&1; tail -2 gpurun_out/round_ncu.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_v2 -c 1 -s 4 -f -o gpurun_out/r01_v2i_final python bench.py --steps 2 --warmup 3 --only-value > gpurun_out/ncu_v2i.log 2>&1; tail -2 gpurun_out/ncu_v2i.log
