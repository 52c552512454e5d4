# end-of-round check on one B200: GPU tests, smoke, both bench arms, ncu launch list of the bench command
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/round_tests.log 2>&1; tail -3 gpurun_out/round_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/round_smoke.log 2>&1; tail -2 gpurun_out/round_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/round_bench.json 2> gpurun_out/round_bench.err; tail -c 600 gpurun_out/round_bench.json; tail -3 gpurun_out/round_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/round_bench_ref.json 2>&1; tail -c 400 gpurun_out/round_bench_ref.json
