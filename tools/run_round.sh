# End-of-round evidence on one B200 (run under gpurun): GPU tests, smoke, both bench arms, then — only after the plain
# runs exited 0 — the ncu launch list of the bench command and one full capture per heavy kernel. Outputs go to gpurun_out/;
# tools/ncu_summary.py turns the .ncu-rep files into the CSVs committed under profiles/.
R=${1:-r02}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/round_tests.log 2>&1; tail -3 gpurun_out/round_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/round_smoke.log 2>&1; tail -2 gpurun_out/round_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/round_bench.json 2> gpurun_out/round_bench.err || exit 1
tail -c 400 gpurun_out/round_bench.json; tail -3 gpurun_out/round_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/round_bench_ref.json 2>&1; tail -c 300 gpurun_out/round_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/round_ncu.log 2>&1; tail -2 gpurun_out/round_ncu.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_v2 -c 1 -s 4 -f -o gpurun_out/${R}_v2_fused python bench.py --steps 2 --warmup 3 --only-value --variant 2 > gpurun_out/ncu_v2.log 2>&1; tail -1 gpurun_out/ncu_v2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_v3 -c 1 -s 4 -f -o gpurun_out/${R}_v3_slot python bench.py --steps 2 --warmup 3 --only-value --variant 3 > gpurun_out/ncu_v3.log 2>&1; tail -1 gpurun_out/ncu_v3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_append -c 1 -s 2 -f -o gpurun_out/${R}_rank_append python tools/append_probe.py > gpurun_out/ncu_append.log 2>&1; tail -1 gpurun_out/ncu_append.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chol_large -c 1 -s 2 -f -o gpurun_out/${R}_large_chol python tools/large_probe.py > gpurun_out/ncu_large.log 2>&1; tail -1 gpurun_out/ncu_large.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grad_tile -c 1 -s 3 -f -o gpurun_out/${R}_grad_tile python tools/grad_probe.py 100 > gpurun_out/ncu_grad.log 2>&1; tail -1 gpurun_out/ncu_grad.log
