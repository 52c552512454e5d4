set -x
python tools/large_probe.py > gpurun_out/large_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chol_large -c 1 -s 2 -o gpurun_out/r01_large_chol python tools/large_probe.py > gpurun_out/ncu_large.log 2>&1
python tools/append_probe.py > gpurun_out/append_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rank_append -c 1 -s 2 -o gpurun_out/r01_rank_append python tools/append_probe.py > gpurun_out/ncu_append.log 2>&1
cat gpurun_out/large_plain.log gpurun_out/append_plain.log; tail -3 gpurun_out/ncu_large.log gpurun_out/ncu_append.log
