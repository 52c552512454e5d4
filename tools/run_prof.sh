python tools/large_probe.py > gpurun_out/large_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chol_large -c 1 -s 2 -f -o gpurun_out/r01_large_chol2 python tools/large_probe.py > gpurun_out/ncu_large.log 2>&1
cat gpurun_out/large_plain.log
