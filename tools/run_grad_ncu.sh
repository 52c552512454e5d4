python tools/grad_probe.py 100 > gpurun_out/grad_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grad_tile -c 1 -s 3 -f -o gpurun_out/r01_grad_tile python tools/grad_probe.py 100 > gpurun_out/ncu_grad.log 2>&1
tail -3 gpurun_out/grad_plain.log; tail -2 gpurun_out/ncu_grad.log
