# gradient kernel: parity tests (both variants), then the timing probe at the vignette shape
timeout 900 python -m pytest tests/test_gpu_grad.py -x -q > gpurun_out/grad_tests.log 2>&1; tail -15 gpurun_out/grad_tests.log
NAGP_DEBUG=1 python tools/grad_probe.py 100 2>&1 | tail -12
