set -x
for e in 1 2 3 4; do NAGP_LIB=gpurun_exp/libnagp_exp$e.so python bench.py --only-value --skip-sanity --steps 5 --warmup 3 > gpurun_out/exp$e.log 2>&1; done
python bench.py --only-value --steps 5 --warmup 3 > gpurun_out/exp0.log 2>&1
NAGP_LIB=gpurun_exp/libnagp_exp9.so python tools/dbg_timeline.py 1000 > gpurun_out/timeline.log 2>&1
tools/chol8_bench > gpurun_out/chol8.log 2>&1
cat gpurun_out/exp*.log gpurun_out/timeline.log gpurun_out/chol8.log
