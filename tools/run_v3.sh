NAGP_VARIANT=3 NAGP_LIB=gpurun_exp/libnagp_t9.so python tools/dbg_timeline.py 1000 2>&1 | tail -30
NAGP_VARIANT=3 NAGP_LIB=gpurun_exp/libnagp_nq.so timeout 300 python bench.py --only-value --steps 5 --warmup 3 2>&1 | tail -1
