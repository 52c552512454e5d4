# runs bench --only-value against every experiment build in gpurun_exp/
for so in gpurun_exp/libnagp_*.so; do
  echo -n "$so "; NAGP_LIB=$PWD/$so timeout -k 10 60 python bench.py --steps 10 --warmup 3 --only-value 2>&1 | tail -1
done
