for v in u4 u4p7; do echo "== $v"; NAGP_LIB=gpurun_exp/libnagp_$v.so timeout 60 python bench.py --steps 10 --warmup 3 --only-value; done
echo "== current"; timeout 60 python bench.py --steps 10 --warmup 3 --only-value
