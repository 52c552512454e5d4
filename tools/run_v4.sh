NAGP_DEBUG=1 NAGP_VARIANT=4 timeout 300 python bench.py --only-value --steps 3 --warmup 3 2>&1 | sort | uniq -c | tail -3
NAGP_DEBUG=1 NAGP_VARIANT=2 timeout 300 python bench.py --only-value --steps 3 --warmup 3 2>&1 | sort | uniq -c | tail -3
