"""Times nagp_factor_store_large / nagp_factor_append at BASELINE configs[4] shape (for ncu launch lists)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
P, n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 2048, int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = syn.make_workload(n + 4 * k, 0, 0, 1, P, seed=20261023, max_depth=4, period=365.0)
eng = Engine(0)
t0 = time.perf_counter()
f = eng.factor_store_large(w.ens, w.t[:n], w.y1[:n], capacity=n + 4 * k, g=w.g[:n], step=w.step, check=False)
t1 = time.perf_counter()
print("store_large ms", (t1 - t0) * 1e3, "bad", int((f.info != 0).sum()))
cur = n
for i in range(4):
    t0 = time.perf_counter()
    eng.factor_append(f, w.t[cur:cur + k], w.y1[cur:cur + k], g_new=w.g[cur:cur + k], check=False)
    t1 = time.perf_counter()
    cur += k
    print("append ms", (t1 - t0) * 1e3)
f.free(); eng.close()
