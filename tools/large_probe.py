"""One nagp_logml_batch call at BASELINE configs[2] shape (1024 particles, n = 512) for ncu captures."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
B, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 512
w = syn.make_workload(n, 0, 0, 1, B, seed=20261018 + 3, max_depth=4, period=365.0)
eng = Engine(0)
for i in range(3):
    t0 = time.perf_counter()
    lm, info = eng.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    print("logml_batch ms", (time.perf_counter() - t0) * 1e3, "bad", int((info != 0).sum()))
eng.close()
