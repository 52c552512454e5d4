"""Times nagp_logml_grad at the vignette shape (n = 150): P = 32 particles, and K x P per-scenario chains."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
n, k, P = 150, 1, 32
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
w = syn.make_workload(n, k, 0, K, P, seed=20261018 + 2)
th, nz = syn.perturbed_theta(w.ens, K, seed=77)
eng = Engine(0)
for name, fn in (("P=32 shared-theta", lambda: eng.logml_grad(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)),
                 (f"K={K} x P=32 per-scenario", lambda: eng.logml_grad(w.ens, w.t[:n + k], w.y1, y2=w.y2, g=w.g[:n + k], step=w.step, theta=th, noise=nz)),
                 ("logml only, same batch", lambda: eng.forecast_instances(w.ens, n, k, 0, w.t[:n + k], w.y1, w.y2, w.logw0, g=w.g[:n + k], step=w.step, theta=th, noise=nz, K=K, want_moments=False))):
    for i in range(3):
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
    print(f"{name}: {dt*1e3:.3f} ms")
eng.close()
