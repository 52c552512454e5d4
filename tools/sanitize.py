"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
eng = Engine(0)
rng = np.random.default_rng(0)
for variant in (2, 1):
    eng.set_variant(variant)
    for (n, k, h, P, K, D) in ((30, 2, 5, 3, 2, 3), (61, 1, 4, 2, 2, 2)):
        w = syn.make_workload(n, k, h, K, P, seed=n)
        zeta, u = rng.standard_normal((K, D, h)), rng.uniform(size=(K, D))
        eng.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
        eng.logml_batch(w.ens, w.t[:n], w.y1)
        r = eng.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
        eng.draw(r["logw"], r["mu"], r["L"], zeta, u=u, u_res=rng.uniform(size=(K, P)), ess_thr=1.0)
        eng.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, zeta, w.ya, w.yb, g=w.g, step=w.step, u=u)
eng.set_variant(0)
n = 250
w = syn.make_workload(n + 20, 0, 0, 1, 2, seed=3)
eng.logml_batch(w.ens, w.t[:n], w.y1[:n], g=w.g[:n], step=w.step)
f = eng.factor_store_large(w.ens, w.t[:n], w.y1[:n], capacity=n + 20, g=w.g[:n], step=w.step)
eng.factor_append(f, w.t[n:n + 3], w.y1[n:n + 3], g_new=w.g[n:n + 3])
eng.factor_append(f, w.t[n + 3:n + 20], w.y1[n + 3:n + 20], g_new=w.g[n + 3:n + 20])
f.free()
w = syn.make_workload(236, 1, 3, 2, 2, seed=4)
eng.forecast_with_nowcasts(w.ens, 236, 1, 3, w.t, w.y1, w.y2, w.logw0, rng.standard_normal((2, 2, 3)), w.ya, w.yb,
                           g=w.g, step=w.step, u=rng.uniform(size=(2, 2)))
eng.close()
print("sanitize pass done")
