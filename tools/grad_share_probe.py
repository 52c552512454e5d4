"""Time of nagp_logml_grad per chain by kind of kernel tree (which trees the per-entry reverse sweep costs most):
`python tools/grad_share_probe.py [K]` at the vignette shape, per-scenario hyperparameters."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nowcastautogp_b200 import kernels as kn, synthetic as syn
from nowcastautogp_b200.engine import Engine
n, k, P = 150, 1, 32
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = syn.make_workload(n, k, 0, K, P, seed=20261018 + 2)
eng = Engine(0)
progs = [bytes(w.ens.prog[w.ens.prog_off[p]:w.ens.prog_off[p + 1]]) for p in range(P)]
def klass(pr):
    if not (2 in pr or 8 in pr): return "one-table"
    if pr == bytes([2]): return "linear-only"
    return "multi-op"
groups = {}
for p, pr in enumerate(progs): groups.setdefault(klass(pr), []).append(p)
groups["all"] = list(range(P))
for p in groups["multi-op"]: groups[f"  #{p} {list(progs[p])}"] = [p]
for name, idx in groups.items():
    ens = kn.pack_ensemble([w.trees[p] for p in idx], np.asarray(w.noise)[idx])
    tk, nk = syn.perturbed_theta(ens, K, seed=77)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        eng.logml_grad(ens, w.t[:n + k], w.y1, y2=w.y2, g=w.g[:n + k], step=w.step, theta=tk, noise=nk)
        ts.append(time.perf_counter() - t0)
    ms = min(ts[1:]) * 1e3
    print(f"{name:40s} {len(idx):3d} x {K}: {ms:8.3f} ms  {ms * 1e3 / (K * len(idx)):7.3f} us/chain (host buffers)")
eng.close()
