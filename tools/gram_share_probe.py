"""Kernel time per instance for the single-leaf and the multi-node particles of the bench ensemble separately
(how much of the step is the Gram pre-pass of the larger trees)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from nowcastautogp_b200 import kernels as kn, synthetic as syn
from nowcastautogp_b200.engine import Engine
w, th, nz, z, u = bench.make_inputs(0)
c = bench.CFG
eng = Engine(0)
K = c["K"]
plen = [int(w.ens.prog_off[p + 1] - w.ens.prog_off[p]) for p in range(w.ens.size)]
for name, idx in (("single-leaf", [p for p in range(w.ens.size) if plen[p] == 1]),
                  ("multi-node", [p for p in range(w.ens.size) if plen[p] > 1]),
                  ("all", list(range(w.ens.size)))):
    ens = kn.pack_ensemble([w.trees[p] for p in idx], np.asarray(w.noise)[idx])
    tk, nk = syn.perturbed_theta(ens, K, seed=77)
    lw0 = np.ascontiguousarray(w.logw0[idx])
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        eng.forecast_instances(ens, c["n"], c["k"], c["h"], w.t, w.y1, w.y2, lw0, w.ya, w.yb, g=w.g, step=w.step,
                               theta=tk, noise=nk)
        ts.append(time.perf_counter() - t0)
    B = K * len(idx)
    ms = min(ts[1:]) * 1e3
    print(f"{name:12s} {len(idx):3d} particles x {K} scenarios: {ms:7.3f} ms (host buffers), "
          f"{ms * 1e-3 * 296 / B * 1.965e9 / 1e3:7.1f} k cycles per instance per CTA")
