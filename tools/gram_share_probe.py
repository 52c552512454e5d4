"""Kernel time per instance by kind of kernel tree in the bench ensemble (how much of the step is Gram production):
`python tools/gram_share_probe.py [variant]` — stationary-only trees compile to one lag table, Linear-only trees to one
non-stationary leaf, the rest go through the multi-op interpreter."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from nowcastautogp_b200 import kernels as kn, synthetic as syn
from nowcastautogp_b200.engine import Engine
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
w, th, nz, z, u = bench.make_inputs(0)
c = bench.CFG
eng = Engine(0)
eng.set_variant(variant)
K = c["K"]
progs = [bytes(w.ens.prog[w.ens.prog_off[p]:w.ens.prog_off[p + 1]]) for p in range(w.ens.size)]
stat = [p for p, pr in enumerate(progs) if not (2 in pr or 8 in pr)]
lin = [p for p, pr in enumerate(progs) if pr == bytes([2])]
rest = [p for p in range(len(progs)) if p not in stat and p not in lin]
slots = 444 if variant in (0, 3) else 296
for name, idx in (("one-table", stat), ("linear-only", lin), ("multi-op", rest), ("all", list(range(w.ens.size)))):
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
    print(f"variant {variant} kernel {eng.last_kernel} {name:12s} {len(idx):3d} particles x {K} scenarios: {ms:7.3f} ms (host buffers), "
          f"{ms * 1e-3 * slots / B * 1.965e9 / 1e3:7.1f} k cycles per instance per matrix slot")
