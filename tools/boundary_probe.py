"""nagp_logml_batch across the size where the tile kernel hands over to the large-path kernel (q = 232 / 233):
`python tools/boundary_probe.py` (4000 instances per size, host buffers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
eng = Engine(0)
for n in (160, 168, 176, 200, 224, 232, 233, 240, 256):
    w = syn.make_workload(n, 0, 0, 1, 4000, seed=5, max_depth=4)
    ts = []
    for i in range(3):
        t0 = time.perf_counter()
        lm, info = eng.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
        ts.append(time.perf_counter() - t0)
    print(n, f"{min(ts)*1e3:.2f} ms", f"{min(ts)*1e6/4000/(n**3/3)*1e6:.3f} ns per MFLOP", eng.last_kernel)
