"""Panel-warp timeline of the v4 kernel from an NAGP_EXP=10 build (clock64 stamps of CTA 0, instance 0)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from nowcastautogp_b200.engine import Engine
from nowcastautogp_b200 import _lib
w, th, nz, z, u = bench.make_inputs(0)
c = bench.CFG
eng = Engine(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for rep in range(2):
    eng.forecast_instances(w.ens, c["n"], c["k"], c["h"], w.t, w.y1, w.y2[:K], w.logw0, w.ya, w.yb, g=w.g,
                           step=w.step, theta=th[:K], noise=nz[:K])
lib = _lib.load()
buf = (C.c_longlong * 8192)()
lib.nagp_debug_read(buf, 8192)
d = np.array(buf[:], dtype=np.int64)
g = d[8000:8006]
print("phases: setup %d tables %d gram %d factor %d logml %d" % tuple(np.diff(g[:6])))
t = d[:20 * 64].reshape(20, 8, 8)
p = t[:, 0, :4]
print("J | wait-pre  trsm+upd  chol8 | period")
for J in range(20):
    per = p[J, 3] - p[J - 1, 3] if J else 0
    print(f"{J:2d} | {p[J,1]-p[J,0]:6d} {p[J,2]-p[J,1]:6d} {p[J,3]-p[J,2]:6d} | {per:6d}")
