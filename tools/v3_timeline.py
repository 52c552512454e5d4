"""Per-step timeline of one instance of the slot kernel (needs a build with -DNAGP_V3_TRACE=n: the n-th instance of
block 0 / slot 0 is stamped): `NAGP_LIB=gpurun_exp/libnagp_trace.so python tools/v3_timeline.py`."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
w, th, nz, z, u = bench.make_inputs(0)
c = bench.CFG
eng = Engine(0)
eng.set_variant(3)
for _ in range(2):
    eng.forecast_instances(w.ens, c["n"], c["k"], c["h"], w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step, theta=th, noise=nz)
N = 32 * 5 * 8 + 16
buf = (C.c_longlong * N)()
eng._lib.nagp_debug_read.argtypes = [C.c_void_p, C.c_int]
assert eng._lib.nagp_debug_read(buf, N) == 0
t = np.array(buf[:32 * 5 * 8]).reshape(32, 5, 8)
g = np.array(buf[32 * 5 * 8:])
t0 = g[0]
print(f"particle {g[8]}: prologue {g[1]-g[0]}, roles {g[2]-g[1]} cycles")
nt = 20
print("J | chain: start  dur | gram: start dur | rows r0,r1,r2: start  a  waitB  c  waitD  d   (cycles; start relative to instance start)")
for J in range(nt):
    ch, gr = t[J, 0], t[J, 1]
    line = f"{J:2d} | {ch[0]-t0:7d} {ch[1]-ch[0]:5d} | {gr[0]-t0:7d} {gr[1]-gr[0]:5d} |"
    for r in range(3):
        x = t[J, 2 + r]
        if J + 1 < nt:
            line += f" [{x[0]-t0:7d} a{x[1]-x[0]:5d} wB{x[2]-x[1]:5d} c{x[3]-x[2]:5d} wD{x[4]-x[3]:5d} d{x[5]-x[4]:5d}]"
        else:
            line += f" [{x[0]-t0:7d} a{x[1]-x[0]:5d}]"
    print(line)
