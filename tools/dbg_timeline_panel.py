"""Per-column timeline of the tile kernel's panel schedule from an NAGP_EXP=9 build (clock64 stamps of CTA 0,
instance 0): chain warp 0 and the six row-owning warps."""
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
t = d[:20 * 64].reshape(20, 8, 8)[:, :, :6]
pub = t[:, 0, 2]
print("publish-to-publish (chain period):", np.diff(pub))
bw = [1, 2, 3, 5, 6, 7]
print("col | chain warp: wait for C + load, factor, store+publish | row owners (median): wait inverse + solve, wait rows + last term + C, lookahead, total")
for J in range(20):
    med = lambda a, b_: np.median([t[J, x, b_] - t[J, x, a] for x in bw])
    nxt = np.median([t[J + 1, x, 0] - t[J, x, 0] for x in bw]) if J < 19 else 0
    print(f"{J:2d} | {int(t[J,0,4]-t[J,0,1]):5d} {int(t[J,0,3]-t[J,0,4]):5d} {int(t[J,0,2]-t[J,0,3]):4d} | "
          f"{med(0,4):6.0f} {med(4,1):6.0f} {med(1,2):6.0f} {nxt:6.0f}")
