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
for J in range(20):
    wait = t[J, 0, 1] - (t[J - 1, 0, 2] if J else t[J, 0, 1])
    chol = t[J, 0, 2] - t[J, 0, 1]
    own = bw[(J + 1) % 6]
    hand = (t[J + 1, 0, 1] - t[J, own, 3]) if J < 19 else 0
    med = lambda a, b_: np.median([t[J, x, b_] - t[J, x, a] for x in bw])
    print(f"{J:2d} | chain: wait {wait:5d} chol {chol:5d} | owner of next row: inverse seen -> chain has C {hand:5d} | rows (median): "
          f"finish+C {med(0,1):5.0f} lookahead {med(1,2):5.0f} wait-inverse {med(2,3):5.0f} solve {med(3,4):5.0f} wait-rows {med(4,5):5.0f}")

n = int(d[8100])
print("SM 0 residents (block, warp, hardware warp slot):", sorted(((int(v) >> 32), (int(v) >> 8) & 0xff, int(v) & 0xff) for v in d[8101:8101 + min(n, 40)]))
print("chain warp, per column [after barrier -> tile loaded -> factored -> stored+published]:")
for J in range(20):
    print("   ", J, int(t[J, 0, 4] - t[J, 0, 1]), int(t[J, 0, 3] - t[J, 0, 4]), int(t[J, 0, 2] - t[J, 0, 3]))
t0 = t[0, 0, 1]
for J in range(4):
    print("col", J, "chain [C, published]:", (t[J, 0, 1:3] - t0).tolist())
    for x in bw:
        print("    row warp", x, "[top, C, lookahead done, inverse seen, solved, rows seen]:", (t[J, x, :6] - t0).tolist())
