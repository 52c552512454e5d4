"""Per-column timeline of the tile kernel from an NAGP_EXP=9 build (clock64 stamps of CTA 0, instance 0)."""
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
print("instance 0 (program:", bytes(w.ens.prog[w.ens.prog_off[0]:w.ens.prog_off[1]]).hex(), ")")
print("phases: setup %d tables %d gram %d factor %d logml %d" % tuple(np.diff(g[:6])))
t = d[:20 * 64].reshape(20, 8, 8)[:, :, :6]
base = t[0, :, 0].min()
print("col owner | per-warp [top->C, C->pre-bar, bar wait, trsm, sync wait] (owner's row) | column wall")
starts = [t[J, J % 8, 1] for J in range(20)]
print("owner C-ready to next owner C-ready (chain period):", np.diff(starts))
for J in range(20):
    ow = J % 8
    row = t[J, ow]
    others = [x for x in range(8) if x != ow]
    look = np.median([t[J, x, 2] - t[J, x, 1] for x in others])
    wait = np.median([t[J, x, 3] - t[J, x, 2] for x in others])
    wall = t[J, :, 5].max() - t[J, :, 0].min()
    print(f"{J:2d} w{ow} | acc+C {row[1]-row[0]:5d} chol {row[2]-row[1]:5d} trsm {row[4]-row[3]:5d} sync {row[5]-row[4]:5d} | "
          f"others: lookahead {look:6.0f} barwait {wait:6.0f} | wall {wall:6d}")
