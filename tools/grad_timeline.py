"""Per-column timeline of one instance of the gradient tile kernel (needs a build with -DNAGP_GRAD_TRACE=n: the n-th
instance of block 0 is stamped): `NAGP_LIB=gpurun_exp/libnagp_gtrace.so python tools/grad_timeline.py`."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
n, k, P, K = 150, 1, 32, 1000
w = syn.make_workload(n, k, 0, K, P, seed=20261018 + 2)
th, nz = syn.perturbed_theta(w.ens, K, seed=77)
eng = Engine(0)
for _ in range(2):
    eng.logml_grad(w.ens, w.t[:n + k], w.y1, y2=w.y2, g=w.g[:n + k], step=w.step, theta=th, noise=nz)
N = 32 * 8 * 8 + 16
buf = (C.c_longlong * N)()
eng._lib.nagp_debug_read_grad.argtypes = [C.c_void_p, C.c_int]
assert eng._lib.nagp_debug_read_grad(buf, N) == 0
t = np.array(buf[:32 * 8 * 8]).reshape(32, 8, 8)
g = np.array(buf[32 * 8 * 8:])
nt = (n + k + 7) // 8
print(f"particle {g[15]}: load {g[1]-g[0]}, sweep {g[2]-g[1]}, tables {g[3]-g[2]}, entries {g[4]-g[3]}, table adjoints + reduction {g[5]-g[4]} cycles")
print("col | per warp: [wait B | finish diag | rows | wait A]   (cycles)")
for I in range(nt - 1, -1, -1):
    line = f"{I:2d} |"
    for wv in range(8):
        x = t[I, wv]
        line += f" [{x[1]-x[0]:5d} {x[2]-x[1]:4d} {x[3]-x[2]:5d} {x[4]-x[3]:5d}]"
    nxt = t[I - 1, 0, 0] if I > 0 else g[2]
    print(line + f"  column {nxt - t[I, 0, 0]:6d}")
print("per warp, relative to the start of the entry pass: [entries done, table adjoints done, reduction done]")
for wv in range(8):
    print("   ", [int(v - t[31, wv, 0]) for v in t[31, wv, 1:4]])
print("absolute stamps (relative to sweep start) for three columns: rows = warps, cols = [before B, after B, after finish, at A, after A]")
for I in (12, 9, 4):
    print("column", I)
    for wv in range(8):
        print("   ", [int(v - g[1]) for v in t[I, wv, :5]])
