"""Phase and per-column timeline of one instance of the tile kernel (needs a build with -DNAGP_V2_TRACE=n: the n-th
instance of block 0 is stamped): `NAGP_LIB=gpurun_exp/libnagp_v2trace.so python tools/v2_timeline.py`."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from nowcastautogp_b200.engine import Engine
w, th, nz, z, u = bench.make_inputs(0)
c = bench.CFG
eng = Engine(0)
eng.set_variant(2)
for _ in range(2):
    eng.forecast_instances(w.ens, c["n"], c["k"], c["h"], w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step, theta=th, noise=nz)
N = 64 + 64 + 8 + 256
buf = (C.c_longlong * N)()
eng._lib.nagp_debug_read_v2.argtypes = [C.c_void_p, C.c_int]
assert eng._lib.nagp_debug_read_v2(buf, N) == 0
col = np.array(buf[:64]).reshape(32, 2)
ph = np.array(buf[64:128]).reshape(8, 8)
p = buf[128]
prog = list(w.ens.prog[w.ens.prog_off[p]:w.ens.prog_off[p + 1]])
t0 = ph[:, 0].min()
names = ["work item", "program+theta", "tables", "Gram", "Cholesky", "epilogue"]
print(f"particle {p} program {prog}")
print("phase arrivals per warp (cycles from the start of the instance; a phase ends when its last warp arrives):")
for i in range(1, 6):
    print(f"  {names[i]:14s} last {int(ph[:, i].max() - t0):7d}  first {int(ph[:, i].min() - t0):7d}   phase length {int(ph[:, i].max() - ph[:, i - 1].max()):7d}")
nt = (c["n"] + c["k"] + c["h"] + 7) // 8
own = np.array(buf[136:136 + 256]).reshape(32, 8)
print("col | inverse published (from the start of the Cholesky) | since the previous | hand-over of the next diagonal tile after publication | factorisation after the hand-over")
c0 = ph[:, 3].max()
for J in range(nt):
    e, hnd = col[J]
    prev = col[J - 1, 0] if J else c0
    ho = f"{int(hnd - e):6d}" if J + 1 < nt else "     -"
    fac = f"{int(e - col[J - 1, 1]):6d}" if J else "     -"
    slack = [int(e - v) for v in own[J] if v > 0]      # > 0: the row owner was already waiting when the inverse came
    print(f"{J:3d} | {int(e - c0):7d} | {int(e - prev):6d} | {ho} | {fac} | row owners waiting since (cycles before publication): {sorted(slack)}")
