"""Per-block-column timeline of one instance of chol_large_kernel (needs a build with -DNAGP_LARGE_TRACE=n):
`NAGP_LIB=gpurun_exp/libnagp_ltrace.so python tools/large_timeline.py [B] [n]`."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
B, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 512
w = syn.make_workload(n, 0, 0, 1, B, seed=20261018 + 3, max_depth=4, period=365.0)
eng = Engine(0)
for i in range(2):
    eng.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
N = 128 * 8 * 4 + 128 * 4
buf = (C.c_longlong * N)()
eng._lib.nagp_debug_read_large.argtypes = [C.c_void_p, C.c_int]
assert eng._lib.nagp_debug_read_large(buf, N) == 0
wt = np.array(buf[:128 * 32]).reshape(128, 8, 4)
bc = np.array(buf[128 * 32:]).reshape(128, 4)
nbc = ((n + 63) // 64) * 2          # block columns of 4 tiles
print("block col | length | diag block: starts after, takes | per warp: [busy until, waited for the diagonal block, row groups solved]")
for J in range(nbc):
    t0 = bc[J, 0]
    end = wt[J, :, 0].max()
    line = f"{J:3d} | {int(end - t0):8d} | {int(bc[J, 1] - t0):7d} {int(bc[J, 2] - bc[J, 1]):7d} |"
    for wv in range(8):
        line += f" [{int(wt[J, wv, 0] - t0):7d} {int(wt[J, wv, 1]):6d} {int(wt[J, wv, 2]):2d}]"
    print(line)
