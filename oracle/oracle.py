"""ctypes binding of the CPU oracle (oracle/nagp_oracle.c) plus a NumPy/SciPy twin.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by the product package. PARITY UNPINNED at the AutoGP boundary (see the C header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the oracle with oracle/Makefile (gcc only)."""
    args = ["make", "-C", _HERE] + (["-B"] if force else [])
    subprocess.run(args, check=True, capture_output=True)


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


def _opt(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype)
    return a.ctypes.data_as(C.c_void_p), a


class Oracle:
    """Thin typed wrapper; `quad=True` binds the __float128 build (same API, `_q` symbols)."""

    def __init__(self, quad: bool = False):
        self.quad = quad
        self.lib = _load("libnagp_oracle_q.so" if quad else "libnagp_oracle.so")
        self._sfx = "_q" if quad else ""

    def _fn(self, name):
        return getattr(self.lib, name + self._sfx)

    def prog_check(self, prog: bytes) -> int:
        f = self._fn("nagp_o_prog_check")
        f.restype = C.c_int64
        buf = np.frombuffer(prog, np.uint8)
        return f(buf.ctypes.data_as(C.c_void_p), C.c_int64(len(prog)))

    def kernel_pair(self, prog: bytes, theta, ti: float, tj: float, delta: Optional[float] = None) -> float:
        f = self._fn("nagp_o_kernel_pair")
        f.restype = C.c_double
        th = np.ascontiguousarray(theta, np.float64)
        buf = np.frombuffer(prog, np.uint8)
        d = abs(ti - tj) if delta is None else delta
        return f(buf.ctypes.data_as(C.c_void_p), C.c_int64(len(prog)), th.ctypes.data_as(C.c_void_p),
                 C.c_double(ti), C.c_double(tj), C.c_double(d))

    def gram(self, prog: bytes, theta, t, diag_lo=0.0, diag_hi=None, m=None, g=None, step=0.0):
        t = np.ascontiguousarray(t, np.float64)
        q = len(t)
        K = np.empty((q, q))
        th = np.ascontiguousarray(theta, np.float64)
        buf = np.frombuffer(prog, np.uint8)
        gp = _opt(g, np.int32)
        f = self._fn("nagp_o_gram")
        f.restype = C.c_int32
        rc = f(buf.ctypes.data_as(C.c_void_p), C.c_int64(len(prog)), th.ctypes.data_as(C.c_void_p),
               C.c_double(diag_lo), C.c_double(diag_lo if diag_hi is None else diag_hi),
               C.c_int64(q if m is None else m), C.c_int64(q), t.ctypes.data_as(C.c_void_p),
               gp[0] if gp else None, C.c_double(step), K.ctypes.data_as(C.c_void_p))
        if rc:
            raise ValueError(f"oracle gram failed rc={rc}")
        return K

    def logml(self, prog: bytes, theta, noise, t, y, jitter=1e-5, g=None, step=0.0):
        t = np.ascontiguousarray(t, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        th = np.ascontiguousarray(theta, np.float64)
        buf = np.frombuffer(prog, np.uint8)
        gp = _opt(g, np.int32)
        out = C.c_double()
        f = self._fn("nagp_o_logml")
        f.restype = C.c_int32
        info = f(buf.ctypes.data_as(C.c_void_p), C.c_int64(len(prog)), th.ctypes.data_as(C.c_void_p),
                 C.c_double(noise), C.c_double(jitter), C.c_int64(len(y)), t.ctypes.data_as(C.c_void_p),
                 gp[0] if gp else None, C.c_double(step), y.ctypes.data_as(C.c_void_p), C.byref(out))
        return out.value, info

    def _instance(self, fname, prog, theta, noise, n, k, h, t, y, ya, yb, jitter, noise_pred, g, step, joint):
        t = np.ascontiguousarray(t, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        th = np.ascontiguousarray(theta, np.float64)
        buf = np.frombuffer(prog, np.uint8)
        gp = _opt(g, np.int32)
        q = n + k + h
        lmn, lmm = C.c_double(), C.c_double()
        mu = np.full(h, np.nan)
        Ls = np.full((h, h), np.nan)
        f = self._fn(fname)
        f.restype = C.c_int32
        head = (buf.ctypes.data_as(C.c_void_p), C.c_int64(len(prog)), th.ctypes.data_as(C.c_void_p),
                C.c_double(noise), C.c_double(jitter), C.c_double(noise_pred),
                C.c_int64(n), C.c_int64(k), C.c_int64(h), t.ctypes.data_as(C.c_void_p),
                gp[0] if gp else None, C.c_double(step), y.ctypes.data_as(C.c_void_p))
        if joint:
            z = np.full(len(y), np.nan)
            Lt = np.full((k + h, q), np.nan)
            info = f(*head, C.c_int64(len(y)), C.c_double(ya), C.c_double(yb), C.byref(lmn), C.byref(lmm),
                     z.ctypes.data_as(C.c_void_p), Lt.ctypes.data_as(C.c_void_p),
                     mu.ctypes.data_as(C.c_void_p), Ls.ctypes.data_as(C.c_void_p))
            return dict(info=info, logml_n=lmn.value, logml_m=lmm.value, z=z, Ltail=Lt, mu=mu, L=Ls)
        info = f(*head, C.c_double(ya), C.c_double(yb), C.byref(lmn), C.byref(lmm),
                 mu.ctypes.data_as(C.c_void_p), Ls.ctypes.data_as(C.c_void_p))
        return dict(info=info, logml_n=lmn.value, logml_m=lmm.value, mu=mu, L=Ls)

    def instance_reference(self, prog, theta, noise, n, k, h, t, y, ya=1.0, yb=0.0, jitter=1e-5,
                           noise_pred=-1.0, g=None, step=0.0):
        """Reference schedule: rebuild(n) → add_data!(m) → predict_mvn (LU) → MvNormal chol."""
        return self._instance("nagp_o_instance_reference", prog, theta, noise, n, k, h, t, y, ya, yb,
                              jitter, noise_pred, g, step, joint=False)

    def instance_joint(self, prog, theta, noise, n, k, h, t, y, ya=1.0, yb=0.0, jitter=1e-5,
                       noise_pred=-1.0, g=None, step=0.0):
        """Joint factorisation (KERNEL_SPEC §6); len(y) is n or n+k."""
        return self._instance("nagp_o_instance_joint", prog, theta, noise, n, k, h, t, y, ya, yb,
                              jitter, noise_pred, g, step, joint=True)

    # ---- double-only entry points -------------------------------------------------------------
    def normalize(self, logw):
        logw = np.ascontiguousarray(logw, np.float64)
        w = np.empty_like(logw)
        ess = C.c_double()
        self.lib.nagp_o_normalize(C.c_int64(len(logw)), logw.ctypes.data_as(C.c_void_p),
                                  w.ctypes.data_as(C.c_void_p), C.byref(ess))
        return w, ess.value

    def draws(self, logw, mu, L, zeta, comp=None, u=None, u_res=None, ess_thr=0.0):
        """mu [K,P,h] or [P,h] (shared); L [K,P,h,h] or [P,h,h]; zeta [K,D,h] → x [h, K*D] (Fortran)."""
        logw = np.ascontiguousarray(logw, np.float64)
        K, P = logw.shape
        zeta = np.ascontiguousarray(zeta, np.float64)
        D, h = zeta.shape[1], zeta.shape[2]
        mu = np.ascontiguousarray(mu, np.float64)
        L = np.ascontiguousarray(L, np.float64)
        mu_stride = P * h if mu.ndim == 3 else 0
        l_stride = P * h * h if L.ndim == 4 else 0
        x = np.empty((K * D, h))
        ess = np.empty(K)
        comp_out = np.empty((K, D), np.int32)
        cp, up, rp = _opt(comp, np.int32), _opt(u, np.float64), _opt(u_res, np.float64)
        self.lib.nagp_o_draws(
            C.c_int64(K), C.c_int64(P), C.c_int64(h), C.c_int64(D), logw.ctypes.data_as(C.c_void_p),
            mu.ctypes.data_as(C.c_void_p), C.c_int64(mu_stride), L.ctypes.data_as(C.c_void_p),
            C.c_int64(l_stride), cp[0] if cp else None, up[0] if up else None, rp[0] if rp else None,
            C.c_double(ess_thr), zeta.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
            ess.ctypes.data_as(C.c_void_p), comp_out.ctypes.data_as(C.c_void_p))
        return x.T, ess, comp_out  # x.T is [h, K*D] column-major view

    def forecast_instances(self, ens, n, k, h, t, y1, y2, logw0, ya=1.0, yb=0.0, jitter=1e-5,
                           noise_pred=-1.0, g=None, step=0.0, use_joint=False, theta_per_scenario=None,
                           noise_per_scenario=None):
        """K×P instances. `ens` is a FlatEnsemble of P kernels; theta_per_scenario [K, total_theta]
        and noise_per_scenario [K,P] switch on per-(scenario, particle) hyperparameters."""
        y2 = np.ascontiguousarray(y2, np.float64)
        K = y2.shape[0]
        P = ens.size
        t = np.ascontiguousarray(t, np.float64)
        y1 = np.ascontiguousarray(y1, np.float64)
        logw0 = np.ascontiguousarray(logw0, np.float64)
        theta = ens.theta if theta_per_scenario is None else np.ascontiguousarray(theta_per_scenario, np.float64)
        noise = ens.noise if noise_per_scenario is None else np.ascontiguousarray(noise_per_scenario, np.float64)
        gp = _opt(g, np.int32)
        logw = np.empty((K, P))
        mu = np.full((K, P, h), np.nan)
        Ls = np.full((K, P, h, h), np.nan)
        info = np.zeros((K, P), np.int32)
        f = self.lib.nagp_o_forecast_instances
        f.restype = C.c_int32
        rc = f(C.c_int64(K), C.c_int64(P), ens.prog.ctypes.data_as(C.c_void_p),
               ens.prog_off.ctypes.data_as(C.c_void_p), theta.ctypes.data_as(C.c_void_p),
               ens.theta_off.ctypes.data_as(C.c_void_p),
               C.c_int64(0 if theta_per_scenario is None else theta.shape[1]),
               noise.ctypes.data_as(C.c_void_p), C.c_int64(0 if noise_per_scenario is None else P),
               C.c_double(jitter), C.c_double(noise_pred), C.c_int64(n), C.c_int64(k), C.c_int64(h),
               t.ctypes.data_as(C.c_void_p), gp[0] if gp else None, C.c_double(step),
               y1.ctypes.data_as(C.c_void_p), y2.ctypes.data_as(C.c_void_p), C.c_double(ya), C.c_double(yb),
               logw0.ctypes.data_as(C.c_void_p), C.c_int32(int(use_joint)),
               logw.ctypes.data_as(C.c_void_p), mu.ctypes.data_as(C.c_void_p),
               Ls.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p))
        return dict(rc=rc, logw=logw, mu=mu, L=Ls, info=info)

    def logml_batch(self, ens, t, y, jitter=1e-5, g=None, step=0.0):
        t = np.ascontiguousarray(t, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        gp = _opt(g, np.int32)
        B = ens.size
        out = np.empty(B)
        info = np.zeros(B, np.int32)
        f = self.lib.nagp_o_logml_batch
        f.restype = C.c_int32
        f(C.c_int64(B), ens.prog.ctypes.data_as(C.c_void_p), ens.prog_off.ctypes.data_as(C.c_void_p),
          ens.theta.ctypes.data_as(C.c_void_p), ens.theta_off.ctypes.data_as(C.c_void_p),
          ens.noise.ctypes.data_as(C.c_void_p), C.c_double(jitter), C.c_int64(len(y)),
          t.ctypes.data_as(C.c_void_p), gp[0] if gp else None, C.c_double(step),
          y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p))
        return out, info

    def set_num_threads(self, n: int) -> None:
        self.lib.nagp_o_set_num_threads(int(n))

    def num_threads(self) -> int:
        return int(self.lib.nagp_o_num_threads())


# ------------------------------------------------------------------------------------------------
# NumPy/SciPy twin: an independent second restatement (vectorised Gram, LAPACK dpotrf through
# scipy — the routine Julia itself calls), used to cross-check the C oracle.
# ------------------------------------------------------------------------------------------------
def gram_np(prog: bytes, theta, t, g=None, step=0.0) -> np.ndarray:
    t = np.asarray(t, np.float64)
    ti, tj = t[:, None], t[None, :]
    if g is None:
        delta = np.abs(ti - tj)
    else:
        g = np.asarray(g, np.int64)
        delta = np.abs(g[:, None] - g[None, :]).astype(np.float64) * step
    st, pos = [], 0
    th = list(theta)
    for op in prog:
        if op == 1:
            st.append(np.full(delta.shape, th[pos])); pos += 1
        elif op == 2:
            c, b, a = th[pos:pos + 3]; pos += 3
            st.append(b + a * ((ti - c) * (tj - c)))
        elif op == 3:
            l, a = th[pos:pos + 2]; pos += 2
            st.append(a * np.exp(-0.5 * (delta / l) ** 2))
        elif op == 4:
            l, gam, a = th[pos:pos + 3]; pos += 3
            st.append(a * np.exp(-np.power(delta / l, gam)))
        elif op == 5:
            l, p, a = th[pos:pos + 3]; pos += 3
            st.append(a * np.exp(-2.0 * np.sin(np.pi * (delta / p)) ** 2 / (l * l)))
        elif op == 6:
            r = st.pop(); st[-1] = st[-1] + r
        elif op == 7:
            r = st.pop(); st[-1] = st[-1] * r
        elif op == 8:
            loc, sc = th[pos:pos + 2]; pos += 2
            si = 0.5 * (1.0 + np.tanh((ti - loc) / sc))
            sj = 0.5 * (1.0 + np.tanh((tj - loc) / sc))
            r = st.pop(); l_ = st.pop()
            st.append((1 - si) * (1 - sj) * l_ + si * sj * r)
        else:
            raise ValueError(op)
    return st[0]


def logml_np(prog: bytes, theta, noise, t, y, jitter=1e-5, g=None, step=0.0) -> float:
    from scipy.linalg import cho_factor, solve_triangular
    y = np.asarray(y, np.float64)
    K = gram_np(prog, theta, t, g, step) + (noise + jitter) * np.eye(len(y))
    L = cho_factor(K, lower=True)[0]
    z = solve_triangular(np.tril(L), y, lower=True)
    return float(-0.5 * (len(y) * np.log(2 * np.pi) + 2 * np.log(np.diag(L)).sum() + z @ z))


# ------------------------------------------------------------------------------------------------
# The TIMED CPU baseline (oracle/nagp_cpu_blocked.c): the reference schedule with blocked, AVX2/FMA-vectorised
# Cholesky / LU at -O3. Checked against the Oracle above in tests/test_cpu_baseline.py; used by bench.py's
# cpu_baseline and --impl reference legs only.
# ------------------------------------------------------------------------------------------------
class BlockedCpu:
    def __init__(self):
        self.lib = _load("libnagp_cpu_blocked.so")

    def set_num_threads(self, n: int) -> None:
        self.lib.nagp_b_set_num_threads(int(n))

    def num_threads(self) -> int:
        return int(self.lib.nagp_b_num_threads())

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def factor_rates(self, n: int, reps: int = 20):
        """Single-thread GFLOP/s of the blocked Cholesky and LU alone at order n."""
        c, l = C.c_double(), C.c_double()
        self.lib.nagp_b_factor_rates(C.c_int64(n), C.c_int64(reps), C.byref(c), C.byref(l))
        return c.value, l.value

    def forecast_instances(self, ens, n, k, h, t, y1, y2, logw0, ya=1.0, yb=0.0, jitter=1e-5, noise_pred=-1.0,
                           g=None, step=0.0, theta_per_scenario=None, noise_per_scenario=None):
        """Reference schedule (three factorisations + LU solves + h x h Cholesky per instance), K x P instances."""
        y2 = np.ascontiguousarray(y2, np.float64)
        K, P = y2.shape[0], ens.size
        t = np.ascontiguousarray(t, np.float64)
        y1 = np.ascontiguousarray(y1, np.float64)
        logw0 = np.ascontiguousarray(logw0, np.float64)
        theta = ens.theta if theta_per_scenario is None else np.ascontiguousarray(theta_per_scenario, np.float64)
        noise = ens.noise if noise_per_scenario is None else np.ascontiguousarray(noise_per_scenario, np.float64)
        gp = _opt(g, np.int32)
        logw, mu = np.empty((K, P)), np.full((K, P, h), np.nan)
        Ls, info = np.full((K, P, h, h), np.nan), np.zeros((K, P), np.int32)
        f = self.lib.nagp_b_forecast_instances
        f.restype = C.c_int32
        rc = f(C.c_int64(K), C.c_int64(P), self._p(ens.prog), self._p(ens.prog_off), self._p(theta), self._p(ens.theta_off),
               C.c_int64(0 if theta_per_scenario is None else theta.shape[1]), self._p(noise),
               C.c_int64(0 if noise_per_scenario is None else P), C.c_double(jitter), C.c_double(noise_pred),
               C.c_int64(n), C.c_int64(k), C.c_int64(h), self._p(t), gp[0] if gp else None, C.c_double(step),
               self._p(y1), self._p(y2), C.c_double(ya), C.c_double(yb), self._p(logw0), self._p(logw), self._p(mu),
               self._p(Ls), self._p(info))
        return dict(rc=rc, logw=logw, mu=mu, L=Ls, info=info)

    def logml_batch(self, ens, t, y, jitter=1e-5, g=None, step=0.0):
        t = np.ascontiguousarray(t, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        gp = _opt(g, np.int32)
        B = ens.size
        out, info = np.empty(B), np.zeros(B, np.int32)
        f = self.lib.nagp_b_logml_batch
        f.restype = C.c_int32
        f(C.c_int64(B), self._p(ens.prog), self._p(ens.prog_off), self._p(ens.theta), self._p(ens.theta_off),
          self._p(ens.noise), C.c_double(jitter), C.c_int64(len(y)), self._p(t), gp[0] if gp else None,
          C.c_double(step), self._p(y), self._p(out), self._p(info))
        return out, info

    def factor_store(self, ens, t, y, n_cap, jitter=1e-5, g=None, step=0.0):
        """Factor every instance on the first len(y) points and keep L [B, n_cap, n_cap] and z [B, n_cap]."""
        t = np.ascontiguousarray(t, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        gp = _opt(g, np.int32)
        B, n = ens.size, len(y)
        L, z, lm = np.zeros((B, n_cap, n_cap)), np.zeros((B, n_cap)), np.empty(B)
        f = self.lib.nagp_b_factor_store
        f.restype = C.c_int32
        rc = f(C.c_int64(B), self._p(ens.prog), self._p(ens.prog_off), self._p(ens.theta), self._p(ens.theta_off),
               self._p(ens.noise), C.c_double(jitter), C.c_int64(n), C.c_int64(n_cap), self._p(t),
               gp[0] if gp else None, C.c_double(step), self._p(y), self._p(L), self._p(z), self._p(lm))
        return dict(rc=rc, L=L, z=z, logml=lm, n=n, n_cap=n_cap)

    def append(self, ens, store, t, y, k, jitter=1e-5, g=None, step=0.0):
        """Rank-append of k points to every stored factor in place; t/g/y cover all store['n'] + k points."""
        t = np.ascontiguousarray(t, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        gp = _opt(g, np.int32)
        B = ens.size
        dl = np.empty(B)
        f = self.lib.nagp_b_append
        f.restype = C.c_int32
        rc = f(C.c_int64(B), self._p(ens.prog), self._p(ens.prog_off), self._p(ens.theta), self._p(ens.theta_off),
               self._p(ens.noise), C.c_double(jitter), C.c_int64(store["n"]), C.c_int64(k), C.c_int64(store["n_cap"]),
               self._p(t), gp[0] if gp else None, C.c_double(step), self._p(y), self._p(store["L"]), self._p(store["z"]),
               self._p(dl))
        store["n"] += k
        return dict(rc=rc, dlogml=dl)
