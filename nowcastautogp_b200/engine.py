"""Typed Python face of the C ABI (include/nagp.h).

Arrays may be NumPy arrays (host) or torch CUDA tensors (device); the library detects which from
the raw pointer. Outputs are allocated as NumPy arrays unless an `out=`-style tensor is passed.
PyTorch is only plumbing here (device memory, streams); every kernel is in libnagp.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .kernels import FlatEnsemble


class NagpError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libnagp error {code}: {message}")
        self.code = code


class PosDefError(ArithmeticError):
    """A Gram / predictive covariance was not positive definite (Julia: PosDefException,
    `/root/reference/test/test_model_fitting.jl:97-98`)."""

    def __init__(self, info: int):
        super().__init__(f"matrix is not positive definite; leading minor of order {info}")
        self.info = info


def _ptr(a, dtype=None):
    """(pointer, keepalive) for a NumPy array, torch tensor or None."""
    if a is None:
        return None, None
    if isinstance(a, np.ndarray):
        if dtype is not None and a.dtype != dtype:
            a = a.astype(dtype)
        if not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a)
        return a.ctypes.data, a
    if hasattr(a, "data_ptr"):  # torch tensor (any device)
        if not a.is_contiguous():
            raise ValueError("torch tensors passed to libnagp must be contiguous")
        return a.data_ptr(), a
    a = np.ascontiguousarray(a, dtype)
    return a.ctypes.data, a


class Engine:
    """One libnagp context (one GPU, one stream). Not thread-safe: one call at a time."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        ctx = C.c_void_p()
        rc = self._lib.nagp_init(device, C.byref(ctx))
        if rc != 0:
            raise NagpError(rc, self._lib.nagp_last_error(None).decode())
        self._ctx = ctx
        self.device = device

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.nagp_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------
    def _check(self, rc: int, raise_posdef: bool = True) -> int:
        if rc < 0:
            raise NagpError(rc, self._lib.nagp_last_error(self._ctx).decode())
        if rc > 0 and raise_posdef:
            raise PosDefError(rc)
        return rc

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        self._check(self._lib.nagp_set_stream(self._ctx, cuda_stream))

    def set_jitter(self, jitter: float) -> None:
        self._check(self._lib.nagp_set_jitter(self._ctx, jitter))

    def set_variant(self, variant: int) -> None:
        self._check(self._lib.nagp_set_variant(self._ctx, variant))

    @property
    def last_kernel(self) -> int:
        """Factorisation kernel of the last fused launch: 1 column, 2 tile, 3 slot, 4 large."""
        return int(self._lib.nagp_last_kernel(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self._lib.nagp_launch_count(self._ctx))

    # ---- (a2) ------------------------------------------------------------------------------------
    def logml_batch(self, ens: FlatEnsemble, t, y, g=None, step: float = 0.0, y_stride: int = 0,
                    logml=None, info=None, check: bool = False):
        B = ens.size
        n = len(t) if not hasattr(t, "numel") else t.numel()
        logml = np.empty(B) if logml is None else logml
        info = np.zeros(B, np.int32) if info is None else info
        keep = [_ptr(ens.prog), _ptr(ens.prog_off), _ptr(ens.theta), _ptr(ens.theta_off), _ptr(ens.noise),
                _ptr(t, np.float64), _ptr(g, np.int32), _ptr(y, np.float64), _ptr(logml), _ptr(info)]
        p = [k[0] for k in keep]
        rc = self._lib.nagp_logml_batch(self._ctx, B, p[0], p[1], p[2], p[3], p[4], n, p[5], p[6], step,
                                        p[7], y_stride, p[8], p[9])
        self._check(rc, raise_posdef=check)
        return logml, info

    # ---- general per-(scenario, particle) path -----------------------------------------------------
    def forecast_instances(self, ens: FlatEnsemble, n, k, h, t, y1, y2, logw0, ya=1.0, yb=0.0, g=None,
                           step=0.0, noise_pred=-1.0, theta=None, noise=None, K=None,
                           logw=None, mu=None, L=None, info=None, check: bool = False,
                           logml_n=None, logml_m=None, want_moments: bool = True):
        """`theta` [K,total] / `noise` [K,P] switch on per-scenario hyperparameters."""
        P = ens.size
        if K is None:
            K = y2.shape[0]
        logw = np.empty((K, P)) if logw is None else logw
        if want_moments:
            mu = np.empty((K, P, h)) if mu is None else mu
            L = np.empty((K, P, h, h)) if L is None else L
        info = np.zeros((K, P), np.int32) if info is None else info
        th = ens.theta if theta is None else theta
        nz = ens.noise if noise is None else noise
        keep = [_ptr(ens.prog), _ptr(ens.prog_off), _ptr(th), _ptr(ens.theta_off), _ptr(nz),
                _ptr(t, np.float64), _ptr(g, np.int32), _ptr(y1, np.float64), _ptr(y2, np.float64),
                _ptr(logw0, np.float64), _ptr(logw), _ptr(mu), _ptr(L), _ptr(info), _ptr(logml_n), _ptr(logml_m)]
        p = [x[0] for x in keep]
        rc = self._lib.nagp_forecast_instances(
            self._ctx, K, P, p[0], p[1], p[2], p[3], 0 if theta is None else int(ens.theta_off[-1]),
            p[4], 0 if noise is None else P, noise_pred, n, k, h, p[5], p[6], step, p[7], p[8], ya, yb,
            p[9], p[10], p[11], p[12], p[13], p[14], p[15])
        self._check(rc, raise_posdef=check)
        return dict(logw=logw, mu=mu, L=L, info=info, logml_n=logml_n, logml_m=logml_m)

    # ---- (f1) gradient of the log marginal likelihood -----------------------------------------------
    def logml_grad(self, ens: FlatEnsemble, t, y1, y2=None, g=None, step: float = 0.0, theta=None, noise=None,
                   K: int = 1, check: bool = False, out=None, y_stride: int = 0):
        """logML over the n + k points [y1 | y2[s]] and its gradient w.r.t. every theta slot and the noise.
        Returns (logml [K,P], grad_theta [K,total], grad_noise [K,P], info [K,P]); `out` may supply those four
        buffers (host or device) to keep the call free of host copies."""
        P = ens.size
        n = int(y_stride) if y_stride else len(y1)
        k = 0 if y2 is None else int(np.shape(y2)[-1])
        if y2 is not None:
            K = int(np.shape(y2)[0])
        elif theta is not None:
            K = int(np.shape(theta)[0])
        total = int(ens.theta_off[-1])
        if out is not None:
            logml, gth, gnz, info = out
        else:
            logml, gth = np.empty((K, P)), np.empty((K, total))
            gnz, info = np.empty((K, P)), np.zeros((K, P), np.int32)
        th = ens.theta if theta is None else theta
        nz = ens.noise if noise is None else noise
        keep = [_ptr(ens.prog), _ptr(ens.prog_off), _ptr(th, np.float64), _ptr(ens.theta_off), _ptr(nz, np.float64),
                _ptr(t, np.float64), _ptr(g, np.int32), _ptr(y1, np.float64), _ptr(y2, np.float64),
                _ptr(logml), _ptr(gth), _ptr(gnz), _ptr(info)]
        p = [x[0] for x in keep]
        rc = self._lib.nagp_logml_grad(self._ctx, K, P, p[0], p[1], p[2], p[3], 0 if theta is None else total,
                                       p[4], 0 if noise is None else P, n, k, p[5], p[6], step, p[7], int(y_stride),
                                       p[8], p[9], p[10], p[11], p[12])
        self._check(rc, raise_posdef=check)
        return logml, gth, gnz, info

    def hmc(self, prog, prog_off, theta_off, slot_kind, slot_a, slot_b, noise_spec, z, noise_z, t, y1, y2=None,
            g=None, step: float = 0.0, y_stride: int = 0, n_leapfrog: int = 10, eps: float = 0.02, momenta=None,
            noise_momenta=None, log_u=None):
        """`nagp_hmc`: K x P chains, `len(log_u)` iterations, integrator on the device. z [K,total] and noise_z
        [K,P] are updated in place. Returns (logml [K,P], n_accept [K,P], info [K,P])."""
        z = np.ascontiguousarray(z, np.float64)
        noise_z = np.ascontiguousarray(noise_z, np.float64)
        K, P = noise_z.shape
        n = int(y_stride) if y_stride else len(y1)
        k = 0 if y2 is None else int(np.shape(y2)[-1])
        n_steps = 0 if log_u is None else len(log_u)
        logml, nacc, info = np.empty((K, P)), np.zeros((K, P), np.int32), np.zeros((K, P), np.int32)
        nk, na, nb = noise_spec
        keep = [_ptr(np.ascontiguousarray(prog, np.uint8)), _ptr(np.ascontiguousarray(prog_off, np.int64)),
                _ptr(np.ascontiguousarray(theta_off, np.int64)), _ptr(np.ascontiguousarray(slot_kind, np.int32)),
                _ptr(np.ascontiguousarray(slot_a, np.float64)), _ptr(np.ascontiguousarray(slot_b, np.float64)),
                _ptr(z), _ptr(noise_z), _ptr(t, np.float64), _ptr(g, np.int32), _ptr(y1, np.float64),
                _ptr(y2, np.float64), _ptr(momenta, np.float64), _ptr(noise_momenta, np.float64),
                _ptr(log_u, np.float64), _ptr(logml), _ptr(nacc), _ptr(info)]
        p = [x[0] for x in keep]
        rc = self._lib.nagp_hmc(self._ctx, K, P, p[0], p[1], p[2], p[3], p[4], p[5], int(nk), float(na), float(nb),
                                p[6], p[7], n, k, p[8], p[9], step, p[10], int(y_stride), p[11], n_steps,
                                int(n_leapfrog), float(eps), p[12], p[13], p[14], p[15], p[16], p[17])
        self._check(rc, raise_posdef=False)
        return z, noise_z, logml, nacc, info

    # ---- (a3)/(a4)/(a7) ----------------------------------------------------------------------------
    def factor_store(self, ens: FlatEnsemble, n, k, h, t, y1, logw0=None, ya=1.0, yb=0.0, g=None, step=0.0,
                     noise_pred=-1.0, check: bool = True) -> "Factor":
        P = ens.size
        logml_n = np.empty(P)
        info = np.zeros(P, np.int32)
        keep = [_ptr(ens.prog), _ptr(ens.prog_off), _ptr(ens.theta), _ptr(ens.theta_off), _ptr(ens.noise),
                _ptr(t, np.float64), _ptr(g, np.int32), _ptr(y1, np.float64), _ptr(logw0, np.float64)]
        p = [x[0] for x in keep]
        handle = C.c_void_p()
        rc = self._lib.nagp_factor_store(self._ctx, P, p[0], p[1], p[2], p[3], p[4], noise_pred, n, k, h,
                                         p[5], p[6], step, p[7], ya, yb, p[8], C.byref(handle),
                                         logml_n.ctypes.data, info.ctypes.data)
        f = Factor(self, handle, P, n, k, h, logml_n, info)
        self._check(rc, raise_posdef=check)
        return f

    def factor_store_large(self, ens: FlatEnsemble, t, y, capacity=None, g=None, step=0.0,
                           check: bool = True) -> "Factor":
        """Appendable whole-factor store for long series (`nagp_factor_store_large`)."""
        P = ens.size
        n = len(t)
        capacity = n if capacity is None else int(capacity)
        logml = np.empty(P)
        info = np.zeros(P, np.int32)
        keep = [_ptr(ens.prog), _ptr(ens.prog_off), _ptr(ens.theta), _ptr(ens.theta_off), _ptr(ens.noise),
                _ptr(np.asarray(t), np.float64), _ptr(None if g is None else np.asarray(g), np.int32),
                _ptr(np.asarray(y), np.float64)]
        p = [x[0] for x in keep]
        handle = C.c_void_p()
        rc = self._lib.nagp_factor_store_large(self._ctx, P, p[0], p[1], p[2], p[3], p[4], n, capacity, p[5], p[6],
                                               step, p[7], C.byref(handle), logml.ctypes.data, info.ctypes.data)
        f = Factor(self, handle, P, n, 0, 0, logml, info)
        self._check(rc, raise_posdef=check)
        return f

    def factor_append(self, factor: "Factor", t_new, y_new, g_new=None, check: bool = True):
        """Rank-append `len(t_new)` points in place; returns (dlogml[P], logml[P], info[P])."""
        kn_ = len(t_new)
        dl, lm = np.empty(factor.P), np.empty(factor.P)
        info = np.zeros(factor.P, np.int32)
        keep = [_ptr(np.asarray(t_new), np.float64), _ptr(None if g_new is None else np.asarray(g_new), np.int32),
                _ptr(np.asarray(y_new), np.float64)]
        rc = self._lib.nagp_factor_append(self._ctx, factor._h, kn_, keep[0][0], keep[1][0], keep[2][0],
                                          dl.ctypes.data, lm.ctypes.data, info.ctypes.data)
        self._check(rc, raise_posdef=check)
        factor.n = int(self._lib.nagp_factor_size(factor._h))
        factor.logml_n = lm
        return dl, lm, info

    def append(self, factor: "Factor", y2, logw=None, mu=None, want_mu: bool = True):
        K = y2.shape[0]
        logw = np.empty((K, factor.P)) if logw is None else logw
        if mu is None and want_mu:
            mu = np.empty((K, factor.P, factor.h))
        keep = [_ptr(y2, np.float64), _ptr(logw), _ptr(mu)]
        self._check(self._lib.nagp_append(self._ctx, factor._h, K, keep[0][0], keep[1][0], keep[2][0]))
        return logw, mu

    def predict(self, factor: "Factor", want_mu: bool = True):
        mu = np.empty((factor.P, factor.h)) if want_mu else None
        L = np.empty((factor.P, factor.h, factor.h))
        self._check(self._lib.nagp_predict(self._ctx, factor._h, None if mu is None else mu.ctypes.data,
                                           L.ctypes.data))
        return mu, L

    # ---- (a5)/(a8) ---------------------------------------------------------------------------------
    def ess(self, logw):
        logw = np.ascontiguousarray(logw, np.float64)
        K, P = logw.shape
        ess = np.empty(K)
        w = np.empty((K, P))
        self._check(self._lib.nagp_ess(self._ctx, K, P, logw.ctypes.data, ess.ctypes.data, w.ctypes.data))
        return ess, w

    def draw(self, logw, mu, L, zeta, comp=None, u=None, u_res=None, ess_thr=0.0, x=None, ess=None,
             comp_out=None, want_aux: bool = True):
        """mu [K,P,h] or [P,h] (shared); L [K,P,h,h] or [P,h,h]; zeta [K,D,h] → x [h, K·D].
        With `want_aux=False` and device `x` the call stays asynchronous on the stream."""
        K, P = logw.shape
        D, h = zeta.shape[1], zeta.shape[2]
        mu_stride = P * h if len(mu.shape) == 3 else 0
        l_stride = P * h * h if len(L.shape) == 4 else 0
        xbuf = np.empty((K * D, h)) if x is None else x
        if want_aux:
            ess = np.empty(K) if ess is None else ess
            comp_out = np.empty((K, D), np.int32) if comp_out is None else comp_out
        keep = [_ptr(logw, np.float64), _ptr(mu, np.float64), _ptr(L, np.float64), _ptr(comp, np.int32),
                _ptr(u, np.float64), _ptr(u_res, np.float64), _ptr(zeta, np.float64), _ptr(xbuf),
                _ptr(ess), _ptr(comp_out)]
        p = [k_[0] for k_ in keep]
        self._check(self._lib.nagp_draw(self._ctx, K, P, h, D, p[0], p[1], mu_stride, p[2], l_stride, p[3],
                                        p[4], p[5], ess_thr, p[6], p[7], p[8], p[9]))
        return (xbuf.T if x is None else xbuf), ess, comp_out

    # ---- fused forecast_with_nowcasts --------------------------------------------------------------
    def forecast_with_nowcasts(self, ens: FlatEnsemble, n, k, h, t, y1, y2, logw0, zeta, ya=1.0, yb=0.0,
                               g=None, step=0.0, noise_pred=-1.0, comp=None, u=None, u_res=None,
                               ess_thr=0.0, x=None, logw=None, ess=None, info=None, check: bool = True,
                               K=None, D=None):
        P = ens.size
        if K is None:
            K = zeta.shape[0]
        if D is None:
            D = zeta.shape[1]
        xbuf = np.empty((K * D, h)) if x is None else x
        info = np.zeros(P, np.int32) if info is None else info
        keep = [_ptr(ens.prog), _ptr(ens.prog_off), _ptr(ens.theta), _ptr(ens.theta_off), _ptr(ens.noise),
                _ptr(t, np.float64), _ptr(g, np.int32), _ptr(y1, np.float64), _ptr(y2, np.float64),
                _ptr(logw0, np.float64), _ptr(comp, np.int32), _ptr(u, np.float64), _ptr(u_res, np.float64),
                _ptr(zeta, np.float64), _ptr(xbuf), _ptr(logw), _ptr(ess), _ptr(info)]
        p = [k_[0] for k_ in keep]
        rc = self._lib.nagp_forecast_with_nowcasts(
            self._ctx, K, P, D, p[0], p[1], p[2], p[3], p[4], noise_pred, n, k, h, p[5], p[6], step, p[7],
            p[8], ya, yb, p[9], p[10], p[11], p[12], ess_thr, p[13], p[14], p[15], p[16], p[17])
        self._check(rc, raise_posdef=check)
        return xbuf.T if x is None else xbuf


    def forecast_with_nowcasts_theta(self, ens: FlatEnsemble, n, k, h, t, y1, y2, logw0, zeta, theta, noise, ya=1.0, yb=0.0,
                                     g=None, step=0.0, noise_pred=-1.0, comp=None, u=None, u_res=None, ess_thr=0.0,
                                     x=None, logw=None, ess=None, info=None, check: bool = False, K=None, D=None):
        """Per-scenario hyperparameters (`theta` [K,total], `noise` [K,P]): instances + ESS/resample + draws in one call."""
        P = ens.size
        if K is None:
            K = zeta.shape[0]
        if D is None:
            D = zeta.shape[1]
        xbuf = np.empty((K * D, h)) if x is None else x
        info = np.zeros((K, P), np.int32) if info is None else info
        keep = [_ptr(ens.prog), _ptr(ens.prog_off), _ptr(theta), _ptr(ens.theta_off), _ptr(noise),
                _ptr(t, np.float64), _ptr(g, np.int32), _ptr(y1, np.float64), _ptr(y2, np.float64),
                _ptr(logw0, np.float64), _ptr(comp, np.int32), _ptr(u, np.float64), _ptr(u_res, np.float64),
                _ptr(zeta, np.float64), _ptr(xbuf), _ptr(logw), _ptr(ess), _ptr(info)]
        p = [k_[0] for k_ in keep]
        rc = self._lib.nagp_forecast_with_nowcasts_theta(
            self._ctx, K, P, D, p[0], p[1], p[2], p[3], int(ens.theta_off[-1]), p[4], P, noise_pred, n, k, h, p[5], p[6],
            step, p[7], p[8], ya, yb, p[9], p[10], p[11], p[12], ess_thr, p[13], p[14], p[15], p[16], p[17])
        self._check(rc, raise_posdef=check)
        return xbuf.T if x is None else xbuf

    # ---- (f4) inverse transformation + per-date quantiles ---------------------------------------------
    def forecast_summary(self, x, spec=(0, 0.0, 0.0, 0.0), probs=None, want_x: bool = True):
        """x [h, N] transformed-space draws (numpy) → (inverse-transformed x [h, N] or None, quantiles [h, nq] or
        None). `spec = (kind, lambda, offset, max_value)` as carried by `transformations.InverseTransform`."""
        x = np.asarray(x, np.float64)
        h, N = x.shape
        xin = np.ascontiguousarray(x.T)                       # (N, h) C-order = column-major (h, N)
        xout = np.empty_like(xin) if want_x else None
        nq = 0 if probs is None else len(probs)
        pr = None if probs is None else np.ascontiguousarray(probs, np.float64)
        q = np.empty((h, nq)) if nq else None
        kind, lam, offset, max_value = spec
        self._check(self._lib.nagp_forecast_summary(
            self._ctx, int(kind), float(lam), float(offset), float(max_value), h, N, xin.ctypes.data,
            None if xout is None else xout.ctypes.data, nq, None if pr is None else pr.ctypes.data,
            None if q is None else q.ctypes.data))
        return (None if xout is None else xout.T), q


class Factor:
    """Device-resident factors of one base model (nagp_factor handle)."""

    def __init__(self, engine: Engine, handle, P, n, k, h, logml_n, info):
        self._engine, self._h = engine, handle
        self.P, self.n, self.k, self.h = P, n, k, h
        self.logml_n, self.info = logml_n, info

    def free(self) -> None:
        if self._h:
            self._engine._lib.nagp_factor_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
