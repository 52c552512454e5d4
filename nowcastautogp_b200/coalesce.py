"""Request coalescing for independent series fitted at the same time (BASELINE configs[3]).

The reference fits one `make_and_fit_model` per jurisdiction in a host loop (`docs/vignettes/getting-started.jl:
540-552`); inside, every SMC / MCMC / HMC step scores the P particles of THAT series (`src/make_and_fit_model.jl:
91`). P = 8…64 instances per launch leaves a 148-SM device almost idle, and the steps of one chain are sequential
by nature. The parallelism is across series: S series × P particles per step.

`CoalescingEngine` wraps one `Engine` for S concurrent clients (one host thread per series, each running the
unchanged `GPModel.fit_smc` logic). A client's `logml_batch` / `logml_grad` call blocks until every still-active
client has submitted its next request; the last one to arrive merges all pending requests with the same time grid
into ONE device call (concatenated ensembles, per-instance observation vectors) and scatters the results back.
Requests on different grids simply become separate calls. The device context is only ever entered by one thread.
Everything else (`ess`, `factor_store`, `predict`, `draw`, …) is forwarded under the same lock.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Tuple

import numpy as np

from .kernels import FlatEnsemble


def concat_ensembles(parts: List[FlatEnsemble]) -> FlatEnsemble:
    prog = np.concatenate([np.asarray(e.prog, np.uint8) for e in parts])
    theta = np.concatenate([np.asarray(e.theta, np.float64) for e in parts])
    noise = np.concatenate([np.asarray(e.noise, np.float64) for e in parts])
    prog_off, theta_off = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    po = to = 0
    for e in parts:
        prog_off.append(np.asarray(e.prog_off[1:], np.int64) + po)
        theta_off.append(np.asarray(e.theta_off[1:], np.int64) + to)
        po += int(e.prog_off[-1]); to += int(e.theta_off[-1])
    return FlatEnsemble(prog, np.concatenate(prog_off), theta, np.concatenate(theta_off), noise)


class _Request:
    __slots__ = ("kind", "key", "ens", "t", "g", "step", "y", "extra", "result", "error")

    def __init__(self, kind, ens, t, g, step, y, extra=None, extra_key=()):
        self.kind, self.ens, self.t, self.g, self.step = kind, ens, np.asarray(t, np.float64), g, float(step)
        self.y = np.asarray(y, np.float64)
        self.extra = extra
        self.key = (kind, len(self.t), self.t.tobytes(), None if g is None else np.asarray(g, np.int32).tobytes(),
                    self.step) + tuple(extra_key)
        self.result = None
        self.error = None


class CoalescingEngine:
    def __init__(self, engine, n_clients: int):
        self.engine = engine
        self._cv = threading.Condition()
        self._active = int(n_clients)
        self._pending: Dict[int, _Request] = {}
        self._generation = 0
        self.device_calls = 0          # merged launches issued
        self.requests = 0              # client requests served

    # ---- client life cycle ----------------------------------------------------------------------------
    def client(self, cid: int) -> "_Client":
        return _Client(self, cid)

    def retire(self, cid: int) -> None:
        """A client is done (or failed): stop waiting for it."""
        with self._cv:
            self._active -= 1
            if self._pending and len(self._pending) >= self._active:
                self._flush()

    # ---- coalescing -------------------------------------------------------------------------------------
    def _submit(self, cid: int, req: _Request):
        with self._cv:
            self._pending[cid] = req
            self.requests += 1
            if len(self._pending) >= self._active:
                self._flush()
            else:
                gen = self._generation
                while self._generation == gen:
                    self._cv.wait()
        if req.error is not None:
            raise req.error
        return req.result

    def _flush(self) -> None:
        """Caller holds the lock and is the only thread that touches the device."""
        groups: Dict[Tuple, List[_Request]] = {}
        for cid in sorted(self._pending):
            r = self._pending[cid]
            groups.setdefault(r.key, []).append(r)
        for reqs in groups.values():
            try:
                self._run_group(reqs)
            except Exception as e:          # every member of the group sees the failure
                for r in reqs:
                    r.error = e
        self._pending = {}
        self._generation += 1
        self._cv.notify_all()

    def _run_group(self, reqs: List[_Request]) -> None:
        r0 = reqs[0]
        n = len(r0.t)
        ens = concat_ensembles([r.ens for r in reqs])
        sizes = [r.ens.size for r in reqs]
        y = np.concatenate([np.broadcast_to(r.y, (sz, n)) for r, sz in zip(reqs, sizes)]).reshape(-1)
        self.device_calls += 1
        if r0.kind == "logml":
            lm, info = self.engine.logml_batch(ens, r0.t, y, g=r0.g, step=r0.step, y_stride=n)
            o = 0
            for r, sz in zip(reqs, sizes):
                r.result = (lm[o:o + sz].copy(), info[o:o + sz].copy()); o += sz
        elif r0.kind == "grad":
            lm, gth, gnz, info = self.engine.logml_grad(ens, r0.t, y, g=r0.g, step=r0.step, y_stride=n)
            o = to = 0
            for r, sz in zip(reqs, sizes):
                nth = int(r.ens.theta_off[-1])
                r.result = (lm[:, o:o + sz].copy(), gth[:, to:to + nth].copy(), gnz[:, o:o + sz].copy(),
                            info[:, o:o + sz].copy())
                o += sz; to += nth
        else:   # "hmc": the chains of every series side by side (K = 1), momenta / uniforms concatenated per iteration
            x0 = r0.extra
            cat = lambda name, axis: np.concatenate([r.extra[name] for r in reqs], axis=axis)
            learn = x0["noise_momenta"] is not None
            Z, NZ, lm, nacc, info = self.engine.hmc(
                ens.prog, ens.prog_off, ens.theta_off, cat("kind", 0), cat("a", 0), cat("b", 0), x0["noise_spec"],
                cat("z", 1), cat("noise_z", 1), r0.t, y, g=r0.g, step=r0.step, y_stride=n, n_leapfrog=x0["n_leapfrog"],
                eps=x0["eps"], momenta=cat("momenta", 1), noise_momenta=cat("noise_momenta", 1) if learn else None,
                log_u=cat("log_u", 1))
            o = to = 0
            for r, sz in zip(reqs, sizes):
                nth = int(r.ens.theta_off[-1])
                r.result = (Z[:, to:to + nth].copy(), NZ[:, o:o + sz].copy(), lm[:, o:o + sz].copy(),
                            nacc[:, o:o + sz].copy(), info[:, o:o + sz].copy())
                o += sz; to += nth

    def _forward(self, name, *args, **kwargs):
        with self._cv:
            return getattr(self.engine, name)(*args, **kwargs)


class _Client:
    """What a `GPModel` sees as its engine."""

    is_coalescing = True      # lockstep fits keep every step a mergeable logML request (no per-series factor store)

    def __init__(self, hub: CoalescingEngine, cid: int):
        self._hub, self._cid = hub, cid

    def logml_batch(self, ens, t, y, g=None, step: float = 0.0, **kw):
        assert not kw.get("y_stride"), "per-instance y goes through the coalescer itself"
        return self._hub._submit(self._cid, _Request("logml", ens, t, g, step, y))

    def logml_grad(self, ens, t, y1, y2=None, g=None, step: float = 0.0, theta=None, noise=None, **kw):
        if y2 is not None or theta is not None or noise is not None:      # per-scenario chains: not a fit-time call
            return self._hub._forward("logml_grad", ens, t, y1, y2=y2, g=g, step=step, theta=theta, noise=noise, **kw)
        return self._hub._submit(self._cid, _Request("grad", ens, t, g, step, y1))

    def hmc(self, prog, prog_off, theta_off, slot_kind, slot_a, slot_b, noise_spec, z, noise_z, t, y1, y2=None, g=None,
            step: float = 0.0, y_stride: int = 0, n_leapfrog: int = 10, eps: float = 0.02, momenta=None,
            noise_momenta=None, log_u=None):
        z, noise_z = np.asarray(z, np.float64), np.asarray(noise_z, np.float64)
        if y2 is not None or y_stride or z.shape[0] != 1 or log_u is None:       # per-scenario chains: not a fit-time call
            return self._hub._forward("hmc", prog, prog_off, theta_off, slot_kind, slot_a, slot_b, noise_spec, z, noise_z, t,
                                      y1, y2=y2, g=g, step=step, y_stride=y_stride, n_leapfrog=n_leapfrog, eps=eps,
                                      momenta=momenta, noise_momenta=noise_momenta, log_u=log_u)
        P = noise_z.shape[1]
        ens = FlatEnsemble(np.asarray(prog, np.uint8), np.asarray(prog_off, np.int64), np.zeros(int(theta_off[-1])),
                           np.asarray(theta_off, np.int64), np.zeros(P))
        extra = dict(kind=np.asarray(slot_kind, np.int32), a=np.asarray(slot_a, np.float64), b=np.asarray(slot_b, np.float64),
                     noise_spec=tuple(noise_spec), z=z, noise_z=noise_z, n_leapfrog=int(n_leapfrog), eps=float(eps),
                     momenta=np.asarray(momenta, np.float64),
                     noise_momenta=None if noise_momenta is None else np.asarray(noise_momenta, np.float64),
                     log_u=np.asarray(log_u, np.float64))
        key = (len(log_u), int(n_leapfrog), float(eps), tuple(noise_spec), noise_momenta is None)
        return self._hub._submit(self._cid, _Request("hmc", ens, t, g, step, y1, extra, key))

    def __getattr__(self, name):
        hub = self._hub
        attr = getattr(hub.engine, name)
        if not callable(attr):
            return attr
        return lambda *a, **k: hub._forward(name, *a, **k)
