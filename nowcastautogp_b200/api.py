"""`make_and_fit_model`, `forecast`, `forecast_with_nowcasts` — the reference's public entry points,
same names, keyword arguments, assertion behaviour and return shapes.

Mirrors `/root/reference/src/make_and_fit_model.jl:17-27,78-93` and
`/root/reference/src/forecasting.jl:29-75,117-167`. Where the reference spawns one task per nowcast
scenario and rebuilds the particle ensemble inside each (`forecasting.jl:131-133`), this module
issues ONE batched device call over all (scenario, particle) instances; kernel-structure proposals and the
HMC integrator stay on the CPU, every likelihood, gradient, factorisation, moment and draw is in libnagp.
Random numbers come from a NumPy `Generator` (the reference uses Julia's task-local Xoshiro), so
outputs agree with the reference in distribution, not draw by draw.
"""
from __future__ import annotations

import warnings
from typing import Callable, List, Optional, Sequence

import numpy as np

from .gpmodel import (GPConfig, GPModel, HMC_DEFAULT, device_noise_spec, device_slot_spec, dtheta_dz_slots,
                      pack_particles, slot_codes, transform_slots)
from .tdata import TData
from . import kernels as kn


def _identity(y):
    return y


def _stabilize_for_fit(y: np.ndarray, *, flat_threshold: float = 1e-3, rng=None) -> np.ndarray:
    """Jitter guard for near-constant series (`src/make_and_fit_model.jl:17-27`)."""
    y = np.asarray(y, np.float64)
    n = len(y)
    if n <= 1:
        return y
    scale = abs(y.sum() / n) + 1
    rel_range = (y.max() - y.min()) / scale
    if rel_range >= flat_threshold:
        return y
    sigma = flat_threshold * scale
    warnings.warn(f"Near-constant series (relative range {rel_range} < {flat_threshold}); adding jitter "
                  f"(sigma = {sigma}) so the GP covariance stays positive-definite (issue #51).")
    rng = np.random.default_rng() if rng is None else rng
    return y + sigma * rng.standard_normal(n)


def linear_schedule(n: int, proportion: float) -> List[int]:
    """`AutoGP.Schedule.linear_schedule` [R]: cumulative counts step, 2·step, …, with n appended."""
    step = max(1, int(round(proportion * n)))
    sched = list(range(step, n + 1, step))
    if not sched or sched[-1] != n:
        sched.append(n)
    return sched


def make_and_fit_model(data: TData, *, n_particles: int = 1, smc_data_proportion: float = 0.1,
                       flat_threshold: float = 1e-3, config: Optional[GPConfig] = None, rng=None,
                       engine=None, **kwargs) -> GPModel:
    """`src/make_and_fit_model.jl:78-93`. `kwargs` go to `fit_smc` (`n_mcmc`, `n_hmc` required)."""
    if "n_mcmc" not in kwargs or "n_hmc" not in kwargs:
        raise TypeError("make_and_fit_model: keyword arguments n_mcmc and n_hmc are required")  # UndefKeywordError
    rng = np.random.default_rng() if rng is None else rng
    n_train = len(data.y)
    y_fit = _stabilize_for_fit(data.y, flat_threshold=flat_threshold, rng=rng)
    model = GPModel(data.ds, y_fit, n_particles=n_particles, config=GPConfig() if config is None else config,
                    rng=rng, engine=engine)
    effective_proportion = max(smc_data_proportion, 1.0 / n_train)
    model.fit_smc(schedule=linear_schedule(n_train, effective_proportion), **kwargs)
    return model


def make_and_fit_models(datas: Sequence[TData], *, n_particles: int = 1, smc_data_proportion: float = 0.1,
                        flat_threshold: float = 1e-3, config: Optional[GPConfig] = None, rng=None, engine=None,
                        share_order: bool = True, rngs: Optional[Sequence] = None, **kwargs) -> List[GPModel]:
    """`make_and_fit_model` for S independent series at once (the per-jurisdiction loop of
    `docs/vignettes/getting-started.jl:540-552`): one host thread per series runs the unchanged SMC logic, and a
    `CoalescingEngine` merges the S concurrent likelihood / gradient requests of every step into one device
    launch of S·P instances (`coalesce.py`). Series with identical dates share one shuffled observation order
    (`share_order`) so that their requests stay on the same time grid. The result of each series is what
    `make_and_fit_model(data, rng=<its spawned generator>, obs_order=<the shared order>)` gives on its own."""
    import threading
    from .coalesce import CoalescingEngine
    from .gpmodel import default_engine
    if "n_mcmc" not in kwargs or "n_hmc" not in kwargs:
        raise TypeError("make_and_fit_models: keyword arguments n_mcmc and n_hmc are required")
    S = len(datas)
    if S == 0:
        return []
    eng = default_engine() if engine is None else engine
    seq = np.random.SeedSequence(rng.integers(2 ** 63) if rng is not None else None)
    spawned = [np.random.default_rng(ss) for ss in seq.spawn(S + 1)]
    if rngs is not None:        # explicit per-series generators (the sharded fit: results independent of the rank count)
        assert len(rngs) == S, "rngs must hold one generator per series"
        spawned[:S] = list(rngs)
    rngs = spawned
    same_dates = all(len(d.ds) == len(datas[0].ds) and np.array_equal(np.asarray(d.ds), np.asarray(datas[0].ds))
                     for d in datas)
    order = rngs[S].permutation(len(datas[0].y)) if (share_order and same_dates and kwargs.get("shuffle", True)) else None
    if max(len(d.y) for d in datas) > 232:
        # Long series (the large path, n > 232): one series' particles already fill the device, and a fit on its own engine
        # keeps an appendable factor store between un-rejuvenated schedule steps (block appends are ~4x cheaper than
        # re-factoring) — which a lockstep fit, whose every step must stay a mergeable request, cannot.
        out = []
        for s in range(S):
            kw = dict(kwargs)
            if order is not None:
                kw["obs_order"] = order
            out.append(make_and_fit_model(datas[s], n_particles=n_particles, smc_data_proportion=smc_data_proportion,
                                          flat_threshold=flat_threshold, config=config, rng=rngs[s], engine=eng, **kw))
        make_and_fit_models.last_stats = {"requests": 0, "device_calls": 0, "one_by_one": True}
        return out
    hub = CoalescingEngine(eng, S)
    models: List[Optional[GPModel]] = [None] * S
    errors: List[Optional[BaseException]] = [None] * S

    def work(s: int) -> None:
        try:
            kw = dict(kwargs)
            if order is not None:
                kw["obs_order"] = order
            models[s] = make_and_fit_model(datas[s], n_particles=n_particles, smc_data_proportion=smc_data_proportion,
                                           flat_threshold=flat_threshold, config=config, rng=rngs[s],
                                           engine=hub.client(s), **kw)
        except BaseException as e:      # noqa: BLE001 - re-raised on the calling thread
            errors[s] = e
        finally:
            hub.retire(s)

    threads = [threading.Thread(target=work, args=(s,), name=f"nagp-fit-{s}") for s in range(S)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for e in errors:
        if e is not None:
            raise e
    for m in models:
        m.engine = eng
    make_and_fit_models.last_stats = {"requests": hub.requests, "device_calls": hub.device_calls}
    return models


def make_and_fit_models_sharded(datas: Sequence[TData], *, seed: int = 0, engine=None, group=None, **kwargs) -> List[GPModel]:
    """`make_and_fit_models` with the series split over the ranks of the current `torch.distributed` group (one
    process per GPU): every rank fits its contiguous share in lockstep on its own device and the fitted models
    travel as `to_dict()` payloads in one `all_gather_object`. Every rank returns all S models (bound to its own
    engine), ready for `forecast_with_nowcasts_sharded`. Series s uses a generator derived from `(seed, s)`, so
    the result does not depend on the number of ranks unless series share an observation order (`share_order`
    couples the series of one rank; pass `share_order=False` for world-size-independent fits)."""
    from .gpmodel import default_engine
    from .sharding import sharded_fit
    eng = default_engine() if engine is None else engine

    def fit_local(idx: List[int]) -> List[dict]:
        # one generator per SERIES, derived from (seed, series index): independent of how the series fall on the ranks
        models = make_and_fit_models([datas[s] for s in idx], rngs=[np.random.default_rng([int(seed), int(s)]) for s in idx],
                                     rng=np.random.default_rng([int(seed), 10 ** 9 + int(idx[0])]), engine=eng, **kwargs)
        return [m.to_dict() for m in models]

    dicts = sharded_fit(fit_local, len(datas), group=group)
    return [GPModel.from_dict(d, engine=eng, rng=np.random.default_rng([int(seed), 10 ** 6 + s]))
            for s, d in enumerate(dicts)]


def _apply(inv_transformation: Callable, x: np.ndarray, engine=None) -> np.ndarray:
    if inv_transformation is _identity:
        return x
    spec = getattr(inv_transformation, "spec", None)
    if spec is not None and engine is not None and x.ndim == 2 and x.size:
        # built-in inverse of get_transformations: one device pass over the whole draw matrix
        return engine.forecast_summary(x, spec)[0]
    try:
        out = inv_transformation(x)                 # vectorised closures (np.exp, scaled logistic …)
        if isinstance(out, np.ndarray) and out.shape == x.shape:
            return out
    except Exception:
        pass
    return np.vectorize(inv_transformation, otypes=[np.float64])(x)


def forecast(model: GPModel, forecast_dates, forecast_draws: int, *, inv_transformation: Callable = _identity,
             forecast_n_hmc: Optional[int] = None) -> np.ndarray:
    """`src/forecasting.jl:29-75`: `(len(forecast_dates), forecast_draws)` samples."""
    dates = np.asarray(list(forecast_dates) if not isinstance(forecast_dates, np.ndarray) else forecast_dates)
    if forecast_n_hmc is None:
        x = model.predict_mvn(dates).rand(forecast_draws, rng=model.rng)
    else:
        x = np.empty((len(dates), forecast_draws))
        for i in range(forecast_draws):
            model.mcmc_parameters(forecast_n_hmc)
            x[:, i] = model.predict_mvn(dates).rand(rng=model.rng)
    return _apply(inv_transformation, np.ascontiguousarray(x), model._engine())


class _ScenarioParams:
    """Per-(scenario, particle) unconstrained hyperparameters of K copies of one particle set, moved
    by batched Metropolis steps: one `nagp_forecast_instances` call scores all K·P proposals."""

    def __init__(self, model: GPModel, K: int):
        self.m = model
        self.K, self.P = K, model.num_particles()
        self.prog_all = b"".join(p.prog for p in model.particles)
        self.codes = np.concatenate([slot_codes(p.prog) for p in model.particles])
        self.codes_k = np.tile(self.codes, K)
        self.z = np.tile(np.concatenate([p.z for p in model.particles]), (K, 1))
        self.noise_z = np.tile(np.array([p.noise_z for p in model.particles]), (K, 1))
        self.ens = pack_particles(model.particles, model.config)
        # particle owning each theta slot (for the per-particle prior term)
        self.owner = np.repeat(np.arange(self.P), np.diff(self.ens.theta_off))

    def theta(self, z, noise_z):
        cfg = self.m.config
        th = transform_slots(self.codes_k, z.reshape(-1), cfg).reshape(z.shape)
        if cfg.noise is not None:
            nz = np.full(noise_z.shape, float(cfg.noise))
        else:
            nz = transform_slots(np.zeros(noise_z.size, np.int8), noise_z.reshape(-1), cfg).reshape(noise_z.shape)
        return np.ascontiguousarray(th), np.ascontiguousarray(nz)

    def jacobian(self, z, th, noise_z, nz):
        """d theta / d z per slot [K, total] and d noise / d noise_z [K, P]."""
        cfg = self.m.config
        jz = dtheta_dz_slots(self.codes_k, z.reshape(-1), th.reshape(-1), cfg).reshape(z.shape)
        if cfg.noise is not None:
            jn = np.zeros_like(noise_z)
        else:
            jn = dtheta_dz_slots(np.zeros(noise_z.size, np.int8), noise_z.reshape(-1), nz.reshape(-1), cfg).reshape(noise_z.shape)
        return jz, jn

    def log_prior(self, z, noise_z):
        lp = np.zeros((self.K, self.P))
        np.add.at(lp.T, self.owner, -0.5 * (z * z).T)
        return lp - 0.5 * noise_z * noise_z


def forecast_with_nowcasts(base_model: GPModel, nowcasts: Sequence[TData], forecast_dates,
                           forecast_draws_per_nowcast: int, *, inv_transformation: Callable = _identity,
                           n_mcmc: int = 0, n_hmc: int = 0, ess_threshold: float = 0.0,
                           forecast_n_hmc: Optional[int] = None, verbose: bool = False, rng=None) -> np.ndarray:
    """`src/forecasting.jl:117-167`: matrix `(len(forecast_dates), len(nowcasts)·D)`, scenario-major
    column blocks. The base model is not mutated."""
    x, _ = _forecast_with_nowcasts(base_model, nowcasts, forecast_dates, forecast_draws_per_nowcast,
                                   n_mcmc=n_mcmc, n_hmc=n_hmc, ess_threshold=ess_threshold,
                                   forecast_n_hmc=forecast_n_hmc, rng=rng)
    return _apply(inv_transformation, x, base_model._engine())


def forecast_with_nowcasts_sharded(base_models: Sequence[GPModel], nowcasts: Sequence[Sequence[TData]],
                                   forecast_dates, forecast_draws_per_nowcast: int, *,
                                   inv_transformation: Callable = _identity, n_mcmc: int = 0, n_hmc: int = 0,
                                   ess_threshold: float = 0.0, forecast_n_hmc: Optional[int] = None,
                                   group=None, device=None, seed: Optional[int] = None):
    """One `forecast_with_nowcasts` per series (the per-jurisdiction loop of
    `docs/vignettes/getting-started.jl:540-552`), with the (series, scenario) pairs partitioned over
    the ranks of the current `torch.distributed` group (one process per GPU) and one final all-gather.
    Returns `(draws, logw)`: per series the `(h, K_s·D)` matrix and the `[K_s, P]` log-weights, on
    every rank. All models must have the same number of particles. Every (series, first-scenario) slice draws from
    its own generator, derived from `seed` (or from the series' model generator when `seed` is None) and the slice,
    so the scenarios of one series computed on different ranks use independent normals and uniforms."""
    from .sharding import sharded_forecast
    D = int(forecast_draws_per_nowcast)
    P = base_models[0].num_particles()
    assert all(m.num_particles() == P for m in base_models), "all series must use the same n_particles"
    h = len(list(forecast_dates))

    def slice_rng(sl):
        # One generator per (series, first scenario) slice. Models rebuilt from the same dicts carry identically seeded
        # generators on every rank, so drawing from `base_models[s].rng` would give every rank the same normals and
        # uniforms: the column blocks of one series computed on different ranks would share their noise.
        if seed is not None:
            base = int(seed)
        else:
            st = base_models[sl.series].rng.bit_generator.state.get("state", {})
            base = int(st.get("state", 0)) % (2 ** 63) if isinstance(st, dict) else 0
        return np.random.default_rng([base, int(sl.series), int(sl.k0)])

    def compute(sl):
        return _forecast_with_nowcasts(base_models[sl.series], nowcasts[sl.series][sl.k0:sl.k1], forecast_dates, D,
                                       n_mcmc=n_mcmc, n_hmc=n_hmc, ess_threshold=ess_threshold,
                                       forecast_n_hmc=forecast_n_hmc, rng=slice_rng(sl))

    # per-scenario refinement makes every (series, scenario) pair its own piece of work: cut the pair list evenly;
    # without it the particles of a series are factored once for all its scenarios, so series stay whole
    draws, logw = sharded_forecast(compute, len(base_models), [len(nc) for nc in nowcasts], h, D, P,
                                   group=group, device=device, split_series=(n_hmc > 0 or n_mcmc > 0))
    return {s: _apply(inv_transformation, np.ascontiguousarray(x), base_models[s]._engine())
            for s, x in draws.items()}, logw


def _forecast_with_nowcasts(base_model: GPModel, nowcasts: Sequence[TData], forecast_dates,
                            forecast_draws_per_nowcast: int, *, n_mcmc: int = 0, n_hmc: int = 0,
                            ess_threshold: float = 0.0, forecast_n_hmc: Optional[int] = None, rng=None):
    """Body of `forecast_with_nowcasts` before the inverse transformation: `(x [h, K·D], logw [K, P])`."""
    assert len(nowcasts) > 0, "nowcasts vector must not be empty"
    assert not (n_mcmc > 0 and n_hmc == 0), "If n_mcmc > 0, n_hmc must also be > 0 for MCMC refinement"
    assert 0.0 <= ess_threshold <= 1.0, "ess_threshold must be between 0 and 1"
    assert forecast_n_hmc is None or forecast_n_hmc > 0, "forecast_n_hmc must be > 0 if specified"
    rng = base_model.rng if rng is None else rng
    D = int(forecast_draws_per_nowcast)
    dates = np.asarray(list(forecast_dates) if not isinstance(forecast_dates, np.ndarray) else forecast_dates)
    base_dict = base_model.to_dict()                                   # forecasting.jl:128
    ds0 = np.asarray(nowcasts[0].ds)
    # do all scenarios share the nowcast dates (create_nowcast_data.jl builds them that way)? One vectorised comparison:
    # a Python-level array_equal per scenario costs more than the device work of the default schedule
    shared_ds = all(len(nc.ds) == len(ds0) for nc in nowcasts)
    if shared_ds and any(nc.ds is not nowcasts[0].ds for nc in nowcasts):
        shared_ds = bool((np.stack([np.asarray(nc.ds) for nc in nowcasts]) == ds0).all())

    def scenario_loop(ncs):
        # the reference's schedule verbatim, one model copy per scenario (forecasting.jl:133-155); every likelihood is
        # still a device call
        blocks, lws = [], []
        for nc in ncs:
            m_ = GPModel.from_dict(base_dict, engine=base_model.engine, rng=rng)
            m_.add_data(nc.ds, nc.y)
            m_.maybe_resample(ess_threshold * m_.num_particles())
            if n_mcmc > 0 and n_hmc > 0:
                m_.mcmc_structure(n_mcmc, n_hmc)
            elif n_hmc > 0:
                m_.mcmc_parameters(n_hmc)
            blocks.append(forecast(m_, dates, D, forecast_n_hmc=forecast_n_hmc))
            lws.append(np.asarray(m_.log_weights, np.float64).copy())
        return np.hstack(blocks), np.stack(lws)

    if n_mcmc > 0 or not shared_ds:
        # structure moves make the programs diverge per scenario
        return scenario_loop(nowcasts)

    # the batched schedules without rejuvenation only read the model: no private copy needed (the reference copies per
    # scenario because add_data! mutates; here the appended observations never enter the model object)
    read_only = n_hmc == 0 and forecast_n_hmc is None
    m = base_model if read_only else GPModel.from_dict(base_dict, engine=base_model.engine, rng=rng)
    alive = np.isfinite(m.log_weights)
    if not alive.all() and alive.any():
        # particles that left the fit with weight -inf (Gram not positive definite) carry no mass: the batched paths run
        # on the others and report -inf for these
        pruned = base_model.to_dict()
        keep = np.nonzero(alive)[0]
        pruned["particles"] = [pruned["particles"][i] for i in keep]
        pruned["log_weights"] = [pruned["log_weights"][i] for i in keep]
        pruned["logml"] = [pruned["logml"][i] for i in keep]
        x_, lw_ = _forecast_with_nowcasts(GPModel.from_dict(pruned, engine=base_model.engine, rng=rng), nowcasts, forecast_dates,
                                          forecast_draws_per_nowcast, n_mcmc=n_mcmc, n_hmc=n_hmc, ess_threshold=ess_threshold,
                                          forecast_n_hmc=forecast_n_hmc, rng=rng)
        lw_full = np.full((len(nowcasts), len(alive)), -np.inf)
        lw_full[:, keep] = lw_
        return x_, lw_full
    eng = m._engine()
    K, P, n, k, h = len(nowcasts), m.num_particles(), len(m.y), len(ds0), len(dates)
    idx = m._obs_idx()
    t, g, step = m._times(np.concatenate([m.ds[idx], ds0.astype(m.ds.dtype), dates.astype(m.ds.dtype)]))
    yt = m.y_transform
    y1 = yt.apply(m.y[idx])
    y2 = np.ascontiguousarray(yt.apply(np.stack([np.asarray(nc.y, np.float64) for nc in nowcasts])))
    ens = m.ensemble()

    if n_hmc == 0 and forecast_n_hmc is None and k > 16:
        # more nowcast points than the scenario-append kernel keeps in registers: one fused factorisation per
        # (scenario, particle) with the shared hyperparameters, then the same ESS/resample/draw kernel
        zeta = rng.standard_normal((K, D, h))
        u = rng.uniform(size=(K, D))
        u_res = rng.uniform(size=(K, P)) if ess_threshold > 0.0 else None
        r = eng.forecast_instances(ens, n, k, h, t, y1, y2, m.log_weights, yt.slope, yt.intercept, g=g, step=step,
                                   check=True)
        x, _, _ = eng.draw(r["logw"], r["mu"], r["L"], zeta, u=u, u_res=u_res, ess_thr=ess_threshold)
        return np.ascontiguousarray(x), r["logw"]

    if n_hmc == 0 and forecast_n_hmc is None:
        # default path: factor once per particle, append K scenarios, ESS/resample, draw — one call
        zeta = rng.standard_normal((K, D, h))
        u = rng.uniform(size=(K, D))
        u_res = rng.uniform(size=(K, P)) if ess_threshold > 0.0 else None
        logw = np.empty((K, P))
        x = eng.forecast_with_nowcasts(ens, n, k, h, t, y1, y2, m.log_weights, zeta, yt.slope, yt.intercept,
                                       g=g, step=step, u=u, u_res=u_res, ess_thr=ess_threshold, logw=logw)
        return np.ascontiguousarray(x), logw

    # per-scenario parameter rejuvenation: all K·P chains advance together
    sp = _ScenarioParams(m, K)
    th, nz = sp.theta(sp.z, sp.noise_z)

    def score(th_, nz_, moments):
        lm_m = np.empty((K, P))
        r = eng.forecast_instances(ens, n, k, h, t, y1, y2, m.log_weights, yt.slope, yt.intercept, g=g,
                                   step=step, theta=th_, noise=nz_, K=K, logml_m=lm_m, want_moments=moments)
        lm_m = np.where(r["info"] == 0, lm_m, -np.inf)
        return r, lm_m

    r, lm = score(th, nz, False)
    logw = r["logw"].copy()                                            # add_data!: forecasting.jl:135
    if not np.all(np.isfinite(lm)):
        from .engine import PosDefError
        raise PosDefError(1)
    # maybe_resample! per scenario (forecasting.jl:138-141). Resampling swaps whole particles — structure, z and noise
    # together. The batch shares one program per particle index across scenarios, so it can express the swap only when
    # every particle has the same program (then a parent's z is still read under its own structure). With
    # heterogeneous programs the scenarios that resample go through the reference's one-model-per-scenario schedule
    # instead (same semantics, one device call per likelihood), the others stay in the batch.
    ess, w = eng.ess(logw)
    need = np.nonzero(ess < ess_threshold * P)[0]
    same_program = all(p_.prog == m.particles[0].prog for p_ in m.particles)
    if len(need) and not same_program:
        slow = set(int(i) for i in need)
        fast = [i for i in range(K) if i not in slow]
        x_out, lw_out = np.empty((h, K * D)), np.empty((K, P))
        xs, ls = scenario_loop([nowcasts[i] for i in need])
        for j, i in enumerate(need):
            x_out[:, i * D:(i + 1) * D] = xs[:, j * D:(j + 1) * D]
            lw_out[i] = ls[j]
        if fast:
            # the ESS depends on the data and the particles only: none of these triggers again
            xf, lf = _forecast_with_nowcasts(base_model, [nowcasts[i] for i in fast], forecast_dates, D, n_mcmc=n_mcmc,
                                             n_hmc=n_hmc, ess_threshold=ess_threshold, forecast_n_hmc=forecast_n_hmc, rng=rng)
            for j, i in enumerate(fast):
                x_out[:, i * D:(i + 1) * D] = xf[:, j * D:(j + 1) * D]
                lw_out[i] = lf[j]
        return x_out, lw_out
    off = ens.theta_off
    for s in need:
        parents = rng.choice(P, size=P, p=w[s])
        sp.z[s] = np.concatenate([sp.z[s, off[a]:off[a + 1]] for a in parents])
        sp.noise_z[s] = sp.noise_z[s, parents]
        lm[s] = lm[s, parents]
        logw[s] = 0.0

    def metropolis(n_steps, step_size=0.15):
        nonlocal lm
        for _ in range(n_steps):
            zp = sp.z + step_size * rng.standard_normal(sp.z.shape)
            nzp = sp.noise_z + (step_size * rng.standard_normal(sp.noise_z.shape) if m.config.noise is None else 0.0)
            thp, nzv = sp.theta(zp, nzp)
            _, lmp = score(thp, nzv, False)
            acc = np.log(rng.uniform(size=(K, P))) < (lmp - lm) + sp.log_prior(zp, nzp) - sp.log_prior(sp.z, sp.noise_z)
            slot_acc = acc[:, sp.owner]
            sp.z = np.where(slot_acc, zp, sp.z)
            sp.noise_z = np.where(acc, nzp, sp.noise_z)
            lm = np.where(acc, lmp, lm)

    learn_noise = m.config.noise is None

    def logpost_grad(z, noise_z):
        """Per (scenario, particle): logML over [train | nowcast] + log N(z; 0, I), and d/dz — one device call."""
        th_, nz_ = sp.theta(z, noise_z)
        lmg, gth, gnz, info = eng.logml_grad(ens, t[:n + k], y1, y2=y2 if k else None, g=None if g is None else g[:n + k],
                                             step=step, theta=th_, noise=nz_, K=K)
        jz, jn = sp.jacobian(z, th_, noise_z, nz_)
        lml = np.where(info == 0, lmg, -np.inf)
        dz = np.where(np.isfinite(gth), gth, 0.0) * jz - z
        dn = (np.where(np.isfinite(gnz), gnz, 0.0) * jn - noise_z) if learn_noise else np.zeros_like(noise_z)
        return lml, lml + sp.log_prior(z, noise_z if learn_noise else np.zeros_like(noise_z)), dz, dn

    def hmc(n_steps):
        """`n_steps` HMC steps on all K x P chains at once (mcmc_parameters! on every scenario's model copy,
        forecasting.jl:148 and :65), integrator on the device: one `nagp_hmc` call."""
        nonlocal lm
        from .engine import NagpError
        L, eps = int(HMC_DEFAULT["n_leapfrog"]), float(HMC_DEFAULT["eps"])
        if not HMC_DEFAULT.get("device", True):
            return hmc_host(n_steps)
        total = sp.z.shape[1]
        mom, mnz, logu = np.empty((n_steps, K, total)), np.empty((n_steps, K, P)), np.empty((n_steps, K, P))
        for it in range(n_steps):
            mom[it] = rng.standard_normal((K, total))
            if learn_noise:
                mnz[it] = rng.standard_normal((K, P))
            logu[it] = np.log(rng.uniform(size=(K, P)))
        kind, sa, sb = device_slot_spec(sp.codes, m.config)
        try:
            sp.z, sp.noise_z, lml, _, _ = eng.hmc(
                ens.prog, ens.prog_off, ens.theta_off, kind, sa, sb, device_noise_spec(m.config), sp.z, sp.noise_z,
                t[:n + k], y1, y2=y2 if k else None, g=None if g is None else g[:n + k], step=step, n_leapfrog=L, eps=eps,
                momenta=mom, noise_momenta=mnz if learn_noise else None, log_u=logu)
        except NagpError as e:
            if e.code != -4:
                raise
            return metropolis(n_steps)
        lm = lml

    def hmc_host(n_steps):
        """The same chains with the integrator on the host: each leapfrog stage is one `nagp_logml_grad` call
        (cross-check of `nagp_hmc`; `HMC_DEFAULT["device"] = False`)."""
        nonlocal lm
        from .engine import NagpError
        L, eps = int(HMC_DEFAULT["n_leapfrog"]), float(HMC_DEFAULT["eps"])
        try:
            lml, lp, dz, dn = logpost_grad(sp.z, sp.noise_z)
        except NagpError as e:
            if e.code != -4:
                raise
            return metropolis(n_steps)
        for _ in range(n_steps):
            pz = rng.standard_normal(sp.z.shape)
            pn = rng.standard_normal(sp.noise_z.shape) if learn_noise else np.zeros_like(sp.noise_z)
            kin0 = np.zeros((K, P)); np.add.at(kin0.T, sp.owner, 0.5 * (pz * pz).T); kin0 += 0.5 * pn * pn
            zq, nq, gz, gn, lp_q, lml_q = sp.z.copy(), sp.noise_z.copy(), dz, dn, lp, lml
            for _l in range(L):
                pz = pz + 0.5 * eps * gz; pn = pn + 0.5 * eps * gn
                zq = zq + eps * pz; nq = nq + eps * pn
                lml_q, lp_q, gz, gn = logpost_grad(zq, nq)
                pz = pz + 0.5 * eps * gz; pn = pn + 0.5 * eps * gn
            kin1 = np.zeros((K, P)); np.add.at(kin1.T, sp.owner, 0.5 * (pz * pz).T); kin1 += 0.5 * pn * pn
            with np.errstate(invalid="ignore"):
                dH = (lp_q - kin1) - (lp - kin0)
            acc = np.log(rng.uniform(size=(K, P))) < np.where(np.isfinite(dH), dH, -np.inf)
            slot_acc = acc[:, sp.owner]
            sp.z = np.where(slot_acc, zq, sp.z); sp.noise_z = np.where(acc, nq, sp.noise_z)
            dz = np.where(slot_acc, gz, dz); dn = np.where(acc, gn, dn)
            lp = np.where(acc, lp_q, lp); lml = np.where(acc, lml_q, lml)
        lm = lml

    if n_hmc > 0:
        hmc(n_hmc)                                              # mcmc_parameters!: forecasting.jl:148

    def draw_block(Dn):
        # predict_mvn + rand for every scenario's rejuvenated particles (forecasting.jl:152-155, :66-67). The mixture
        # weights are the log-weights fixed at add_data! (rejuvenation leaves them untouched [R]); moments and draws stay on
        # the device: nagp_forecast_with_nowcasts_theta would also recompute the weights, so the two-call form is kept here.
        th_, nz_ = sp.theta(sp.z, sp.noise_z)
        r_, _ = score(th_, nz_, True)
        zeta = rng.standard_normal((K, Dn, h))
        u = rng.uniform(size=(K, Dn))
        x_, _, _ = eng.draw(logw, r_["mu"], r_["L"], zeta, u=u)
        return np.ascontiguousarray(x_)                                # [h, K*Dn]

    if forecast_n_hmc is None:
        x = draw_block(D)
    else:
        x = np.empty((h, K * D))
        for i in range(D):                                             # forecasting.jl:63-68
            hmc(forecast_n_hmc)
            x[:, i::D] = draw_block(1)
    return x, logw
