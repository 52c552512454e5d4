"""Request coalescing for series fitted in lockstep: the merge/scatter logic against a fake engine on
the CPU, and — on the GPU — `make_and_fit_models` against one-at-a-time `make_and_fit_model`."""
import threading

import numpy as np
import pytest

from nowcastautogp_b200 import kernels as kn
from nowcastautogp_b200.coalesce import CoalescingEngine, concat_ensembles


class FakeEngine:
    """Deterministic stand-in: 'logml' of an instance = sum(theta) + noise + y[0]; records every device call."""

    def __init__(self):
        self.calls = []

    def logml_batch(self, ens, t, y, g=None, step=0.0, y_stride=0):
        B, n = ens.size, len(t)
        self.calls.append(("logml", B, n))
        y = np.asarray(y).reshape(B, n)
        lm = np.array([ens.theta[ens.theta_off[b]:ens.theta_off[b + 1]].sum() + ens.noise[b] + y[b, 0] for b in range(B)])
        return lm, np.zeros(B, np.int32)

    def logml_grad(self, ens, t, y1, g=None, step=0.0, y_stride=0):
        B, n = ens.size, len(t)
        self.calls.append(("grad", B, n))
        lm, info = self.logml_batch(ens, t, y1, y_stride=y_stride)
        self.calls.pop()
        return lm[None], np.asarray(ens.theta)[None] * 2.0, np.asarray(ens.noise)[None] * 3.0, info[None]

    def hmc(self, prog, prog_off, theta_off, kind, a, b, noise_spec, z, noise_z, t, y1, y2=None, g=None, step=0.0,
            y_stride=0, n_leapfrog=10, eps=0.02, momenta=None, noise_momenta=None, log_u=None):
        P = noise_z.shape[1]
        self.calls.append(("hmc", P, len(t)))
        y = np.asarray(y1).reshape(P, len(t))
        # "final state" = z + sum of its momenta; logml = y[0] of the chain's series + number of slots
        ns = np.diff(theta_off)
        return (z + momenta.sum(0)[None], noise_z + (0 if noise_momenta is None else noise_momenta.sum(0)[None]),
                (y[:, 0] + ns)[None], np.full((1, P), len(log_u), np.int32), np.zeros((1, P), np.int32))

    def ess(self, logw):
        return np.ones(len(logw)), None


def _ens(seed, P):
    rng = np.random.default_rng(seed)
    trees = [kn.Plus(kn.Linear(*rng.uniform(0.1, 1, 3)), kn.Periodic(*rng.uniform(0.1, 1, 3))) if rng.uniform() < 0.5
             else kn.GammaExponential(*rng.uniform(0.1, 1, 3)) for _ in range(P)]
    return kn.pack_ensemble(trees, rng.uniform(0.01, 0.1, P))


def test_concat_ensembles_offsets():
    a, b = _ens(1, 3), _ens(2, 2)
    c = concat_ensembles([a, b])
    assert c.size == 5 and c.prog_off[-1] == len(c.prog) and c.theta_off[-1] == len(c.theta)
    for i, (e, j) in enumerate([(a, 0), (a, 1), (a, 2), (b, 0), (b, 1)]):
        assert bytes(c.prog[c.prog_off[i]:c.prog_off[i + 1]]) == bytes(e.prog[e.prog_off[j]:e.prog_off[j + 1]])
        assert np.array_equal(c.theta[c.theta_off[i]:c.theta_off[i + 1]], e.theta[e.theta_off[j]:e.theta_off[j + 1]])


def test_requests_merge_per_grid_and_scatter_back():
    fake = FakeEngine()
    S = 6
    hub = CoalescingEngine(fake, S)
    t_a, t_b = np.linspace(0, 1, 10), np.linspace(0, 1, 12)
    out = [None] * S

    def work(s):
        try:
            cl = hub.client(s)
            res = []
            for rnd in range(4 if s != 2 else 2):                      # client 2 finishes early
                ens = _ens(100 * s + rnd, 2 + s % 3)
                t = t_b if (s == 5 and rnd == 1) else t_a               # one request on another grid
                y = np.full(len(t), float(s))
                if rnd % 2 == 0:
                    lm, info = cl.logml_batch(ens, t, y)
                    res.append((ens, lm))
                else:
                    lm, gth, gnz, info = cl.logml_grad(ens, t, y)
                    assert np.array_equal(gth[0], ens.theta * 2.0) and np.array_equal(gnz[0], ens.noise * 3.0)
                    res.append((ens, lm[0]))
                assert cl.ess(np.zeros((1, 2)))[0][0] == 1.0           # forwarded call
            out[s] = res
        finally:
            hub.retire(s)

    ths = [threading.Thread(target=work, args=(s,)) for s in range(S)]
    [th.start() for th in ths]
    [th.join(timeout=30) for th in ths]
    assert not any(th.is_alive() for th in ths), "coalescer deadlocked"
    for s in range(S):
        for ens, lm in out[s]:
            want = [ens.theta[ens.theta_off[b]:ens.theta_off[b + 1]].sum() + ens.noise[b] + s for b in range(ens.size)]
            assert np.allclose(lm, want, rtol=0, atol=1e-12)
    # rounds 0 and 1 have six clients (round 1 on two grids), rounds 2 and 3 five: 5 merged launches for 22 requests
    assert hub.requests == 22 and hub.device_calls == 5
    assert sorted(c[1] for c in fake.calls if c[2] == 12) == [2 + 5 % 3]


def test_error_reaches_every_member_of_the_group():
    class Boom(FakeEngine):
        def logml_batch(self, *a, **k):
            raise RuntimeError("device failure")
    hub = CoalescingEngine(Boom(), 2)
    seen = []

    def work(s):
        try:
            hub.client(s).logml_batch(_ens(s, 2), np.linspace(0, 1, 5), np.zeros(5))
        except RuntimeError as e:
            seen.append(str(e))
        finally:
            hub.retire(s)
    ths = [threading.Thread(target=work, args=(s,)) for s in range(2)]
    [th.start() for th in ths]
    [th.join(timeout=30) for th in ths]
    assert seen == ["device failure"] * 2


@pytest.mark.gpu
def test_lockstep_fit_equals_one_at_a_time(engine):
    import nowcastautogp_b200 as ng
    from nowcastautogp_b200.api import make_and_fit_models
    rng = np.random.default_rng(0)
    dates = np.arange(np.datetime64("2024-01-01"), np.datetime64("2024-02-20"))
    S = 5
    datas = [ng.TData(dates, 20 + 3 * np.sin(np.arange(len(dates)) / (3.0 + s)) + 0.3 * rng.standard_normal(len(dates)),
                      transformation=lambda v: v) for s in range(S)]
    kw = dict(n_particles=4, smc_data_proportion=0.25, n_mcmc=3, n_hmc=2)
    models = make_and_fit_models(datas, rng=np.random.default_rng(7), engine=engine, **kw)
    st = make_and_fit_models.last_stats
    assert st["device_calls"] * 3 < st["requests"], st               # coalescing happened
    # the same generators and observation order, one series at a time
    seq = np.random.SeedSequence(np.random.default_rng(7).integers(2 ** 63))
    rngs = [np.random.default_rng(ss) for ss in seq.spawn(S + 1)]
    order = rngs[S].permutation(len(dates))
    for s in range(S):
        solo = ng.make_and_fit_model(datas[s], rng=rngs[s], engine=engine, obs_order=order, **kw)
        assert [p.prog for p in solo.particles] == [p.prog for p in models[s].particles]
        assert all(np.array_equal(a.z, b.z) for a, b in zip(solo.particles, models[s].particles))
        assert np.array_equal(solo.log_weights, models[s].log_weights)
    x = ng.forecast(models[0], dates[-1] + np.arange(1, 5), 10)
    assert x.shape == (4, 10) and np.isfinite(x).all()


@pytest.mark.gpu
def test_long_series_are_fitted_one_at_a_time_with_the_appendable_store(engine):
    """n > 232: `make_and_fit_models` runs the series one by one on the plain engine, so the un-rejuvenated schedule
    steps extend an appendable factor store instead of re-factoring; the result is what `make_and_fit_model` gives with
    the same generators and observation order."""
    import nowcastautogp_b200 as ng
    from nowcastautogp_b200.api import make_and_fit_models
    rng = np.random.default_rng(3)
    n, S = 260, 2
    dates = np.datetime64("2019-01-06") + 7 * np.arange(n)
    datas = [ng.TData(dates, 30 + 4 * np.sin(np.arange(n) / (5.0 + s)) + 0.02 * np.arange(n) + 0.4 * rng.standard_normal(n),
                      transformation=lambda v: v) for s in range(S)]
    kw = dict(n_particles=3, smc_data_proportion=0.25, n_mcmc=0, n_hmc=0)
    models = make_and_fit_models(datas, rng=np.random.default_rng(11), engine=engine, **kw)
    assert make_and_fit_models.last_stats.get("one_by_one")
    assert all(m.append_steps > 0 for m in models)                      # the store served schedule steps
    seq = np.random.SeedSequence(np.random.default_rng(11).integers(2 ** 63))
    rngs = [np.random.default_rng(ss) for ss in seq.spawn(S + 1)]
    order = rngs[S].permutation(n)
    for s in range(S):
        solo = ng.make_and_fit_model(datas[s], rng=rngs[s], engine=engine, obs_order=order, **kw)
        assert [p.prog for p in solo.particles] == [p.prog for p in models[s].particles]
        assert np.array_equal(solo.log_weights, models[s].log_weights)


def test_hmc_requests_merge():
    fake = FakeEngine()
    S = 4
    hub = CoalescingEngine(fake, S)
    t = np.linspace(0, 1, 9)
    out = [None] * S

    def work(s):
        try:
            ens = _ens(7 * s, 2 + s)
            total, P = int(ens.theta_off[-1]), ens.size
            rng = np.random.default_rng(s)
            z, nz = rng.standard_normal((1, total)), rng.standard_normal((1, P))
            mom, mnz, lu = rng.standard_normal((3, total)), rng.standard_normal((3, P)), rng.standard_normal((3, P))
            Z, NZ, lm, nacc, info = hub.client(s).hmc(ens.prog, ens.prog_off, ens.theta_off, np.zeros(total, np.int32),
                                                      np.zeros(total), np.ones(total), (0, -1.5, 1.0), z, nz, t,
                                                      np.full(len(t), float(s)), momenta=mom, noise_momenta=mnz, log_u=lu)
            out[s] = (np.allclose(Z, z + mom.sum(0)), np.allclose(NZ, nz + mnz.sum(0)),
                      np.allclose(lm[0], s + np.diff(ens.theta_off)), (nacc == 3).all(), Z.shape == z.shape)
        finally:
            hub.retire(s)
    ths = [threading.Thread(target=work, args=(s,)) for s in range(S)]
    [th.start() for th in ths]
    [th.join(timeout=30) for th in ths]
    assert all(o is not None and all(o) for o in out), out
    assert fake.calls == [("hmc", 2 + 3 + 4 + 5, 9)]


@pytest.mark.gpu
def test_sharded_fit_single_rank_roundtrips_models(engine):
    """make_and_fit_models_sharded without a process group (world = 1): the models come back through their
    `to_dict()` payloads and forecast as usual."""
    import nowcastautogp_b200 as ng
    rng = np.random.default_rng(4)
    dates = np.arange(np.datetime64("2024-01-01"), np.datetime64("2024-02-10"))
    datas = [ng.TData(dates, 15 + np.cos(np.arange(len(dates)) / (2.0 + s)) + 0.2 * rng.standard_normal(len(dates)),
                      transformation=lambda v: v) for s in range(3)]
    models = ng.make_and_fit_models_sharded(datas, seed=3, engine=engine, n_particles=3, smc_data_proportion=0.5,
                                            n_mcmc=2, n_hmc=1)
    assert len(models) == 3 and all(m.n_obs == len(dates) for m in models)
    again = ng.make_and_fit_models_sharded(datas, seed=3, engine=engine, n_particles=3, smc_data_proportion=0.5,
                                           n_mcmc=2, n_hmc=1)
    for a, b in zip(models, again):                      # seeded: reproducible
        assert [p.prog for p in a.particles] == [p.prog for p in b.particles]
        assert np.array_equal(a.log_weights, b.log_weights)
    x = ng.forecast(models[2], dates[-1] + np.arange(1, 4), 8)
    assert x.shape == (3, 8) and np.isfinite(x).all()
