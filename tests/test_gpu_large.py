"""GPU parity of the large path (232 < q <= 4096: factor in HBM) and the in-place rank-append, through
the C ABI, against the CPU oracle on the same seeded inputs. Tolerance 1e-9 relative (north_star)."""
import numpy as np
import pytest

from nowcastautogp_b200 import kernels as kn
from nowcastautogp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("use_grid", [False, True])
@pytest.mark.parametrize("n,P", [(233, 5), (256, 3), (300, 6), (512, 4), (631, 2), (759, 2), (761, 2), (900, 2)])   # from 96 tile rows: two rows per warp
def test_large_logml_matches_oracle(engine, oracle, n, P, use_grid):
    w = syn.make_workload(n, 0, 0, 1, P, seed=900 + n)
    g = w.g[:n] if use_grid else None
    got, info = engine.logml_batch(w.ens, w.t[:n], w.y1, g=g, step=w.step)
    want, winfo = oracle.logml_batch(w.ens, w.t[:n], w.y1, g=g, step=w.step)
    assert (info == 0).all() and (winfo == 0).all()
    assert rel(got, want) < RTOL


def test_large_many_instances_dynamic_queue(engine, oracle):
    """More instances than resident CTAs (persistent grid + instance queue), per-instance y."""
    n, P = 240, 700
    w = syn.make_workload(n, 0, 0, 1, 7, seed=11)
    trees = [w.trees[i % 7] for i in range(P)]
    noise = np.array([w.noise[i % 7] for i in range(P)])
    ens = kn.pack_ensemble(trees, noise)
    got, info = engine.logml_batch(ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    want, _ = oracle.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    assert (info == 0).all()
    assert rel(got, np.tile(want, P // 7)) < RTOL


@pytest.mark.parametrize("n", [400, 800])      # one / two rows per warp below the diagonal block
@pytest.mark.parametrize("use_grid", [False, True])
def test_large_forecast_instances_match_oracle(engine, oracle, use_grid, n):
    k, h, P, K = 2, 6, 3, 2
    w = syn.make_workload(n, k, h, K, P, seed=41)
    g = w.g if use_grid else None
    r = engine.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=g, step=w.step)
    assert (r["info"] == 0).all()
    for s in range(K):
        y = np.concatenate([w.y1, w.y2[s]])
        for p, tr in enumerate(w.trees):
            prog, th = kn.flatten(tr)
            o = oracle.instance_joint(prog, th, w.noise[p], n, k, h, w.t, y, w.ya, w.yb, g=g, step=w.step)
            assert o["info"] == 0
            assert abs(r["logw"][s, p] - (w.logw0[p] + o["logml_m"] - o["logml_n"])) < RTOL * max(1.0, abs(o["logml_m"]))
            assert rel(r["mu"][s, p], o["mu"]) < RTOL
            assert rel(r["L"][s, p], o["L"]) < 1e-8      # h x h factor of a Schur complement: cond-limited


def test_large_fast_path_factor_append_predict(engine, oracle):
    """forecast_with_nowcasts' default schedule on a long series: factor once, append K scenarios."""
    n, k, h, P, K, D = 300, 1, 5, 4, 6, 3
    w = syn.make_workload(n, k, h, K, P, seed=5)
    rng = np.random.default_rng(3)
    zeta = rng.standard_normal((K, D, h))
    u = rng.uniform(size=(K, D))
    logw = np.empty((K, P))
    x = engine.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, zeta, w.ya, w.yb, g=w.g,
                                      step=w.step, u=u, logw=logw)
    want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
    xo, _, _ = oracle.draws(want["logw"], want["mu"], want["L"], zeta, u=u)
    assert rel(logw, want["logw"]) < RTOL
    assert rel(x, xo) < 1e-8


@pytest.fixture(params=[False, True], ids=["by-rows", "streamed"])
def append_path(request, monkeypatch):
    """Appends of two or more tile rows go through the factorisation kernel restricted to the new rows; with
    NAGP_APPEND_STREAM set (read per call) they take the row-streaming kernel like the small ones."""
    if request.param:
        monkeypatch.setenv("NAGP_APPEND_STREAM", "1")
    else:
        monkeypatch.delenv("NAGP_APPEND_STREAM", raising=False)
    return request.param


@pytest.mark.parametrize("use_grid", [False, True])
@pytest.mark.parametrize("n0,adds", [(260, [1, 7, 8, 70]), (64, [3, 200]), (250, [1, 1, 1]), (10, [300]), (100, [29, 33, 64])])
def test_rank_append_matches_full_refactor(engine, oracle, n0, adds, use_grid, append_path):
    """SMC data annealing: logML after each in-place append == logML of a from-scratch factorisation of
    the grown series (oracle), and dlogml is the increment."""
    P = 4
    ntot = n0 + sum(adds)
    w = syn.make_workload(ntot, 0, 0, 1, P, seed=77 + n0)
    g = w.g if use_grid else None
    f = engine.factor_store_large(w.ens, w.t[:n0], w.y1[:n0], capacity=ntot, g=None if g is None else g[:n0], step=w.step)
    want0, _ = oracle.logml_batch(w.ens, w.t[:n0], w.y1[:n0], g=None if g is None else g[:n0], step=w.step)
    assert rel(f.logml_n, want0) < RTOL
    cur, prev = n0, want0
    for kx in adds:
        dl, lm, info = engine.factor_append(f, w.t[cur:cur + kx], w.y1[cur:cur + kx],
                                            g_new=None if g is None else g[cur:cur + kx])
        cur += kx
        want, _ = oracle.logml_batch(w.ens, w.t[:cur], w.y1[:cur], g=None if g is None else g[:cur], step=w.step)
        assert (info == 0).all() and f.n == cur
        assert rel(lm, want) < RTOL
        assert np.abs(dl - (want - prev)).max() < 1e-9 * np.abs(want).max()
        prev = want
    with pytest.raises(Exception):
        engine.factor_append(f, w.t[:1], w.y1[:1], g_new=None if g is None else g[:1])   # capacity exceeded
    f.free()


def test_daily_series_2048_append_consistency(engine, oracle):
    """BASELINE config 5 shape (n = 2048 daily points): one-shot factorisation == 1024 + append(1024),
    and both match the oracle for one particle (the oracle needs seconds per instance at this size)."""
    n, P = 2048, 3
    w = syn.make_workload(n, 0, 0, 1, P, seed=2048, period=365.0)
    full, info = engine.logml_batch(w.ens, w.t, w.y1, g=w.g, step=w.step)
    assert (info == 0).all()
    f = engine.factor_store_large(w.ens, w.t[:1024], w.y1[:1024], capacity=n, g=w.g[:1024], step=w.step)
    dl, lm, info = engine.factor_append(f, w.t[1024:], w.y1[1024:], g_new=w.g[1024:])
    assert (info == 0).all()
    assert rel(lm, full) < RTOL
    one = kn.pack_ensemble(w.trees[:1], w.noise[:1])
    want, _ = oracle.logml_batch(one, w.t, w.y1, g=w.g, step=w.step)
    assert abs(full[0] - want[0]) < RTOL * abs(want[0])
    f.free()


def test_large_not_positive_definite_reports_info(engine):
    """A Constant kernel with zero noise and zero jitter is rank one: pivot 2 fails (LAPACK-style info)."""
    n = 300
    ens = kn.pack_ensemble([kn.Constant(1.0)], [0.0])
    engine.set_jitter(0.0)
    try:
        t = np.linspace(0, 1, n)
        _, info = engine.logml_batch(ens, t, np.zeros(n))
        assert info[0] > 0 and info[0] <= 8
    finally:
        engine.set_jitter(1e-5)


def test_fit_smc_unrejuvenated_steps_use_rank_append(engine, oracle):
    """`fit_smc!` with `n_mcmc = 0` (/root/reference/src/make_and_fit_model.jl:88-91: linear_schedule + fit_smc!): after the
    first schedule step every step is a rank-append on the stored factors (observations in shuffled arrival order, lag
    grid of the whole series). The log-weight trajectory must follow the oracle's from-scratch logML of the same
    observations at every step."""
    import nowcastautogp_b200 as ng
    from nowcastautogp_b200.api import linear_schedule
    from nowcastautogp_b200.gpmodel import GPModel, pack_particles
    n, P = 96, 6
    rng = np.random.default_rng(17)
    ds = np.datetime64("2023-01-01") + 7 * np.arange(n)
    y = np.log(50.0) + np.sin(2 * np.pi * np.arange(n) / 52.0) + 0.15 * rng.standard_normal(n)
    m = GPModel(ds, y, n_particles=P, rng=np.random.default_rng(3), engine=engine)
    sched = linear_schedule(n, 0.125)
    traj = []
    orig = m.maybe_resample

    def spy(thr):                       # record the weights before any resampling resets them; never resample
        traj.append((m.n_obs, m._logml.copy()))
        return False
    m.maybe_resample = spy
    m.fit_smc(schedule=sched, n_mcmc=0, n_hmc=0)
    m.maybe_resample = orig
    assert m.append_steps == len(sched) - 1 and [s for s, _ in traj] == sched
    ens = pack_particles(m.particles, m.config)
    t_all, g_all, step_all = m._times(m.ds)
    ys = m.y_transform.apply(m.y)
    for step, lm in traj:
        idx = m.obs_order[:step]
        want, winfo = oracle.logml_batch(ens, t_all[idx], ys[idx], g=g_all[idx], step=step_all)
        ok = winfo == 0
        assert ok.any() and np.array_equal(np.isfinite(lm), ok)
        assert rel(lm[ok], want[ok]) < 1e-9, step
    # and the cumulative log-weights are the last logML (weights start at 0, nothing resampled)
    assert rel(m.log_weights[ok], want[ok]) < 1e-9
