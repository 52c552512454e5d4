"""GPU parity of nagp_logml_grad: d logML / d theta and d logML / d noise against Richardson-
extrapolated central differences of the CPU oracle's __float128 log marginal likelihood (no analytic gradient
code on the oracle side: the check is independent of the device's reverse-mode formulas)."""
import numpy as np
import pytest

from nowcastautogp_b200 import kernels as kn
from nowcastautogp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["tile", "column"], autouse=True)
def grad_variant(engine, request):
    """0: tile kernel (in-place K^-1 on the tensor pipe, lag-binned reverse mode); 1: the first column kernel,
    kept as the fallback for unsorted grids and n beyond the shared-memory-resident size."""
    engine.set_variant(request.param)
    yield request.param
    engine.set_variant(0)


def fd_grad(oracle_q, prog, theta, noise, t, y, g, step):
    """4th-order central differences in every theta slot and the noise (steps exactly representable)."""
    theta = np.asarray(theta, float)

    def f(th, nz):
        lm, info = oracle_q.logml(prog, th, nz, t, y, g=g, step=step)
        assert info == 0
        return lm

    def d1(eval_at, x0):
        out = []
        for h in (4e-4, 2e-4):
            hp = h * max(1.0, abs(x0))
            xp, xm = x0 + hp, x0 - hp
            out.append((eval_at(xp) - eval_at(xm)) / (xp - xm))
        return (4.0 * out[1] - out[0]) / 3.0

    gth = np.empty(len(theta))
    for j in range(len(theta)):
        def at(x, j=j):
            th = theta.copy(); th[j] = x
            return f(th, noise)
        gth[j] = d1(at, theta[j])
    gnz = d1(lambda x: f(theta, x), noise)
    return gth, gnz


ALL_NODES = [
    kn.Plus(kn.Times(kn.Linear(0.3, 0.2, 0.7), kn.Periodic(0.9, 0.25, 0.8)), kn.GammaExponential(0.4, 1.3, 0.6)),
    kn.ChangePoint(kn.SquaredExponential(0.3, 0.9), kn.Plus(kn.Constant(0.4), kn.GammaExponential(0.2, 0.8, 0.5)), 0.45, 0.05),
    kn.Periodic(0.7, 0.33, 1.1),
    kn.Times(kn.Constant(0.6), kn.Linear(-0.1, 0.5, 1.2)),
]


@pytest.mark.parametrize("use_grid", [False, True])
@pytest.mark.parametrize("n", [12, 37, 64])
def test_grad_every_node_type(engine, oracle_q, n, use_grid):
    w = syn.make_workload(n, 0, 0, 1, 2, seed=n)
    noise = np.array([0.05, 0.2, 0.1, 0.02])
    ens = kn.pack_ensemble(ALL_NODES, noise)
    g = w.g[:n] if use_grid else None
    lm, gth, gnz, info = engine.logml_grad(ens, w.t[:n], w.y1, g=g, step=w.step)
    assert (info == 0).all()
    for p, tr in enumerate(ALL_NODES):
        prog, th = kn.flatten(tr)
        want_lm, _ = oracle_q.logml(prog, th, noise[p], w.t[:n], w.y1, g=g, step=w.step)
        assert abs(lm[0, p] - want_lm) < 1e-9 * abs(want_lm)
        wth, wnz = fd_grad(oracle_q, prog, th, noise[p], w.t[:n], w.y1, g, w.step)
        got = gth[0, ens.theta_off[p]:ens.theta_off[p + 1]]
        names = kn.theta_slot_names(prog)
        scale = max(np.abs(wth).max(), abs(wnz), 1.0)
        for j, nm in enumerate(names):
            assert abs(got[j] - wth[j]) < 2e-6 * scale, (p, nm, got[j], wth[j])
        assert abs(gnz[0, p] - wnz) < 2e-6 * scale


# One tree per route through the tile kernel's differentiation pass (csrc/nagp_grad_tile.cu): leaves under the root's
# Plus nodes are summarised by lag sums / weighted moments, what is left is swept in registers (<= 9 compiled ops) or
# by the local-memory interpreter.
L1, L2, L3 = kn.Linear(0.3, 0.2, 0.7), kn.Linear(-0.2, 0.1, 0.4), kn.Linear(0.6, 0.3, 0.2)
SE, PER, GE = kn.SquaredExponential(0.3, 0.9), kn.Periodic(0.9, 0.25, 0.8), kn.GammaExponential(0.4, 1.3, 0.6)
ROUTES = [
    kn.Plus(SE, kn.Plus(L1, kn.Plus(kn.Constant(0.4), PER))),                              # nothing left to interpret
    kn.Plus(L1, kn.Times(L2, SE)),                                                         # moments + a three-op rest
    kn.Plus(kn.Times(L1, SE), kn.Plus(kn.Constant(0.3), kn.ChangePoint(GE, L2, 0.45, 0.05))),   # two terms re-joined
    kn.Times(L1, kn.Times(SE, kn.Times(L2, PER))),                                         # seven ops in registers
    kn.Plus(PER, kn.Times(L1, kn.Times(L2, kn.Times(L3, kn.Times(L1, kn.Times(L2, SE)))))),     # too many leaves: interpreter
]


@pytest.mark.parametrize("use_grid", [False, True])
def test_grad_every_route_of_the_differentiation_pass(engine, oracle_q, use_grid):
    n = 37
    w = syn.make_workload(n, 0, 0, 1, 2, seed=5)
    noise = np.array([0.05, 0.2, 0.1, 0.02, 0.08])
    ens = kn.pack_ensemble(ROUTES, noise)
    g = w.g[:n] if use_grid else None
    lm, gth, gnz, info = engine.logml_grad(ens, w.t[:n], w.y1, g=g, step=w.step)
    assert (info == 0).all()
    for p, tr in enumerate(ROUTES):
        prog, th = kn.flatten(tr)
        wth, wnz = fd_grad(oracle_q, prog, th, noise[p], w.t[:n], w.y1, g, w.step)
        got = gth[0, ens.theta_off[p]:ens.theta_off[p + 1]]
        scale = max(np.abs(wth).max(), abs(wnz), 1.0)
        assert np.abs(got - wth).max() < 2e-6 * scale, (p, got, wth)
        assert abs(gnz[0, p] - wnz) < 2e-6 * scale


def test_grad_on_a_grid_with_gaps(engine, oracle_q):
    """Missing observations: grid indices with holes (partner rows come from the inverse index map, and the moments of
    the additive Linear / Constant leaves still use t_b = t_a - lag * step)."""
    w = syn.make_workload(60, 0, 0, 1, 2, seed=21)
    keep = np.sort(np.random.default_rng(3).choice(60, size=41, replace=False))
    t, g, y = w.t[keep], w.g[keep], w.y1[keep]
    trees = [ROUTES[0], ROUTES[1], ROUTES[2], ALL_NODES[1]]
    noise = np.array([0.05, 0.2, 0.1, 0.07])
    ens = kn.pack_ensemble(trees, noise)
    lm, gth, gnz, info = engine.logml_grad(ens, t, y, g=g, step=w.step)
    assert (info == 0).all()
    for p, tr in enumerate(trees):
        prog, th = kn.flatten(tr)
        wth, wnz = fd_grad(oracle_q, prog, th, noise[p], t, y, g, w.step)
        got = gth[0, ens.theta_off[p]:ens.theta_off[p + 1]]
        scale = max(np.abs(wth).max(), abs(wnz), 1.0)
        assert np.abs(got - wth).max() < 2e-6 * scale, (p, got, wth)
        assert abs(gnz[0, p] - wnz) < 2e-6 * scale


def test_grad_tile_kernel_agrees_with_column_kernel_at_its_largest_size(engine, grad_variant):
    """m = 216 (27 tile rows: the last size the tile kernel takes) and m = 209 (ragged last tile): both kernels
    differentiate the same logML."""
    if grad_variant != 0:
        pytest.skip("cross-check runs once")
    for m in (216, 209):
        w = syn.make_workload(m, 0, 0, 1, 4, seed=m)
        engine.set_variant(0)
        lm0, g0, n0, i0 = engine.logml_grad(w.ens, w.t[:m], w.y1, g=w.g[:m], step=w.step)
        engine.set_variant(1)
        lm1, g1, n1, i1 = engine.logml_grad(w.ens, w.t[:m], w.y1, g=w.g[:m], step=w.step)
        engine.set_variant(0)
        assert (i0 == 0).all() and (i1 == 0).all()
        scale = max(np.abs(g1).max(), np.abs(n1).max(), 1.0)
        assert np.abs(g0 - g1).max() < 1e-7 * scale and np.abs(n0 - n1).max() < 1e-7 * scale
        assert np.abs(lm0 - lm1).max() < 1e-9 * np.abs(lm1).max()


def test_grad_prior_sampled_trees_per_scenario(engine, oracle_q):
    """K scenarios with their own hyperparameters and nowcast values: the per-scenario HMC batch."""
    n, k, P, K = 40, 2, 5, 3
    w = syn.make_workload(n, k, 0, K, P, seed=9)
    theta_k, noise_k = syn.perturbed_theta(w.ens, K, seed=4)
    lm, gth, gnz, info = engine.logml_grad(w.ens, w.t[:n + k], w.y1, y2=w.y2, g=w.g[:n + k], step=w.step,
                                           theta=theta_k, noise=noise_k)
    assert (info == 0).all() and lm.shape == (K, P) and gth.shape == (K, w.ens.theta_off[-1])
    for s in (0, K - 1):
        y = np.concatenate([w.y1, w.y2[s]])
        for p, tr in enumerate(w.trees):
            prog, _ = kn.flatten(tr)
            th = theta_k[s, w.ens.theta_off[p]:w.ens.theta_off[p + 1]]
            wth, wnz = fd_grad(oracle_q, prog, th, noise_k[s, p], w.t[:n + k], y, w.g[:n + k], w.step)
            got = gth[s, w.ens.theta_off[p]:w.ens.theta_off[p + 1]]
            scale = max(np.abs(wth).max(), abs(wnz), 1.0)
            assert np.abs(got - wth).max() < 5e-6 * scale, (s, p)
            assert abs(gnz[s, p] - wnz) < 5e-6 * scale


def test_grad_vignette_size_is_consistent(engine):
    """n = 150 (BASELINE configs[1]): the gradient is the derivative of the device's own logML — a directional
    finite difference of nagp_logml_batch along a random direction agrees to 1e-5 relative."""
    n, P = 150, 8
    w = syn.make_workload(n, 0, 0, 1, P, seed=77)
    lm, gth, gnz, info = engine.logml_grad(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    assert (info == 0).all()
    rng = np.random.default_rng(1)
    names = kn.theta_slot_names(w.ens.prog.tobytes())
    d = rng.standard_normal(len(w.ens.theta)) * np.abs(w.ens.theta)
    d[[nm in ("scale", "gamma") for nm in names]] = 0.0          # keep gamma <= 2 and the fixed sharpness untouched
    h = 1e-5
    ens_p = kn.FlatEnsemble(w.ens.prog, w.ens.prog_off, w.ens.theta + h * d, w.ens.theta_off, w.ens.noise)
    ens_m = kn.FlatEnsemble(w.ens.prog, w.ens.prog_off, w.ens.theta - h * d, w.ens.theta_off, w.ens.noise)
    lp, _ = engine.logml_batch(ens_p, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    lmn, _ = engine.logml_batch(ens_m, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    fd = (lp - lmn) / (2 * h)
    an = np.array([gth[0, w.ens.theta_off[p]:w.ens.theta_off[p + 1]] @ d[w.ens.theta_off[p]:w.ens.theta_off[p + 1]]
                   for p in range(P)])
    assert np.abs(fd - an).max() < 1e-5 * max(np.abs(an).max(), 1.0)


def test_hmc_parameter_move(engine):
    """mcmc_parameters! as gradient HMC on every particle at once: accepts, stays finite, conserves the
    Hamiltonian well enough for a healthy acceptance rate, and improves a deliberately poor start."""
    import nowcastautogp_b200 as ng
    rng = np.random.default_rng(3)
    dates = np.arange(np.datetime64("2024-01-01"), np.datetime64("2024-03-01"))
    vals = 10 + 2 * np.sin(np.arange(len(dates)) / 4.0) + 0.1 * rng.standard_normal(len(dates))
    m = ng.GPModel(dates, vals, n_particles=6, rng=np.random.default_rng(5), engine=engine)
    m.n_obs = len(vals)
    m._logml = m.logml(m.particles, m._obs_idx())
    before = m._logml.copy()
    rate = m.mcmc_parameters(10)
    assert 0.3 < rate <= 1.0
    after = m.logml(m.particles, m._obs_idx())
    assert np.isfinite(after).all() and np.allclose(after, m._logml, rtol=1e-9, atol=1e-9)
    prior = lambda ps: np.array([-0.5 * (p.z @ p.z + p.noise_z ** 2) for p in ps])
    assert np.median(after) > np.median(before)          # prior draws are far from the posterior mode


def test_device_hmc_walks_the_host_chain(engine):
    """`nagp_hmc` (leapfrog, z -> theta maps and accept/reject on the device, the iteration replayed as a CUDA
    graph) against the host integrator calling `nagp_logml_grad` per stage, fed the same momenta and uniforms:
    same accept decisions, same final states to 1e-8 (the two differ only in libm ulps of the z -> theta maps)."""
    import nowcastautogp_b200 as ng
    rng = np.random.default_rng(11)
    dates = np.arange(np.datetime64("2024-01-01"), np.datetime64("2024-03-15"))
    vals = 10 + 2 * np.sin(np.arange(len(dates)) / 5.0) + 0.1 * rng.standard_normal(len(dates))
    for cfg in (ng.GPConfig(), ng.GPConfig(noise=0.05)):
        base = ng.GPModel(dates, vals, n_particles=7, config=cfg, rng=np.random.default_rng(5), engine=engine)
        base.n_obs = len(vals)
        base._logml = base.logml(base.particles, base._obs_idx())
        d = base.to_dict()
        out = []
        for device in (True, False):
            m = ng.GPModel.from_dict(d, engine=engine, rng=np.random.default_rng(99))
            rate = m.mcmc_parameters(6, {"device": device, "eps": 0.03})
            out.append((rate, m))
        (r_dev, m_dev), (r_host, m_host) = out
        assert r_dev == r_host and r_dev > 0.2
        for a, b in zip(m_dev.particles, m_host.particles):
            assert a.prog == b.prog
            np.testing.assert_allclose(a.z, b.z, rtol=1e-8, atol=1e-8)
            assert abs(a.noise_z - b.noise_z) < 1e-8
        np.testing.assert_allclose(m_dev._logml, m_host._logml, rtol=1e-8)
        again = m_dev.logml(m_dev.particles, m_dev._obs_idx())
        np.testing.assert_allclose(again, m_dev._logml, rtol=1e-9)


def test_per_scenario_device_hmc_in_forecast_with_nowcasts(engine):
    """forecast_with_nowcasts(n_hmc > 0): K x P chains in one nagp_hmc call; same draws as the host integrator."""
    import nowcastautogp_b200 as ng
    from nowcastautogp_b200 import api
    rng = np.random.default_rng(3)
    dates = np.arange(np.datetime64("2024-01-01"), np.datetime64("2024-02-20"))
    vals = 30 + 4 * np.sin(np.arange(len(dates)) / 6.0) + 0.3 * rng.standard_normal(len(dates))
    data = ng.TData(dates, vals, transformation=lambda v: v)
    model = ng.make_and_fit_model(data, n_particles=4, smc_data_proportion=0.5, n_mcmc=2, n_hmc=1,
                                  rng=np.random.default_rng(1), engine=engine)
    nds = dates[-1] + np.arange(1, 3)
    nowcasts = ng.create_nowcast_data([[vals[-1] * 1.01, vals[-1] * 1.02], [vals[-1] * 0.97, vals[-1] * 0.99],
                                       [vals[-1], vals[-1] * 1.05]], list(nds))
    fdates = nds[-1] + np.arange(1, 5)
    xs = []
    for device in (True, False):
        api.HMC_DEFAULT["device"] = device
        try:
            xs.append(ng.forecast_with_nowcasts(model, nowcasts, fdates, 6, n_hmc=3, rng=np.random.default_rng(8)))
        finally:
            api.HMC_DEFAULT.pop("device", None)
    assert xs[0].shape == (4, 18) and np.isfinite(xs[0]).all()
    np.testing.assert_allclose(xs[0], xs[1], rtol=1e-6, atol=1e-6)


def test_gradient_by_differences_for_long_series(engine):
    """Beyond the gradient kernels' size the host takes central differences of the device logML (one batched call through
    the large-path kernel). At a size both routes take, the two gradients agree; at n = 300 the HMC move runs on it."""
    from nowcastautogp_b200.gpmodel import GPModel
    n = 120
    w = syn.make_workload(n, 0, 0, 1, 6, seed=31)
    dates = np.datetime64("2020-01-05") + 7 * np.arange(n)
    m = GPModel(dates, w.y1, n_particles=6, rng=np.random.default_rng(5), engine=engine)
    m.fit_smc(schedule=[n], n_mcmc=0, n_hmc=0, shuffle=False)
    idx = m._obs_idx()
    lp0, dz0, dn0 = m._logpost_grad(m.particles, idx)
    m._force_fd_gradient = True
    lp1, dz1, dn1 = m._logpost_grad(m.particles, idx)
    m._force_fd_gradient = False
    ok = np.isfinite(lp0)
    assert ok.any() and np.allclose(lp0[ok], lp1[ok], rtol=1e-10)
    for a_, b_, good in zip(dz0, dz1, ok):
        if good:
            assert np.abs(a_ - b_).max() < 1e-4 * max(1.0, np.abs(a_).max())
    assert np.abs(dn0[ok] - dn1[ok]).max() < 1e-4 * max(1.0, np.abs(dn0[ok]).max())
    # a long series: the move runs, keeps every particle finite and accepts something
    n = 300
    t = np.arange(n)
    y = np.sin(2 * np.pi * t / 52) + 0.01 * t + 0.1 * np.random.default_rng(2).standard_normal(n)
    dates = np.datetime64("2016-01-03") + 7 * np.arange(n)
    m = GPModel(dates, y, n_particles=4, rng=np.random.default_rng(6), engine=engine)
    m.fit_smc(schedule=[n], n_mcmc=0, n_hmc=0, shuffle=False)
    before = np.array(m._logml, copy=True)
    rate = m.mcmc_parameters(3)
    assert 0.0 <= rate <= 1.0 and np.isfinite(np.asarray(m._logml)[np.isfinite(before)]).all()
    assert rate > 0.0
