"""Edge cases of the C ABI on the GPU: size boundaries between the shared-memory and HBM paths, degenerate
shapes (one point, no nowcast, no forecast dates, one particle), the limits of the scenario-append kernel,
argument errors, and a property check at BASELINE.json's full sizes where the oracle is too slow."""
import ctypes as C

import numpy as np
import pytest

import nowcastautogp_b200 as ng
from nowcastautogp_b200 import kernels as kn
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import NagpError

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 16, 231, 232, 233, 234, 240])
def test_logml_size_boundaries(engine, oracle, n):
    """n = 232 is the last size of the shared-memory tile kernel; 233+ runs with the factor in HBM."""
    w = syn.make_workload(max(n, 3), 0, 0, 1, 3, seed=300 + n)
    t, y, g = w.t[:n], w.y1[:n], w.g[:n]
    got, info = engine.logml_batch(w.ens, t, y, g=g, step=w.step)
    want, _ = oracle.logml_batch(w.ens, t, y, g=g, step=w.step)
    assert (info == 0).all()
    assert rel(got, want) < RTOL


@pytest.mark.parametrize("n,k,h", [(20, 0, 0), (20, 3, 0), (20, 0, 5), (1, 1, 1), (220, 2, 10), (225, 1, 9)])
def test_degenerate_point_layouts(engine, oracle, n, k, h):
    """No nowcast points, no forecast dates, both, the smallest problem, and q straddling the 232 boundary."""
    P, K = 2, 2
    w = syn.make_workload(max(n, 3), k, h, K, P, seed=17 + n + k + h)
    if n < 3:
        w = syn.Workload(n, k, h, w.t[3 - n:], w.g[3 - n:] - w.g[3 - n], w.step, w.y1[3 - n:], w.y2, w.ya, w.yb, w.trees,
                         w.noise, w.logw0, w.ens)
    r = engine.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
    assert (r["info"] == 0).all()
    for s in range(K):
        y = np.concatenate([w.y1, w.y2[s]]) if k else w.y1
        for p, tr in enumerate(w.trees):
            prog, th = kn.flatten(tr)
            o = oracle.instance_joint(prog, th, w.noise[p], n, k, h, w.t, y, w.ya, w.yb, g=w.g, step=w.step)
            assert abs(r["logw"][s, p] - (w.logw0[p] + o["logml_m"] - o["logml_n"])) < RTOL * max(1.0, abs(o["logml_m"]))
            if h:
                assert rel(r["mu"][s, p], o["mu"]) < RTOL and rel(r["L"][s, p], o["L"]) < 1e-8


def test_many_nowcast_points(engine, oracle):
    """k = 16 is the most the scenario-append kernel keeps in registers; the fused entry point refuses k = 17
    with NAGP_E_SIZE and the host API routes such calls through the general path instead."""
    n, h, P, K, D = 40, 3, 2, 3, 2
    rng = np.random.default_rng(0)
    for k in (16, 17):
        w = syn.make_workload(n, k, h, K, P, seed=k)
        zeta, u = rng.standard_normal((K, D, h)), rng.uniform(size=(K, D))
        want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
        if k == 16:
            logw = np.empty((K, P))
            engine.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, zeta, w.ya, w.yb, g=w.g,
                                          step=w.step, u=u, logw=logw)
            assert rel(logw, want["logw"]) < RTOL
        else:
            with pytest.raises(NagpError) as ei:
                engine.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, zeta, w.ya, w.yb, g=w.g,
                                              step=w.step, u=u)
            assert ei.value.code == -4
            got = engine.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
            assert rel(got["logw"], want["logw"]) < RTOL


def test_host_api_many_nowcast_points(engine):
    dates = np.arange(np.datetime64("2024-01-01"), np.datetime64("2024-03-01"))
    vals = 10 + np.sin(np.arange(len(dates)) / 5.0) + 0.1 * np.random.default_rng(1).standard_normal(len(dates))
    data = ng.create_transformed_data(dates, vals, transformation=lambda v: v)
    m = ng.make_and_fit_model(data, n_particles=2, n_mcmc=3, n_hmc=2, rng=np.random.default_rng(2), engine=engine)
    nd = np.arange(np.datetime64("2024-03-01"), np.datetime64("2024-03-19"))       # 18 nowcast points
    sc = ng.create_nowcast_data(np.tile(vals[-18:, None], (1, 3)) + 0.01, nd)
    fd = np.arange(np.datetime64("2024-03-19"), np.datetime64("2024-03-23"))
    r = ng.forecast_with_nowcasts(m, sc, fd, 4)
    assert r.shape == (4, 12) and np.isfinite(r).all()


def test_argument_errors(engine):
    w = syn.make_workload(20, 1, 2, 1, 2, seed=1)
    with pytest.raises(NagpError):                       # too large for any path
        engine.logml_batch(w.ens, np.linspace(0, 1, 5000), np.zeros(5000))
    lib, ctx = engine._lib, engine._ctx
    assert lib.nagp_logml_batch(ctx, 0, None, None, None, None, None, 5, None, None, 0.0, None, 0, None, None) == -1
    assert b"null" in lib.nagp_last_error(ctx)
    assert lib.nagp_set_variant(ctx, 9) == -1 and lib.nagp_set_jitter(ctx, -1.0) == -1
    bad = C.c_void_p()
    assert lib.nagp_init(10_000, C.byref(bad)) == -1 and not bad.value


def test_full_size_properties(engine):
    """BASELINE configs[1] at full size (32 particles x 1000 scenarios, n=150, h=9) — too big for the oracle in a
    test, so size-independent properties: (1) the fast path and the general path agree on every (scenario,
    particle) log-weight; (2) draws are linear in the supplied normals: x(zeta) - x(0) = L_c zeta, so doubling
    zeta doubles the deviation bit for bit; (3) scenario s of the batch equals the same scenario run alone."""
    n, k, h, P, K, D = 150, 1, 9, 32, 1000, 20
    w = syn.make_workload(n, k, h, K, P, seed=20261018 + 2)
    rng = np.random.default_rng(5)
    zeta, u = rng.standard_normal((K, D, h)), rng.uniform(size=(K, D))
    lw_fast = np.empty((K, P))
    x1 = engine.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, zeta, w.ya, w.yb, g=w.g, step=w.step,
                                       u=u, logw=lw_fast)
    gen = engine.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
    assert (gen["info"] == 0).all()
    assert rel(lw_fast, gen["logw"]) < RTOL
    x0 = engine.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, np.zeros_like(zeta), w.ya, w.yb,
                                       g=w.g, step=w.step, u=u)
    x2 = engine.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, 2.0 * zeta, w.ya, w.yb, g=w.g,
                                       step=w.step, u=u)
    assert x1.shape == (h, K * D) and np.isfinite(x1).all()
    assert np.allclose(x2 - x0, 2.0 * (x1 - x0), rtol=1e-12, atol=1e-12 * np.abs(x1).max())
    s = 517
    xs = engine.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2[s:s + 1], w.logw0, zeta[s:s + 1], w.ya, w.yb,
                                       g=w.g, step=w.step, u=u[s:s + 1])
    assert np.array_equal(np.asarray(xs), np.asarray(x1)[:, s * D:(s + 1) * D])
