"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): log marginal likelihoods and posterior moments within 1e-9
relative in FP64; draws bit-identical given identical (mu, L, component, normals).
"""
import numpy as np
import pytest

from nowcastautogp_b200 import kernels as kn
from nowcastautogp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(params=[1, 2, 3, 4], ids=["column-kernel", "tile-kernel", "slot-kernel", "large-kernel"])
def veng(engine, request):
    """The engine pinned to one factorisation kernel (1 = shared-memory column, 2 = DMMA tile, 3 = slot kernel
    wherever it applies: lag-grid times and 17 <= q <= 168; the tile kernel elsewhere, 4 = the factor-in-HBM kernel of the
    large path at every size)."""
    engine.set_variant(request.param)
    yield engine
    engine.set_variant(0)


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _oracle_joint_all(oracle, w, K, use_grid):
    P = w.ens.size
    out = dict(logw=np.empty((K, P)), mu=np.empty((K, P, w.h)), L=np.empty((K, P, w.h, w.h)))
    for s in range(K):
        y = np.concatenate([w.y1, w.y2[s]])
        for p, tr in enumerate(w.trees):
            prog, th = kn.flatten(tr)
            r = oracle.instance_joint(prog, th, w.noise[p], w.n, w.k, w.h, w.t, y, w.ya, w.yb,
                                      g=w.g if use_grid else None, step=w.step)
            assert r["info"] == 0
            out["logw"][s, p] = w.logw0[p] + r["logml_m"] - r["logml_n"]
            out["mu"][s, p] = r["mu"]
            out["L"][s, p] = r["L"]
    return out


@pytest.mark.parametrize("use_grid", [False, True])
@pytest.mark.parametrize("n,k,h,P", [(30, 2, 5, 6), (150, 1, 9, 8), (123, 1, 4, 5), (10, 2, 10, 3)])
def test_logml_batch_matches_oracle(veng, oracle, n, k, h, P, use_grid):
    w = syn.make_workload(n, k, h, 2, P, seed=100 + n + P)
    g = w.g[:n] if use_grid else None
    got, info = veng.logml_batch(w.ens, w.t[:n], w.y1, g=g, step=w.step)
    want, winfo = oracle.logml_batch(w.ens, w.t[:n], w.y1, g=g, step=w.step)
    assert (info == 0).all() and (winfo == 0).all()
    assert rel(got, want) < RTOL


@pytest.mark.parametrize("use_grid", [False, True])
@pytest.mark.parametrize("n,k,h,P,K", [(30, 2, 5, 4, 3), (150, 1, 9, 6, 2), (40, 0, 6, 3, 1)])
def test_forecast_instances_match_oracle(veng, oracle, n, k, h, P, K, use_grid):
    w = syn.make_workload(n, k, h, K, P, seed=7 + n)
    g = w.g if use_grid else None
    got = veng.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=g, step=w.step,
                                    K=K)
    want = _oracle_joint_all(oracle, w, K, use_grid)
    assert (got["info"] == 0).all()
    assert rel(got["logw"], want["logw"]) < RTOL
    assert rel(got["mu"], want["mu"]) < RTOL
    assert rel(got["L"], want["L"]) < RTOL


def test_forecast_instances_per_scenario_theta(veng, oracle):
    n, k, h, P, K = 60, 1, 4, 5, 4
    w = syn.make_workload(n, k, h, K, P, seed=21)
    th, nz = syn.perturbed_theta(w.ens, K, seed=5)
    got = veng.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                    theta=th, noise=nz)
    want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                     use_joint=True, theta_per_scenario=th, noise_per_scenario=nz)
    assert rel(got["logw"], want["logw"]) < RTOL
    assert rel(got["mu"], want["mu"]) < RTOL
    assert rel(got["L"], want["L"]) < RTOL


def test_reference_schedule_agrees(veng, oracle):
    """Device (one joint factorisation) vs the oracle's REFERENCE schedule (three factorisations,
    LU solves in predict_mvn, Cholesky of Sigma*)."""
    n, k, h, P, K = 150, 1, 9, 8, 3
    w = syn.make_workload(n, k, h, K, P, seed=33)
    got = veng.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb)
    want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, use_joint=False)
    assert rel(got["logw"], want["logw"]) < RTOL
    assert rel(got["mu"], want["mu"]) < 1e-8     # LU vs Cholesky: cond(K)·eps, see DESIGN.md
    assert rel(got["L"], want["L"]) < 1e-8


def test_factor_append_predict_fast_path(veng, oracle):
    n, k, h, P, K = 150, 2, 9, 8, 16
    w = syn.make_workload(n, k, h, K, P, seed=44)
    f = veng.factor_store(w.ens, n, k, h, w.t, w.y1, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
    logw, mu = veng.append(f, w.y2)
    _, L = veng.predict(f, want_mu=False)
    want = _oracle_joint_all(oracle, w, K, True)
    assert rel(logw, want["logw"]) < RTOL
    assert rel(mu, want["mu"]) < RTOL
    assert rel(L, want["L"][0]) < RTOL
    ref_lm, _ = oracle.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    assert rel(f.logml_n, ref_lm) < RTOL
    f.free()


def test_predict_without_nowcast(veng, oracle):
    n, h, P = 80, 12, 4
    w = syn.make_workload(n, 0, h, 1, P, seed=51)
    f = veng.factor_store(w.ens, n, 0, h, w.t, w.y1, None, w.ya, w.yb)
    mu, L = veng.predict(f)
    for p, tr in enumerate(w.trees):
        prog, th = kn.flatten(tr)
        r = oracle.instance_joint(prog, th, w.noise[p], n, 0, h, w.t, w.y1, w.ya, w.yb)
        assert rel(mu[p], r["mu"]) < RTOL and rel(L[p], r["L"]) < RTOL
    f.free()


def test_draws_bit_identical(engine, oracle):
    rng = np.random.default_rng(3)
    K, P, h, D = 7, 5, 9, 20
    logw = rng.standard_normal((K, P))
    mu = rng.standard_normal((K, P, h))
    L = np.tril(rng.standard_normal((K, P, h, h)))
    zeta = rng.standard_normal((K, D, h))
    comp = rng.integers(0, P, (K, D)).astype(np.int32)
    x, ess, _ = engine.draw(logw, mu, L, zeta, comp=comp)
    xo, esso, _ = oracle.draws(logw, mu, L, zeta, comp=comp)
    assert np.array_equal(np.asarray(x), np.asarray(xo))          # bit-identical
    assert rel(ess, esso) < 1e-12
    # inverse-CDF component pick and resampling from supplied uniforms
    u = rng.uniform(size=(K, D))
    u_res = rng.uniform(size=(K, P))
    for thr in (0.0, 1.0):
        x2, _, c2 = engine.draw(logw, mu, L, zeta, u=u, u_res=u_res, ess_thr=thr)
        xo2, _, co2 = oracle.draws(logw, mu, L, zeta, u=u, u_res=u_res, ess_thr=thr)
        assert np.array_equal(c2, co2)
        assert np.array_equal(np.asarray(x2), np.asarray(xo2))
    # scenario-shared factors (stride 0)
    x3, _, _ = engine.draw(logw, mu, L[0], zeta, comp=comp)
    xo3, _, _ = oracle.draws(logw, mu, L[0], zeta, comp=comp)
    assert np.array_equal(np.asarray(x3), np.asarray(xo3))


def test_fused_forecast_with_nowcasts(veng, oracle):
    n, k, h, P, K, D = 150, 1, 9, 8, 50, 20
    w = syn.make_workload(n, k, h, K, P, seed=61)
    rng = np.random.default_rng(9)
    zeta = rng.standard_normal((K, D, h))
    u = rng.uniform(size=(K, D))
    logw = np.empty((K, P))
    ess = np.empty(K)
    x = veng.forecast_with_nowcasts(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, zeta, w.ya, w.yb, g=w.g,
                                      step=w.step, u=u, logw=logw, ess=ess)
    assert x.shape == (h, K * D)
    want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                     use_joint=False)
    assert rel(logw, want["logw"]) < RTOL
    xo, esso, _ = oracle.draws(want["logw"], want["mu"], want["L"], zeta, u=u)
    assert rel(ess, esso) < 1e-8
    assert rel(x, xo) < 1e-8
    # draw stage in isolation is bit-exact: feed the oracle the device's own moments
    f = veng.factor_store(w.ens, n, k, h, w.t, w.y1, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
    lw, mu = veng.append(f, w.y2)
    _, L = veng.predict(f, want_mu=False)
    xb, _, comp = veng.draw(lw, mu, L, zeta, u=u)
    xob, _, _ = oracle.draws(lw, mu, L, zeta, comp=comp)
    assert np.array_equal(np.asarray(xb), np.asarray(xob))
    assert np.array_equal(np.asarray(xb), np.asarray(x))
    f.free()


def test_not_positive_definite_reports_info(veng):
    # flat series + zero noise/jitter → singular Gram (issue #51 regression, test_model_fitting.jl:97-98)
    from nowcastautogp_b200.engine import PosDefError
    ens = kn.pack_ensemble([kn.Constant(1.0)], [0.0])
    veng.set_jitter(0.0)
    try:
        t = np.linspace(0, 1, 12)
        y = np.ones(12)
        lm, info = veng.logml_batch(ens, t, y)
        assert info[0] == 2 and np.isnan(lm[0])
        with pytest.raises(PosDefError):
            veng.logml_batch(ens, t, y, check=True)
    finally:
        veng.set_jitter(1e-5)


def test_bad_program_rejected(engine):
    from nowcastautogp_b200.engine import NagpError
    ens = kn.pack_ensemble([kn.Constant(1.0)], [0.1])
    ens.prog[0] = 6  # Plus with empty stack
    with pytest.raises(NagpError):
        engine.logml_batch(ens, np.linspace(0, 1, 5), np.zeros(5))


def test_failures_in_later_tile_columns_mixed_batch(veng, oracle):
    # Half of a large batch fails (negative "noise" makes the Gram indefinite) somewhere inside the factorisation,
    # the other half is fine: every instance must report the oracle's leading-minor index or the oracle's value,
    # and the kernel must get through the batch (two resident matrices per SM, failures flagged by the chain warp
    # while the row owners are a column behind).
    n, P, K = 150, 32, 12
    w = syn.make_workload(n, 1, 9, K, P, seed=77)
    theta_k, noise_k = syn.perturbed_theta(w.ens, K, seed=5)
    noise_k = np.array(noise_k, copy=True)
    noise_k[:, 1::2] = -0.02 - 0.01 * np.arange(P // 2)[None, :]
    got = veng.forecast_instances(w.ens, n, 1, 9, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                  theta=theta_k, noise=noise_k)
    info = np.asarray(got["info"]).reshape(K, P)
    assert (info[:, 0::2] == 0).all()
    assert (info[:, 1::2] > 0).all() and (info[:, 1::2] > 8).any()      # some fail beyond the first tile column
    logw = np.asarray(got["logw"]).reshape(K, P)
    assert np.isnan(logw[:, 1::2]).all() and np.isfinite(logw[:, 0::2]).all()
    for s in (0, K - 1):
        y = np.concatenate([w.y1, w.y2[s]])
        for p in range(P):
            prog, _ = kn.flatten(w.trees[p])
            off = w.ens.theta_off
            r = oracle.instance_joint(prog, theta_k[s, off[p]:off[p + 1]], noise_k[s, p], n, 1, 9, w.t, y, w.ya, w.yb,
                                      g=w.g, step=w.step)
            assert r["info"] == info[s, p], (s, p)
            if r["info"] == 0:
                assert rel(logw[s, p], w.logw0[p] + r["logml_m"] - r["logml_n"]) < RTOL


def test_forecast_with_nowcasts_theta_one_call(veng, oracle):
    """nagp_forecast_with_nowcasts_theta = nagp_forecast_instances + nagp_draw without the round trip: draws bit-identical
    to the oracle's draw routine fed the oracle's own moments only up to the moments' 1e-9, so compare against the
    two-call device path bit for bit and against the oracle at the moment tolerance."""
    n, k, h, P, K, D = 60, 1, 4, 5, 6, 7
    w = syn.make_workload(n, k, h, K, P, seed=91)
    th, nz = syn.perturbed_theta(w.ens, K, seed=92)
    rng = np.random.default_rng(93)
    zeta, u = rng.standard_normal((K, D, h)), rng.uniform(size=(K, D))
    logw = np.empty((K, P))
    x1 = veng.forecast_with_nowcasts_theta(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, zeta, th, nz, w.ya, w.yb, g=w.g,
                                           step=w.step, u=u, logw=logw)
    r = veng.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step, theta=th, noise=nz)
    x2, _, _ = veng.draw(r["logw"], r["mu"], r["L"], zeta, u=u)
    assert np.array_equal(np.asarray(x1), np.asarray(x2)) and np.array_equal(logw, r["logw"])
    want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                     use_joint=True, theta_per_scenario=th, noise_per_scenario=nz)
    xo, _, _ = oracle.draws(want["logw"], want["mu"], want["L"], zeta, u=u)
    assert rel(logw, want["logw"]) < RTOL and rel(x1, xo) < 1e-8
