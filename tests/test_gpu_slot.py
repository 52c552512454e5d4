"""GPU parity of the slot kernel (nagp_fused_v3.cu: three matrices in flight per SM, recycled tile pool, Gram on
demand) against the CPU oracle, with the kernel that ran asserted through nagp_last_kernel.

Tolerance: 1e-9 relative on log-weights, logML, mu and L (BASELINE.json north_star)."""
import numpy as np
import pytest

from nowcastautogp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

RTOL = 1e-9
SLOT, TILE = 3, 2


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture
def seng(engine):
    engine.set_variant(SLOT)
    yield engine
    engine.set_variant(0)


# tile rows 3 (smallest), 9, 20 (vignette shape), 21 (largest); k = 0 and k = 2; h = 1 and h = 12
@pytest.mark.parametrize("n,k,h,P,K", [(12, 1, 4, 3, 5), (60, 1, 4, 5, 7), (150, 1, 9, 8, 6), (150, 2, 16, 6, 3),
                                       (100, 0, 6, 4, 2), (140, 3, 1, 4, 4), (155, 1, 12, 32, 2)])
def test_slot_kernel_matches_oracle(seng, oracle, n, k, h, P, K):
    w = syn.make_workload(n, k, h, K, P, seed=900 + n + h)
    th, nz = syn.perturbed_theta(w.ens, K, seed=n)
    got = seng.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                  theta=th, noise=nz)
    assert seng.last_kernel == SLOT
    want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                     use_joint=True, theta_per_scenario=th, noise_per_scenario=nz)
    assert (got["info"] == 0).all()
    assert rel(got["logw"], want["logw"]) < RTOL
    assert rel(got["mu"], want["mu"]) < RTOL
    assert rel(got["L"], want["L"]) < RTOL


def test_slot_kernel_agrees_with_tile_kernel_on_a_large_batch(seng):
    """Many more instances than matrix slots (3 per SM): every slot streams through tens of instances."""
    n, k, h, P, K = 150, 1, 9, 32, 60
    w = syn.make_workload(n, k, h, K, P, seed=5)
    th, nz = syn.perturbed_theta(w.ens, K, seed=6)
    args = (w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb)
    got = seng.forecast_instances(*args, g=w.g, step=w.step, theta=th, noise=nz)
    assert seng.last_kernel == SLOT
    seng.set_variant(TILE)
    ref = seng.forecast_instances(*args, g=w.g, step=w.step, theta=th, noise=nz)
    assert seng.last_kernel == TILE
    assert (got["info"] == 0).all() and (ref["info"] == 0).all()
    for key in ("logw", "mu", "L"):
        assert rel(got[key], ref[key]) < 1e-11, key


def test_slot_kernel_logml_and_factor_store(seng, oracle):
    n, k, h, P = 120, 2, 6, 9
    w = syn.make_workload(n, k, h, 4, P, seed=77)
    got, info = seng.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    assert seng.last_kernel == SLOT
    want, _ = oracle.logml_batch(w.ens, w.t[:n], w.y1, g=w.g[:n], step=w.step)
    assert (info == 0).all() and rel(got, want) < RTOL
    # scenario-shared fast path on top of a factor stored by the slot kernel (rows >= n kept for proj / Ltail)
    f = seng.factor_store(w.ens, n, k, h, w.t, w.y1, w.logw0, w.ya, w.yb, g=w.g, step=w.step)
    assert seng.last_kernel == SLOT
    logw, mu = seng.append(f, w.y2)
    _, L = seng.predict(f, want_mu=False)
    ref = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                    use_joint=True)
    assert rel(logw, ref["logw"]) < RTOL and rel(mu, ref["mu"]) < RTOL and rel(L, ref["L"][0]) < RTOL
    f.free()


def test_slot_kernel_reports_failed_instances(seng, oracle):
    """Instances whose Gram is not positive definite report the oracle's leading-minor index; their neighbours in
    the same slot stream are unaffected."""
    n, k, h, P, K = 90, 1, 5, 6, 40
    w = syn.make_workload(n, k, h, K, P, seed=31)
    th, nz = syn.perturbed_theta(w.ens, K, seed=32)
    nz = np.array(nz, copy=True)
    nz[1::3, ::2] = -50.0              # negative diagonal: the first pivot fails
    got = seng.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                  theta=th, noise=nz)
    assert seng.last_kernel == SLOT
    want = oracle.forecast_instances(w.ens, n, k, h, w.t, w.y1, w.y2, w.logw0, w.ya, w.yb, g=w.g, step=w.step,
                                     use_joint=True, theta_per_scenario=th, noise_per_scenario=nz)
    assert (got["info"] == want["info"]).all() and (got["info"] > 0).any()
    ok = got["info"] == 0
    assert rel(got["logw"][ok], want["logw"][ok]) < RTOL
    assert rel(got["mu"][ok], want["mu"][ok]) < RTOL
    assert np.isnan(got["logw"][~ok]).all()
