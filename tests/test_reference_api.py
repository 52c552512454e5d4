"""The reference's own test items, restated against the Python mirror of its public API.

CPU part (no GPU): the exactly-pinned boundary tests of `create_nowcast_data` / `TData`
(`/root/reference/test/test_nowcast_functions.jl:52-107`, `test/test_helper_functions.jl:25-98`) and the
host helpers (`_stabilize_for_fit`, `linear_schedule`, assertion behaviour).
GPU part (`-m gpu`): `make_and_fit_model` / `forecast` / `forecast_with_nowcasts` items
(`test/test_nowcast_functions.jl:142-295`, `test/test_forecasting.jl`, `test/test_model_fitting.jl:23-138`,
`test/test_gpconfig.jl:37-43`) — like the reference they assert shapes, finiteness, sign and the one
band, because the reference pins no numeric GP output; numeric parity lives in test_gpu_parity.py."""
import warnings

import numpy as np
import pytest

import nowcastautogp_b200 as ng
from nowcastautogp_b200.api import _stabilize_for_fit, linear_schedule


def D(s):
    return np.datetime64(s, "D")


def drange(a, b):
    return np.arange(D(a), D(b) + 1)


DATES3 = np.array([D("2024-01-01"), D("2024-01-02"), D("2024-01-03")])
VEC = [[10.0, 11.0, 12.0], [9.5, 10.8, 11.2], [10.2, 11.1, 12.1]]
MAT = np.array([[10.0, 9.5, 10.2], [11.0, 10.8, 11.1], [12.0, 11.2, 12.1]])


# ---- create_nowcast_data / TData: exact equality, as in the reference -------------------------------
def test_create_nowcast_data_vector_input():          # test_nowcast_functions.jl:52-61
    r = ng.create_nowcast_data(VEC, DATES3)
    assert len(r) == 3
    assert all(np.array_equal(x.ds, DATES3) for x in r)
    assert r[0].y.tolist() == [10.0, 11.0, 12.0] and r[0].values.tolist() == [10.0, 11.0, 12.0]
    assert r[1].y.tolist() == [9.5, 10.8, 11.2] and r[2].y.tolist() == [10.2, 11.1, 12.1]


def test_create_nowcast_data_matrix_equals_vector():  # :63-80
    rv, rm = ng.create_nowcast_data(VEC, DATES3), ng.create_nowcast_data(MAT, DATES3)
    assert len(rm) == 3
    for a, b in zip(rv, rm):
        assert np.array_equal(a.ds, b.ds) and np.array_equal(a.y, b.y) and np.array_equal(a.values, b.values)


def test_create_nowcast_data_transformation():        # :82-90
    r = ng.create_nowcast_data([[1.0, 2.0], [1.5, 2.5]], DATES3[:2], transformation=np.log)
    assert np.allclose(r[0].y, np.log([1.0, 2.0]), rtol=0, atol=0)
    assert r[0].values.tolist() == [1.0, 2.0]
    assert r[0].y.dtype == np.float64 and r[0].values.dtype == np.float64   # :137-139 type promotion


def test_create_nowcast_data_assertions():            # :92-107
    with pytest.raises(AssertionError):
        ng.create_nowcast_data([], DATES3)
    with pytest.raises(AssertionError):
        ng.create_nowcast_data([[1.0, 2.0], [1.0, 2.0, 3.0]], DATES3[:2])
    with pytest.raises(AssertionError):
        ng.create_nowcast_data(VEC, DATES3[:2])
    with pytest.raises(AssertionError):
        ng.TData(DATES3, [1.0, 2.0], transformation=lambda v: v)   # test_helper_functions.jl:67-70


def test_tdata_fields_and_promotion():                # test_helper_functions.jl:25-65
    t = ng.TData(DATES3, [1, 2, 3], transformation=lambda v: v * 0.5)
    assert t.y.tolist() == [0.5, 1.0, 1.5] and t.values.tolist() == [1.0, 2.0, 3.0]
    assert t.y.dtype == t.values.dtype == np.float64 and len(t) == 3
    t2 = ng.create_transformed_data(iter(DATES3), iter([4.0, 5.0, 6.0]), transformation=np.sqrt)
    assert np.array_equal(t2.y, np.sqrt([4.0, 5.0, 6.0]))


def test_stabilize_for_fit():                         # test_model_fitting.jl:126-138
    y = np.array([1.0, 5.0, 2.0, 8.0])
    assert _stabilize_for_fit(y) is not None and np.array_equal(_stabilize_for_fit(y), y)   # identity on healthy data
    flat = np.full(10, 75000.0)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        out = _stabilize_for_fit(flat, rng=np.random.default_rng(1))
    assert len(wlist) == 1 and not np.array_equal(out, flat) and out.std() > 0
    assert abs(out.mean() - 75000.0) < 5 * 1e-3 * 75001.0


def test_linear_schedule():
    assert linear_schedule(10, 0.1) == list(range(1, 11))
    assert linear_schedule(10, 0.5) == [5, 10]
    assert linear_schedule(157, 0.1)[-1] == 157 and linear_schedule(157, 0.1)[0] == 16
    assert linear_schedule(7, 1.0 / 7)[-1] == 7


def test_forecast_with_nowcasts_assertions_need_no_device():   # forecasting.jl:123-126 fire before any device work
    with pytest.raises(AssertionError):
        ng.forecast_with_nowcasts(None, [], DATES3, 5)
    t = ng.TData(DATES3[:1], [1.0], transformation=lambda v: v)
    with pytest.raises(AssertionError):
        ng.forecast_with_nowcasts(None, [t], DATES3, 5, n_mcmc=5, n_hmc=0)
    with pytest.raises(AssertionError):
        ng.forecast_with_nowcasts(None, [t], DATES3, 5, ess_threshold=1.5)
    with pytest.raises(AssertionError):
        ng.forecast_with_nowcasts(None, [t], DATES3, 5, forecast_n_hmc=0)


def test_make_and_fit_model_requires_mcmc_kwargs():   # test_gpconfig.jl:37-43 (UndefKeywordError)
    t = ng.TData(DATES3, [1.0, 2.0, 3.0], transformation=lambda v: v)
    with pytest.raises(TypeError):
        ng.make_and_fit_model(t, n_particles=1)


# ---- GPU: the reference's integration items --------------------------------------------------------------
VALUES10 = [10.0, 15.0, 12.0, 18.0, 22.0, 25.0, 20.0, 16.0, 14.0, 11.0]


@pytest.fixture(scope="module")
def base_model(engine):
    data = ng.create_transformed_data(drange("2024-01-01", "2024-01-10"), VALUES10, transformation=lambda x: x)
    return ng.make_and_fit_model(data, n_particles=1, n_mcmc=5, n_hmc=5, rng=np.random.default_rng(123), engine=engine)


NOWCAST_DATES = np.array([D("2024-01-11"), D("2024-01-12")])
SINGLE_DATES = NOWCAST_DATES[:1]
ident = lambda x: x   # noqa: E731


@pytest.mark.gpu
def test_fwn_basic_shape(base_model):                 # test_nowcast_functions.jl:142-153
    sc = [ng.TData(NOWCAST_DATES, [12.0, 13.0], transformation=ident), ng.TData(NOWCAST_DATES, [11.5, 12.8], transformation=ident)]
    r = ng.forecast_with_nowcasts(base_model, sc, np.array([D("2024-01-13"), D("2024-01-14")]), 10)
    assert r.shape == (2, 20) and np.isfinite(r).all()


@pytest.mark.gpu
def test_fwn_single_and_transform(base_model):        # :155-177
    single = [ng.TData(SINGLE_DATES, [12.0], transformation=ident)]
    assert ng.forecast_with_nowcasts(base_model, single, [D("2024-01-12")], 5).shape == (1, 5)
    logged = [ng.TData(SINGLE_DATES, [np.log(12.0)], transformation=ident)]
    r = ng.forecast_with_nowcasts(base_model, logged, [D("2024-01-12")], 3, inv_transformation=np.exp)
    assert r.shape == (1, 3) and (r > 0).all()


@pytest.mark.gpu
def test_fwn_mcmc_options_resampling_and_dates(base_model):   # :179-209
    single = [ng.TData(SINGLE_DATES, [12.0], transformation=ident)]
    fd = [D("2024-01-12")]
    assert ng.forecast_with_nowcasts(base_model, single, fd, 2, n_mcmc=0, n_hmc=2).shape == (1, 2)
    assert ng.forecast_with_nowcasts(base_model, single, fd, 2, n_mcmc=2, n_hmc=2).shape == (1, 2)
    assert ng.forecast_with_nowcasts(base_model, single, fd, 2, n_mcmc=0, n_hmc=0).shape == (1, 2)
    assert ng.forecast_with_nowcasts(base_model, single, fd, 2, ess_threshold=0.5).shape == (1, 2)
    two = [ng.TData(SINGLE_DATES, [12.0], transformation=ident), ng.TData(SINGLE_DATES, [11.8], transformation=ident)]
    r = ng.forecast_with_nowcasts(base_model, two, drange("2024-01-12", "2024-01-15"), 3)
    assert r.shape == (4, 6)


@pytest.mark.gpu
def test_fwn_consistency_and_base_model_not_mutated(base_model):   # :238-246, forecasting.jl:101
    single = [ng.TData(SINGLE_DATES, [12.0], transformation=ident)]
    before = base_model.to_dict()
    r1 = ng.forecast_with_nowcasts(base_model, single, [D("2024-01-12")], 5)
    r2 = ng.forecast_with_nowcasts(base_model, single, [D("2024-01-12")], 5)
    assert r1.shape == r2.shape and np.isfinite(r1).all() and np.isfinite(r2).all()
    after = base_model.to_dict()
    assert before["n_obs"] == after["n_obs"] and before["log_weights"] == after["log_weights"]


@pytest.mark.gpu
def test_fwn_two_particles_all_paths(engine):         # the threading regression item :248-281
    data = ng.create_transformed_data(drange("2024-01-01", "2024-01-10"), VALUES10, transformation=ident)
    m = ng.make_and_fit_model(data, n_particles=2, n_mcmc=5, n_hmc=3, rng=np.random.default_rng(456), engine=engine)
    nc = [ng.TData(NOWCAST_DATES, [12.0, 13.0], transformation=ident)]
    fd = [D("2024-01-13")]
    r = ng.forecast_with_nowcasts(m, nc, fd, 3, n_mcmc=2, n_hmc=2)
    assert r.shape == (1, 3) and np.isfinite(r).all()
    r = ng.forecast_with_nowcasts(m, nc, fd, 3, n_mcmc=2, n_hmc=2, forecast_n_hmc=1)
    assert r.shape == (1, 3) and np.isfinite(r).all()


@pytest.mark.gpu
def test_fwn_with_create_nowcast_data(base_model):    # :283-295
    sc = ng.create_nowcast_data(np.array([[12.0, 11.8], [13.0, 12.5]]), NOWCAST_DATES)
    assert ng.forecast_with_nowcasts(base_model, sc, [D("2024-01-13")], 3).shape == (1, 6)


@pytest.mark.gpu
def test_forecast_shapes_and_transforms(base_model):  # test_forecasting.jl:32-116
    fd = drange("2024-01-11", "2024-01-15")
    r = ng.forecast(base_model, fd, 20)
    assert r.shape == (5, 20) and np.issubdtype(r.dtype, np.floating)
    assert ng.forecast(base_model, fd[:1], 1).shape == (1, 1)
    assert (ng.forecast(base_model, fd, 10, inv_transformation=np.exp) > 0).all()
    lg = ng.forecast(base_model, fd, 10, inv_transformation=lambda x: 1.0 / (1.0 + np.exp(-x)))
    assert ((lg > 0) & (lg < 1)).all()
    assert ng.forecast(base_model, fd, 4, forecast_n_hmc=2).shape == (5, 4)
    a, b = ng.forecast(base_model, fd, 5), ng.forecast(base_model, fd, 5)
    assert a.shape == b.shape and not np.array_equal(a, b)


@pytest.mark.gpu
def test_flat_and_constant_series_issue_51(engine):   # test_model_fitting.jl:97-124
    flat_dates = drange("2024-01-01", "2024-01-10")
    fdates = drange("2024-01-11", "2024-01-18")
    flat = [75000.0, 75100.0, 74950.0, 75050.0, 75000.0, 74980.0, 75020.0, 75010.0, 74990.0, 75005.0]
    for vals, seed, band in ((flat, 51, True), ([75000.0] * 10, 52, False)):
        data = ng.create_transformed_data(flat_dates, vals, transformation=np.log)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = ng.make_and_fit_model(data, smc_data_proportion=0.5, n_particles=1, n_mcmc=5, n_hmc=3,
                                      rng=np.random.default_rng(seed), engine=engine)
        assert isinstance(m, ng.GPModel)
        fc = ng.forecast(m, fdates, 25, inv_transformation=np.exp)
        assert fc.shape == (8, 25) and np.isfinite(fc).all() and (fc >= 0).all()
        if band:
            assert 50_000 < fc.mean() < 100_000


@pytest.mark.gpu
def test_fwn_per_scenario_resampling_swaps_whole_particles(engine):
    """forecasting.jl:138-141 with `ess_threshold = 1` and `n_hmc > 0`: every scenario resamples, and resampling swaps
    whole particles. The particles here have three hyperparameter slots each but DIFFERENT structures (Linear /
    GammaExponential / Periodic), so a parent's z must never be read under another particle's program: the call has
    to agree with the reference's one-model-per-scenario schedule run by hand with the same generator."""
    from nowcastautogp_b200.api import _forecast_with_nowcasts
    from nowcastautogp_b200.gpmodel import GPModel, Particle
    data = ng.create_transformed_data(drange("2024-01-01", "2024-01-10"), VALUES10, transformation=ident)
    m = GPModel(data.ds, data.y, n_particles=3, rng=np.random.default_rng(8), engine=engine)
    m.particles = [Particle(bytes([2]), np.array([0.3, -0.2, 0.1]), -0.5),       # Linear: intercept is an identity slot
                   Particle(bytes([4]), np.array([0.4, 0.0, -0.3]), -0.4),       # GammaExponential
                   Particle(bytes([5]), np.array([-0.1, 0.6, 0.2]), -0.6)]       # Periodic
    m.fit_smc(schedule=[len(data.y)], n_mcmc=0, n_hmc=0, shuffle=False)
    sc = ng.create_nowcast_data(np.array([[12.0, 11.8, 12.4], [13.0, 12.5, 13.3]]), NOWCAST_DATES)
    fd = np.array([D("2024-01-13"), D("2024-01-14")])
    x, lw = _forecast_with_nowcasts(m, sc, fd, 4, n_hmc=2, ess_threshold=1.0, rng=np.random.default_rng(31))
    assert x.shape == (2, 12) and np.isfinite(x).all()
    assert np.array_equal(lw, np.zeros((3, 3)))                                   # weights are reset after resampling
    # by hand: one model copy per scenario, in scenario order, same generator
    rng = np.random.default_rng(31)
    blocks = []
    for nc in sc:
        c = GPModel.from_dict(m.to_dict(), engine=engine, rng=rng)
        c.add_data(nc.ds, nc.y)
        assert c.maybe_resample(1.0 * c.num_particles())
        c.mcmc_parameters(2)
        blocks.append(ng.forecast(c, fd, 4))
    assert np.array_equal(x, np.hstack(blocks))
    # scenarios that do not trigger stay in the batched path: shapes and finiteness with a threshold nobody reaches
    x2, lw2 = _forecast_with_nowcasts(m, sc, fd, 4, n_hmc=2, ess_threshold=0.0, rng=np.random.default_rng(31))
    assert x2.shape == (2, 12) and np.isfinite(x2).all() and np.isfinite(lw2).all()


@pytest.mark.gpu
def test_zero_weight_particle_does_not_break_forecasts(engine):
    """A particle whose Gram is not positive definite leaves `fit_smc` with weight -inf (it carries no mass). Forecasts,
    `add_data` and `forecast_with_nowcasts` must keep working on the particles that count (the MvNormal constructor of the
    reference never sees a zero-weight component's factorisation fail the mixture)."""
    from nowcastautogp_b200.gpmodel import GPModel, Particle
    data = ng.create_transformed_data(drange("2024-01-01", "2024-01-10"), VALUES10, transformation=ident)
    m = GPModel(data.ds, data.y, n_particles=2, rng=np.random.default_rng(5), engine=engine)
    m.particles = [Particle(bytes([4]), np.array([0.2, 0.0, -0.3]), -0.4),
                   Particle(bytes([2]), np.array([0.3, 40.0, 40.0]), -200.0)]  # Linear, amplitude 1e16: rank 2 + jitter, not PD in FP64
    m.config.noise = None
    m.fit_smc(schedule=[len(data.y)], n_mcmc=0, n_hmc=0, shuffle=False, ess_fraction=0.0)
    assert np.isneginf(m.log_weights[1]) and np.isfinite(m.log_weights[0])     # the fit flags it, keeps it, gives it no mass
    fd = drange("2024-01-11", "2024-01-13")
    x = ng.forecast(m, fd, 6)
    assert x.shape == (3, 6) and np.isfinite(x).all()
    sc = [ng.TData(SINGLE_DATES, [12.0], transformation=ident)]
    r = ng.forecast_with_nowcasts(m, sc, [D("2024-01-12")], 4)
    assert r.shape == (1, 4) and np.isfinite(r).all()
