"""`get_transformations` (host mirror, restating the reference's own test items in
`test/test_helper_functions.jl:100-426`), the oracle's inverse / quantile restatements pinned against numpy, and
— on the GPU — `nagp_forecast_summary` against the oracle."""
import numpy as np
import pytest

from nowcastautogp_b200.transformations import (InverseTransform, fit_boxcox_lambda, get_transformations,
                                                inverse_reference)
from oracle import summary as osum

values = [10.0, 15.0, 12.0, 18.0, 22.0, 25.0, 20.0, 16.0, 14.0, 11.0]
values_with_zero = [0.0, 15.0, 12.0, 0.0, 22.0, 25.0, 0.0, 16.0, 14.0, 11.0]
test_values = [0.5, 1.0, 2.0, 5.0, 10.0, 20.0, 50.0]
positive_values = [0.1, 1.0, 5.0, 10.0, 100.0]
percentage_values = [10.0, 25.0, 50.0, 75.0, 90.0]
boxcox_values = [1.0, 2.0, 5.0, 10.0, 20.0]


def roundtrip(fwd, inv, vals, atol=0.0, rtol=0.0):
    for v in vals:
        rec = inv(fwd(v))
        assert isinstance(rec, float)
        assert abs(rec - v) <= atol + rtol * abs(v), (v, rec)


# ---- reference test items, test/test_helper_functions.jl ----------------------------------------------------
def test_percentage():                                   # :100-113, :167-181
    for data in (test_values, values_with_zero):
        fwd, inv = get_transformations("percentage", data)
        assert callable(fwd) and callable(inv)
        roundtrip(fwd, inv, percentage_values, atol=1e-10)


def test_positive():                                     # :115-129, :183-197
    for data in (positive_values, values_with_zero):
        fwd, inv = get_transformations("positive", data)
        roundtrip(fwd, inv, [v for v in positive_values if v > 0], atol=1e-6)


def test_boxcox():                                       # :131-144, :199-212
    for data in (boxcox_values, values_with_zero):
        fwd, inv = get_transformations("boxcox", data)
        roundtrip(fwd, inv, boxcox_values, atol=1e-6)


def test_boxcox_fallback_on_flat_data():                 # :146-165 (issue #51)
    flat = [75000.0, 75100.0, 74950.0, 75050.0, 75000.0, 74980.0, 75020.0, 75010.0, 74990.0, 75005.0]
    fwd, inv = get_transformations("boxcox", flat)
    assert abs(fwd(flat[0]) - np.log(flat[0])) <= 1e-9 * np.log(flat[0])
    roundtrip(fwd, inv, flat, rtol=1e-9)
    healthy, _ = get_transformations("boxcox", values)
    assert not np.isclose(healthy(values[0]), np.log(values[0]), rtol=1e-9, atol=0)


def test_boxcox_edge_cases():                            # :214-243
    small = [1.0e-8, 1.0e-6, 1.0e-4, 0.001, 0.01, 0.1, 1.0, 10.0]
    fwd, inv = get_transformations("boxcox", small)
    roundtrip(fwd, inv, small, atol=1e-6)
    for y in [-100.0, -50.0, -20.0, -10.0, 100.0, 50.0, 20.0, 10.0]:
        r = inv(y)
        assert r >= 0.0 and np.isfinite(r)


def test_boxcox_negative_lambda():                       # :245-265
    dec = [100.0, 50.0, 25.0, 12.5, 6.25, 3.125]
    fwd, inv = get_transformations("boxcox", dec)
    roundtrip(fwd, inv, dec, atol=1e-4)
    for y in [-5.0, -2.0, -1.0, -0.5, -0.1, 0.0, 0.1, 0.5, 1.0, 2.0, 5.0]:
        r = inv(y)
        assert r >= 0.0 and np.isfinite(r)


def test_boxcox_zero_lambda_and_stability():             # :267-306
    fwd, inv = get_transformations("boxcox", [1.0, 2.718, 7.389, 20.086, 54.598])
    for y in [-10.0, -5.0, -1.0, 0.0, 1.0, 5.0, 10.0]:
        assert inv(y) >= 0.0 and np.isfinite(inv(y))
    roundtrip(fwd, inv, [1.0, 2.718, 7.389, 20.086, 54.598], atol=1e-5)
    extreme = [1.0e-10, 1.0e-5, 1.0e-2, 1.0, 1.0e2, 1.0e5, 1.0e8]
    fwd, inv = get_transformations("boxcox", extreme)
    for v in extreme:
        assert np.isfinite(fwd(v))
    roundtrip(fwd, inv, extreme, rtol=1e-3)


def test_integer_and_mixed_data():                       # :308-426
    for data in ([1, 2, 5, 8, 10, 15, 20, 25, 30], [1, 2.5, 5, 7.8, 10, 12.3, 15], [1, 2, 5, 8, 10, 15, 20, 25, 30, 0],
                 np.float32([1.0, 2.0, 3.0, 4.0, 5.0]), [0, 1, 2, 3, 4, 5]):
        fwd, inv = get_transformations("boxcox", data)
        roundtrip(fwd, inv, [float(v) for v in data if v > 0], atol=1e-6)
    fwd, inv = get_transformations("positive", [0, 1, 2, 3, 4, 5])
    roundtrip(fwd, inv, [1.0, 2.0, 3.0, 4.0, 5.0], atol=1e-6)
    fwd, inv = get_transformations("percentage", [0, 10, 25, 50, 75, 90])
    roundtrip(fwd, inv, [10.0, 25.0, 50.0, 75.0, 90.0], atol=1e-6)


def test_unknown_name_and_bad_values():                  # :428-, src/transformations.jl:53-54,168
    with pytest.raises(AssertionError):
        get_transformations("unknown", values)
    with pytest.raises(AssertionError):
        get_transformations("positive", [])
    with pytest.raises(AssertionError):
        get_transformations("positive", [1.0, -2.0])


def test_boxcox_lambda_is_the_likelihood_maximiser():
    from scipy import stats
    rng = np.random.default_rng(0)
    x = np.exp(rng.standard_normal(200)) ** 1.7 + 0.3
    assert abs(fit_boxcox_lambda(x) - stats.boxcox_normmax(x, method="mle")) < 1e-5


# ---- oracle pinning -----------------------------------------------------------------------------------------
SPECS = [(0, 0.0, 0.0, 0.0), (1, 0.0, 0.0, 0.0), (1, 0.0, 5.5, 0.0), (2, 0.0, 0.0, 0.0), (2, 0.0, 5.0, 0.0),
         (3, 0.31, 0.0, 25.0), (3, -0.45, 0.5, 100.0), (3, 0.0, 0.25, 9.0), (3, -2.0, 0.0, 3.0)]


def _inputs(rng, n):
    return np.concatenate([rng.standard_normal(n) * 3.0, [-800.0, -60.0, -1e-3, 0.0, 1e-3, 2.0, 40.0, 800.0],
                           np.linspace(0.4, 0.6, 9)])        # lam = -2: lam*y + 1 crosses 1e-10 and 0 near 0.5


def test_oracle_inverse_matches_host_mirror_and_scipy():
    from scipy.special import expit, inv_boxcox
    y = _inputs(np.random.default_rng(1), 64)
    for spec in SPECS:
        want = np.array([osum.inverse_scalar(*spec, float(v)) for v in y])
        got = inverse_reference(*spec, y)
        assert np.all(np.isfinite(want) | np.isinf(want))
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)
    # independent library forms where no clamp is active
    mid = y[(y > -30) & (y < 30)]
    np.testing.assert_allclose([osum.inverse_scalar(2, 0, 0, 0, float(v)) for v in mid], expit(mid) * 100, rtol=1e-13)
    ok = mid[0.31 * mid + 1 > 1e-3]
    np.testing.assert_allclose([osum.inverse_scalar(3, 0.31, 0, 25.0, float(v)) for v in ok], inv_boxcox(ok, 0.31),
                               rtol=1e-13)


def test_oracle_quantile_is_type7():
    rng = np.random.default_rng(2)
    for n in (1, 2, 3, 10, 101, 1000):
        v = rng.standard_normal(n)
        for p in (0.0, 0.025, 0.25, 0.5, 0.75, 0.975, 1.0, 1.0 / 3.0):
            assert abs(osum.quantile_type7(v.tolist(), p) - np.quantile(v, p, method="linear")) <= 1e-15 * max(1.0, np.abs(v).max())


# ---- GPU: nagp_forecast_summary vs oracle ---------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("spec", SPECS)
def test_device_inverse_transform_matches_oracle(engine, spec):
    """Tolerance: 8 ulp of the value before the offset is subtracted (CUDA exp <= 1 ulp, pow <= 2 ulp, against the
    host libm's; `exp(y) - offset` cancels, so the bound is on |result| + offset)."""
    y = _inputs(np.random.default_rng(3), 500)
    h = 7
    y = y[:(len(y) // h) * h].reshape(h, -1)
    got, _ = engine.forecast_summary(y, spec)
    want = np.array([[osum.inverse_scalar(*spec, float(v)) for v in row] for row in y])
    big = np.isinf(want)
    assert np.array_equal(np.isinf(got), big)
    scale = np.abs(want[~big]) + spec[2]
    assert np.all(np.abs(got[~big] - want[~big]) <= 8 * 2.2e-16 * scale)


@pytest.mark.gpu
@pytest.mark.parametrize("h,N", [(1, 1), (1, 2), (4, 3), (9, 1000), (3, 20000), (4, 4097)])
def test_device_quantiles_bit_identical(engine, h, N):
    """Identity transformation: the order statistics are selected exactly, so the quantiles equal the oracle's
    bit for bit (including ties, negative values and signed zeros)."""
    rng = np.random.default_rng(h * 100003 + N)
    x = rng.standard_normal((h, N)) * np.exp(rng.standard_normal((h, 1)) * 3)
    if N >= 1000:
        x[:, ::7] = np.round(x[:, ::7], 1)           # ties
        x[0, :50] = 0.0
        x[0, 50:60] = -0.0
    probs = [0.0, 0.025, 0.25, 0.5, 0.75, 0.975, 1.0, 1.0 / 3.0]
    xo, q = engine.forecast_summary(x, (0, 0.0, 0.0, 0.0), probs)
    assert np.array_equal(xo, x)
    for r in range(h):
        for j, p in enumerate(probs):
            assert q[r, j] == osum.quantile_type7(x[r].tolist(), p), (r, p)


@pytest.mark.gpu
def test_device_summary_of_transformed_draws(engine):
    """The vignette's last step: inverse Box-Cox / log of the draws, then 25/50/75 % bands per date."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((9, 20000)) * 0.4 + 3.0
    for name in ("positive", "boxcox", "percentage"):
        _, inv = get_transformations(name, values)
        assert isinstance(inv, InverseTransform)
        xo, q = engine.forecast_summary(x, inv.spec, [0.25, 0.5, 0.75])
        want = inv(x)
        np.testing.assert_allclose(xo, want, rtol=1e-14)
        np.testing.assert_allclose(q, np.quantile(want, [0.25, 0.5, 0.75], axis=1).T, rtol=1e-13)
    with pytest.raises(Exception):
        engine.forecast_summary(x, (0, 0, 0, 0), [1.5])
