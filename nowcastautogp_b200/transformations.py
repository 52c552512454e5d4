"""`get_transformations(transform_name, values)` — host mirror of `src/transformations.jl:134-170`.

Returns `(forward, inverse)` exactly like the reference: closures that work on a scalar or, vectorised, on
a numpy array. The three built-in inverses carry a `spec` (kind, lambda, offset, max value), so
`forecast` / `forecast_with_nowcasts` can hand the whole `(h, K·D)` draw matrix to the device kernel
(`nagp_forecast_summary`) instead of looping a Python closure over it; an arbitrary user closure
still runs on the host, as in the reference (`src/forecasting.jl:50,73,166`).

Third-party pieces restated here (absent from /root/reference, versions from `Project.toml`):
* `LogExpFunctions.logit / logistic` (compat 0.3.29): `logit(x) = log(x / (1 - x))`; `logistic(x) = e / (1 + e)`,
  `e = exp(x)`, 0 below -744.44…, 1 above 36.73… [R].
* `BoxCox.fit(BoxCoxTransformation, x)` (compat 0.3.7): λ maximising the profile log-likelihood
  `-n/2 · log var_uncorrected(bc_λ(x)) + (λ - 1) Σ log x`, unbounded derivative-free search started at λ = 0
  [R: NLopt BOBYQA, tolerances 1e-8]; `bc_λ(x) = log x` when `|λ| <= 1e-8`, else `(x^λ - 1) / λ`.
"""
from __future__ import annotations

import logging
import math
from typing import Callable, Tuple

import numpy as np
from scipy import optimize

log = logging.getLogger("nowcastautogp_b200")

KIND_IDENTITY, KIND_POSITIVE, KIND_PERCENTAGE, KIND_BOXCOX = 0, 1, 2, 3
_LOGISTIC_LOWER, _LOGISTIC_UPPER = -744.4400719213812, 36.7368005696771


def _get_offset(values: np.ndarray) -> float:
    """`src/transformations.jl:52-62`: half the minimum positive value when the data touch zero, else 0."""
    assert len(values) > 0, "Values array must not be empty"
    assert np.all(values >= 0), "All values must be non-negative for the selected transformations"
    if values.min() == 0:
        return float(values[values > 0].min() / 2)
    return 0.0


def logit(x):
    x = np.asarray(x, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(x / (1.0 - x))


def logistic(x):
    x = np.asarray(x, np.float64)
    with np.errstate(over="ignore"):
        e = np.exp(x)
        mid = e / (1.0 + e)
    return np.where(x < _LOGISTIC_LOWER, 0.0, np.where(x > _LOGISTIC_UPPER, 1.0, mid))


def boxcox(lam: float, x, atol: float = 1e-8):
    x = np.asarray(x, np.float64)
    if abs(lam) <= atol:
        return np.log(x)
    return (np.power(x, lam) - 1.0) / lam


def boxcox_loglikelihood(lam: float, x: np.ndarray) -> float:
    z = boxcox(lam, x)
    var = float(np.var(z))                       # uncorrected
    if not np.isfinite(var) or var <= 0.0:
        return -np.inf
    return -0.5 * len(x) * math.log(var) + (lam - 1.0) * float(np.log(x).sum())


def fit_boxcox_lambda(x: np.ndarray) -> float:
    """λ of `fit(BoxCoxTransformation, x)`: unbounded maximiser of the profile log-likelihood."""
    x = np.asarray(x, np.float64)
    neg = lambda lam: -boxcox_loglikelihood(float(lam), x)
    # bracket outwards from λ = 0, then Brent
    lo, hi, f0 = -1.0, 1.0, neg(0.0)
    for _ in range(12):
        if neg(lo) > f0 and neg(hi) > f0:
            break
        lo, hi = lo * 2.0, hi * 2.0
    res = optimize.minimize_scalar(neg, bracket=(lo, 0.0, hi), method="brent", options={"xtol": 1e-10})
    # On near-constant data the profile likelihood is flat and the maximiser is arbitrary: the reference's unbounded
    # optimiser then returns a pathological λ and `get_transformations` falls back to the log transformation (issue
    # #51, `src/transformations.jl:164-168`). Signal the same case instead of returning whatever Brent stopped at:
    # no measurable likelihood gain over λ = 0, or a non-finite optimum.
    if not np.isfinite(res.fun) or not np.isfinite(res.x) or (f0 - res.fun) < 1e-8 * max(1.0, abs(f0)):
        return float("nan")
    return float(res.x)


def inverse_reference(kind: int, lam: float, offset: float, max_value: float, y):
    """Elementwise inverse transformation, the reference's formulas and clamping rules
    (`src/transformations.jl:6-44` Box-Cox, `:145-146` percentage, `:149-150` positive). Vectorised numpy."""
    y = np.asarray(y, np.float64)
    if kind == KIND_IDENTITY:
        return y.copy()
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        if kind == KIND_POSITIVE:
            return np.maximum(np.exp(y) - offset, 0.0)
        if kind == KIND_PERCENTAGE:
            return np.maximum(logistic(y) * 100.0 - offset, 0.0)
        if kind != KIND_BOXCOX:
            raise AssertionError(f"Unknown transformation kind: {kind}")
        v = lam * y + 1.0
        if lam > 0:
            res = np.power(np.maximum(v, 1.0e-10), 1.0 / lam) - offset
        elif lam < 0:
            normal = np.power(np.where(v > 1.0e-10, v, 1.0), 1.0 / lam) - offset
            tiny = np.minimum(np.power(np.where(v > 0, v, 1.0), 1.0 / lam), 1000.0 * max_value) - offset
            res = np.where(v > 1.0e-10, normal, np.where(v <= 0, 0.0, tiny))
        else:
            res = np.exp(y) - offset
        return np.maximum(res, 0.0)


class _Transform:
    """A closure of the reference: callable on a scalar (returns a float) or a numpy array."""

    def __init__(self, fn: Callable, name: str):
        self._fn, self.__name__ = fn, name

    def __call__(self, y):
        out = self._fn(y)
        return float(out) if np.ndim(y) == 0 else out


class InverseTransform(_Transform):
    """Built-in inverse transformation; `spec = (kind, lambda, offset, max_value)` is what the device kernel takes."""

    def __init__(self, kind: int, lam: float = 0.0, offset: float = 0.0, max_value: float = 0.0):
        self.spec = (int(kind), float(lam), float(offset), float(max_value))
        super().__init__(lambda y: inverse_reference(*self.spec, y), f"inverse[{kind}]")


def get_transformations(transform_name: str, values) -> Tuple[Callable, Callable]:
    """`get_transformations(transform_name, values)` → `(forward_transform, inverse_transform)`
    (`src/transformations.jl:134-170`). Supported: "percentage", "positive", "boxcox"."""
    vals = np.asarray(values, np.float64)
    offset = _get_offset(vals)
    if transform_name == "percentage":
        log.info("Using percentage transformation")
        return (_Transform(lambda y: logit((np.asarray(y, np.float64) + offset) / 100.0), "percentage"),
                InverseTransform(KIND_PERCENTAGE, 0.0, offset))
    if transform_name == "positive":
        log.info("Using positive transformation with offset = %s", offset)
        return (_Transform(lambda y: np.log(np.asarray(y, np.float64) + offset), "positive"),
                InverseTransform(KIND_POSITIVE, 0.0, offset))
    if transform_name == "boxcox":
        max_value = float(vals.max())
        shifted = vals + offset
        lam = fit_boxcox_lambda(shifted)
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            transformed = boxcox(lam, shifted)
        bc_range = float(transformed.max() - transformed.min()) if np.all(np.isfinite(transformed)) else float("nan")
        log_range = float(np.log(shifted).max() - np.log(shifted).min())
        if not np.all(np.isfinite(transformed)) or bc_range <= 1.0e-2 * log_range:
            log.warning("Box-Cox transformation degenerate (lambda = %s, transformed range = %s); "
                        "falling back to log transformation (issue #51).", lam, bc_range)
            return get_transformations("positive", values)
        log.info("Using Box-Cox transformation with lambda = %s and offset = %s", lam, offset)
        fwd = _Transform(lambda y: boxcox(lam, np.asarray(y, np.float64) + offset), "boxcox")
        fwd.lam = lam
        return fwd, InverseTransform(KIND_BOXCOX, lam, offset, max_value)
    raise AssertionError(f"Unknown transform_name: {transform_name}")
