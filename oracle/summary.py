"""CPU restatement of the forecast summary: the reference's built-in inverse transformations
and Julia's default `quantile`.

TEST INFRASTRUCTURE ONLY — imported by tests/ (and bench.py's CPU-baseline leg), never by the product package.
Scalar, branch-for-branch restatements (small inputs only):
* `inverse_scalar` follows `/root/reference/src/transformations.jl:6-44` (`_inv_boxcox`), `:145-146`
  (percentage: `max(logistic(y) * 100 - offset, 0)`) and `:149-150` (positive: `max(exp(y) - offset, 0)`).
  `logistic` is LogExpFunctions' (not vendored; restated [R]): `exp(x) / (1 + exp(x))`, 0 below
  -744.4400719213812 and 1 above 36.7368005696771.
* `quantile_type7` follows Statistics.jl `quantile(v, p)` with its defaults alpha = beta = 1 (not vendored;
  restated [R]) — pinned in tests/test_summary.py against numpy's `quantile(..., method="linear")`, which is
  the same Hyndman-Fan type 7 definition.
"""
from __future__ import annotations

import math
from typing import Sequence

KIND_IDENTITY, KIND_POSITIVE, KIND_PERCENTAGE, KIND_BOXCOX = 0, 1, 2, 3


def _exp(x: float) -> float:
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


def logistic(x: float) -> float:
    if x < -744.4400719213812:
        return 0.0
    if x > 36.7368005696771:
        return 1.0
    e = _exp(x)
    return e / (1.0 + e)


def _pow(b: float, e: float) -> float:
    try:
        return math.pow(b, e)
    except OverflowError:
        return math.inf


def inv_boxcox(lam: float, offset: float, max_value: float, y: float) -> float:
    v = lam * y + 1.0
    if lam > 0:
        result = _pow(max(v, 1.0e-10), 1.0 / lam) - offset
    elif lam < 0:
        if v > 1.0e-10:
            result = _pow(v, 1.0 / lam) - offset
        elif v <= 0:
            result = 0.0
        else:
            result = min(_pow(v, 1.0 / lam), 1000.0 * max_value) - offset
    else:
        result = _exp(y) - offset
    return max(result, 0.0)


def inverse_scalar(kind: int, lam: float, offset: float, max_value: float, y: float) -> float:
    if kind == KIND_IDENTITY:
        return y
    if kind == KIND_POSITIVE:
        return max(_exp(y) - offset, 0.0)
    if kind == KIND_PERCENTAGE:
        return max(logistic(y) * 100.0 - offset, 0.0)
    if kind == KIND_BOXCOX:
        return inv_boxcox(lam, offset, max_value, y)
    raise AssertionError(kind)


def quantile_type7(values: Sequence[float], p: float) -> float:
    v = sorted(values)
    n = len(v)
    if n == 1:
        return v[0]
    aleph = n * p + (1.0 - p)
    j = min(max(int(math.trunc(aleph)), 1), n - 1)
    gamma = min(max(aleph - j, 0.0), 1.0)
    a, b = v[j - 1], v[j]
    if math.isfinite(a) and math.isfinite(b):
        return a + gamma * (b - a)
    return (1.0 - gamma) * a + gamma * b
