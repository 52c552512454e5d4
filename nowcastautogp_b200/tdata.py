"""`TData`, `create_transformed_data`, `create_nowcast_data` — the scenario containers of the
reference, same names, argument meaning and assertion behaviour.

Mirrors `/root/reference/src/TData.jl:46-74` and `/root/reference/src/create_nowcast_data.jl:27-40,71-76`.
Dates are `numpy.datetime64[D]` (or anything `np.asarray(..., "datetime64[D]")` accepts) in place of
Julia's `Date`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterable, List, Sequence

import numpy as np


def _as_dates(ds) -> np.ndarray:
    arr = np.asarray(list(ds) if not isinstance(ds, np.ndarray) else ds)
    if arr.dtype.kind == "M":
        return arr
    if arr.dtype.kind in "OUS":
        return arr.astype("datetime64[D]")
    return arr  # numeric time axis is allowed too (AutoGP accepts Real ds)


@dataclass(frozen=True, eq=False)
class TData:
    """Container for transformed time-series data: `ds`, `y` (transformed), `values` (raw).

    `TData(ds, values, transformation=f)` applies `f` elementwise (`src/TData.jl:55`), promotes `y`
    and `values` to a common floating type (`:58-61`) and asserts equal lengths (`:52`).
    """
    ds: np.ndarray
    y: np.ndarray
    values: np.ndarray

    def __init__(self, ds, values, *, transformation: Callable):
        ds = _as_dates(ds)
        vals = np.asarray(values)
        assert len(ds) == len(vals), "length of `ds` should match length of `values`"
        y = np.asarray([transformation(v) for v in vals.tolist()]) if len(vals) else np.asarray([], float)
        common = np.promote_types(y.dtype if y.size else vals.dtype, vals.dtype)
        object.__setattr__(self, "ds", ds)
        object.__setattr__(self, "y", y.astype(common))
        object.__setattr__(self, "values", vals.astype(common))

    def __len__(self) -> int:
        return len(self.ds)


def create_transformed_data(ds: Iterable, values: Iterable, *, transformation: Callable) -> TData:
    """Convenience constructor from any iterables (`src/TData.jl:72-74`)."""
    return TData(list(ds) if not isinstance(ds, np.ndarray) else ds,
                 list(values) if not isinstance(values, np.ndarray) else values,
                 transformation=transformation)


def create_nowcast_data(nowcasts, dates: Sequence, *, transformation: Callable = lambda y: y) -> List[TData]:
    """Vector-of-vectors or matrix (columns = scenarios) → `list[TData]`.

    `src/create_nowcast_data.jl:27-40` (vector method, three asserts) and `:71-76` (matrix method:
    `eachcol` then the vector method). All scenarios share the one `dates` vector — the fact that
    lets the device factor once per particle.
    """
    if isinstance(nowcasts, np.ndarray) and nowcasts.ndim == 2:
        nowcasts = [nowcasts[:, j] for j in range(nowcasts.shape[1])]
    nowcasts = list(nowcasts)
    assert all(len(v) == len(dates) for v in nowcasts), "Length of each nowcast must match length of dates"
    assert len(nowcasts) > 0, "nowcasts must not be empty"
    first_length = len(nowcasts[0])
    assert all(len(v) == first_length for v in nowcasts), "All vectors in nowcasts must have the same length"
    return [create_transformed_data(dates, nc, transformation=transformation) for nc in nowcasts]
