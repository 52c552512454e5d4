"""Multi-GPU partition of the (series, nowcast scenario) work and the final gather.

The instances of the hot path are independent per (series, scenario, particle): the reference runs
one task per scenario on a private model copy (`/root/reference/src/forecasting.jl:131-133`) and one
`forecast_with_nowcasts` call per series (`docs/vignettes/getting-started.jl:540-552`). The only
couplings are the per-scenario weight normalisation over that scenario's P particles and the final
`hcat` (`forecasting.jl:166`). So: one process per GPU, all particles of a (series, scenario) pair on
one GPU, pairs split contiguously across ranks — series first, so that in the scenario-shared fast
path each particle is factored on exactly one GPU — and ONE collective at the end: a single all-gather of a
packed buffer holding the draws `[h, K·D]` and the log-weights `[K, P]` (NCCL over NVLink on the GPU box, gloo in CPU tests).
No collective on the data path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class Slice:
    series: int
    k0: int     # first scenario (inclusive)
    k1: int     # last scenario (exclusive)


def partition(n_series: int, n_scenarios: Sequence[int] | int, world: int, split_series: bool = False) -> List[List[Slice]]:
    """Contiguous split of the flattened (series, scenario) pairs over `world` ranks.

    Whole series are kept together whenever there are at least as many series as ranks (C4: 53 series
    on 8 GPUs → 7/7/7/7/7/6/6/6); otherwise — or always with `split_series=True` — the pair list is cut into
    `world` near-equal contiguous runs, which splits a series' scenarios across ranks (C2: one series, K scenarios
    → K/world each; C4 with `split_series`: 6625 pairs per rank instead of 7000/6000). Splitting costs nothing when
    every scenario is its own piece of work (per-scenario hyperparameters: `n_hmc > 0`); in the scenario-shared
    regime it would factor the particles of a boundary series on two GPUs, hence the default.
    Every pair is owned by exactly one rank; a rank's slices are in global order."""
    if world < 1:
        raise ValueError("world must be >= 1")
    ks = [int(n_scenarios)] * n_series if np.isscalar(n_scenarios) else [int(k) for k in n_scenarios]
    if len(ks) != n_series:
        raise ValueError("n_scenarios must have one entry per series")
    out: List[List[Slice]] = [[] for _ in range(world)]
    if n_series >= world and not split_series:
        base, extra = divmod(n_series, world)
        s = 0
        for r in range(world):
            cnt = base + (1 if r < extra else 0)
            out[r] = [Slice(i, 0, ks[i]) for i in range(s, s + cnt) if ks[i] > 0]
            s += cnt
        return out
    total = sum(ks)
    bounds = [(total * r) // world for r in range(world + 1)]
    starts = np.concatenate([[0], np.cumsum(ks)])
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        for i in range(n_series):
            a, b = max(lo, starts[i]), min(hi, starts[i + 1])
            if a < b:
                out[r].append(Slice(i, int(a - starts[i]), int(b - starts[i])))
    return out


_PACKED: Dict[tuple, tuple] = {}     # (cap, D, h, P, world, device) -> (send, recv) packed buffers, reused across calls


def _packed_buffers(cap: int, D: int, h: int, P: int, world: int, dev):
    import torch
    key = (cap, D, h, P, world, str(dev))
    if key not in _PACKED:
        n = cap * (D * h + P)
        _PACKED[key] = (torch.zeros(n, dtype=torch.float64, device=dev),
                        torch.zeros((world, n), dtype=torch.float64, device=dev))
    return _PACKED[key]


def sharded_forecast(compute: Callable, n_series: int, n_scenarios: Sequence[int] | int, h: int, D: int, P: int,
                     group=None, device=None, in_place: bool = False, split_series: bool = False):
    """Run `compute` on this rank's slices and gather with ONE collective.

    Host form (default): `compute(slice) -> (x [h, (k1-k0)·D], logw [(k1-k0), P])` as NumPy arrays; returns
    `(draws, logw)` dicts of NumPy arrays keyed by series on EVERY rank: `draws[s]` is the reference's `(h, K_s·D)`
    matrix in scenario-major column order, `logw[s]` is `[K_s, P]`.
    Device form (`in_place=True`): `compute(slice, x_out, lw_out)` fills `x_out [(k1-k0)·D, h]` (the column-major
    `(h, (k1-k0)·D)` block, i.e. what `nagp_draw` writes) and `lw_out [(k1-k0), P]`, both views of the packed send
    buffer on `device`; nothing bounces through the host and the dicts hold device tensors (`draws[s]` a
    `(h, K_s·D)` transposed view).
    Either way every rank's results travel in one packed buffer `[pairs·D·h draws | pairs·P log-weights]` and one
    `all_gather_into_tensor` (NCCL over NVLink on the GPU box; the list form of `all_gather` under gloo). Works
    without an initialised process group (world = 1). Ranks may own different numbers of pairs (padded to the
    largest). `split_series`: see `partition`."""
    import torch
    import torch.distributed as dist

    use_dist = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if use_dist else 1
    rank = dist.get_rank(group) if use_dist else 0
    ks = [int(n_scenarios)] * n_series if np.isscalar(n_scenarios) else [int(k) for k in n_scenarios]
    parts = partition(n_series, ks, world, split_series)
    mine = parts[rank]
    n_pairs = [sum(s.k1 - s.k0 for s in p) for p in parts]
    cap = max(max(n_pairs), 1)
    dev = torch.device("cpu") if device is None else torch.device(device)
    send, recv = _packed_buffers(cap, D, h, P, world, dev)
    sx = send[:cap * D * h].view(cap, D, h)          # pair-major draws: the column blocks of x, transposed
    sl_ = send[cap * D * h:].view(cap, P)
    off = 0
    for sl in mine:
        kk = sl.k1 - sl.k0
        if in_place:
            compute(sl, sx[off:off + kk].view(kk * D, h), sl_[off:off + kk])
        else:
            x, logw = compute(sl)
            x = np.asarray(x, np.float64)
            if x.shape != (h, kk * D):
                raise ValueError(f"compute returned draws of shape {x.shape}, expected {(h, kk * D)}")
            sx[off:off + kk] = torch.from_numpy(np.ascontiguousarray(x.T)).view(kk, D, h).to(dev)
            sl_[off:off + kk] = torch.from_numpy(np.ascontiguousarray(np.asarray(logw, np.float64).reshape(kk, P))).to(dev)
        off += kk
    if use_dist:
        if dev.type == "cpu":     # gloo: the list form is available on every build
            dist.all_gather(list(recv.unbind(0)), send, group=group)
        else:
            dist.all_gather_into_tensor(recv, send, group=group)
        got = recv
    else:
        got = send[None]
    gx = got[:, :cap * D * h].view(world, cap, D, h)
    gl = got[:, cap * D * h:].view(world, cap, P)
    if in_place:
        draws, logws = {}, {}
        whole = {sl.series for p in parts for sl in p if sl.k0 == 0 and sl.k1 == ks[sl.series]}
        for s in range(n_series):
            if s not in whole:
                draws[s] = torch.empty((ks[s] * D, h), dtype=torch.float64, device=dev)
                logws[s] = torch.empty((ks[s], P), dtype=torch.float64, device=dev)
        for r in range(world):
            off = 0
            for sl in parts[r]:
                kk = sl.k1 - sl.k0
                if sl.series in whole:        # one rank owns the whole series: a view of the gathered buffer
                    draws[sl.series] = gx[r, off:off + kk].view(kk * D, h)
                    logws[sl.series] = gl[r, off:off + kk]
                else:
                    draws[sl.series][sl.k0 * D:sl.k1 * D] = gx[r, off:off + kk].view(kk * D, h)
                    logws[sl.series][sl.k0:sl.k1] = gl[r, off:off + kk]
                off += kk
        return {s: x.t() for s, x in draws.items()}, logws
    if got.is_cuda:
        # one device->host copy of the gathered buffer into pinned memory (cached per shape): pageable .cpu() of the 61 MB
        # of C4 was a third of the call
        hkey = ("host",) + tuple(got.shape)
        if hkey not in _PACKED:
            _PACKED[hkey] = (torch.empty(tuple(got.shape), dtype=torch.float64, pin_memory=True),)
        host = _PACKED[hkey][0]
        host.copy_(got, non_blocking=True)
        torch.cuda.current_stream(got.device).synchronize()
        gxn = host[:, :cap * D * h].reshape(world, cap, D, h).numpy()
        gln = host[:, cap * D * h:].reshape(world, cap, P).numpy()
    else:
        gxn, gln = gx.numpy(), gl.numpy()
    draws = {s: np.empty((h, ks[s] * D)) for s in range(n_series)}
    logws = {s: np.empty((ks[s], P)) for s in range(n_series)}
    for r in range(world):
        off = 0
        for sl in parts[r]:
            kk = sl.k1 - sl.k0
            draws[sl.series][:, sl.k0 * D:sl.k1 * D] = gxn[r, off:off + kk].reshape(kk * D, h).T
            logws[sl.series][sl.k0:sl.k1] = gln[r, off:off + kk]
            off += kk
    return draws, logws


def sharded_fit(fit_local: Callable[[List[int]], List[dict]], n_series: int, group=None) -> List[dict]:
    """Fit `n_series` independent series across the ranks of the current process group: rank r runs
    `fit_local(series indices)` -> one serialised model (`GPModel.to_dict()`) per index on its contiguous share
    (the same series-first split as the forecast: 53 series on 8 GPUs -> 7/7/7/7/7/6/6/6) and the dicts are
    exchanged with one `all_gather_object` (a few KB per model: structures, z, weights — no bulk data). Returns the
    `n_series` dicts in series order on every rank. No collective on the data path."""
    import torch.distributed as dist

    use_dist = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if use_dist else 1
    rank = dist.get_rank(group) if use_dist else 0
    parts = partition(n_series, 1, world)
    mine = [sl.series for sl in parts[rank]]
    local = fit_local(mine) if mine else []
    if len(local) != len(mine):
        raise ValueError("fit_local must return one model per series index")
    if use_dist:
        gathered: List = [None] * world
        dist.all_gather_object(gathered, list(zip(mine, local)), group=group)
    else:
        gathered = [list(zip(mine, local))]
    out: List = [None] * n_series
    for chunk in gathered:
        for s, d in chunk:
            out[s] = d
    if any(d is None for d in out):
        raise RuntimeError("sharded_fit: a series was not fitted by any rank")
    return out
