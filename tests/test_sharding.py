"""Host-side multi-GPU logic on CPU: the (series, scenario) partition and the final gather, over a
world_size-2 `gloo` process group (the GPU box runs the same code over NCCL). The compute step is a
deterministic stand-in — the kernels are covered by the `-m gpu` parity tests."""
import os
import socket

import numpy as np
import pytest

from nowcastautogp_b200.sharding import Slice, partition, sharded_fit, sharded_forecast


def _pairs(parts):
    return [(s.series, k) for p in parts for s in p for k in range(s.k0, s.k1)]


@pytest.mark.parametrize("S,K,world", [(53, 1000, 8), (1, 1000, 8), (1, 7, 8), (3, 5, 2), (5, [3, 0, 4, 1, 2], 2),
                                       (2, [1, 9], 4), (8, 10, 8), (1, 1, 1)])
def test_partition_covers_every_pair_once(S, K, world):
    parts = partition(S, K, world)
    ks = [K] * S if np.isscalar(K) else K
    want = [(s, k) for s in range(S) for k in range(ks[s])]
    assert len(parts) == world
    assert _pairs(parts) == want            # exactly once each, global order preserved rank by rank
    if S >= world:                          # whole series per rank, sizes differ by at most one series
        assert all(sl.k0 == 0 and sl.k1 == ks[sl.series] for p in parts for sl in p)
        per_rank = [len({sl.series for sl in p} | set()) for p in parts]
        n_nonempty = sum(1 for k in ks if k > 0)
        assert sum(per_rank) == n_nonempty
    else:                                   # near-equal pair counts
        cnt = [sum(sl.k1 - sl.k0 for sl in p) for p in parts]
        assert max(cnt) - min(cnt) <= 1


@pytest.mark.parametrize("S,K,world", [(53, 1000, 8), (5, [3, 0, 7, 1, 4], 2), (9, 10, 4)])
def test_partition_split_series_is_even_and_complete(S, K, world):
    """`split_series=True` (per-scenario regime): the pair list is cut into near-equal contiguous runs even when there
    are more series than ranks; every pair is still owned exactly once, in global order."""
    parts = partition(S, K, world, split_series=True)
    ks = [K] * S if np.isscalar(K) else list(K)
    counts = [sum(sl.k1 - sl.k0 for sl in p) for p in parts]
    assert sum(counts) == sum(ks) and max(counts) - min(counts) <= 1
    seen = [(sl.series, k) for p in parts for sl in p for k in range(sl.k0, sl.k1)]
    assert seen == [(s, k) for s in range(S) for k in range(ks[s])]


def test_partition_c4_shape():
    parts = partition(53, 1000, 8)
    assert [len(p) for p in parts] == [7, 7, 7, 7, 7, 6, 6, 6]


def _fake_compute(h, D, P):
    def compute(sl: Slice):
        kk = sl.k1 - sl.k0
        cols = np.arange(sl.k0 * D, sl.k1 * D)
        x = 1000.0 * sl.series + cols[None, :] + 0.01 * np.arange(h)[:, None]
        lw = -1.0 * sl.series - np.arange(sl.k0, sl.k1)[:, None] * 0.5 - 0.001 * np.arange(P)[None, :]
        assert x.shape == (h, kk * D)
        return x, lw
    return compute


def _expected(S, ks, h, D, P):
    comp = _fake_compute(h, D, P)
    dr, lw = {}, {}
    for s in range(S):
        dr[s], lw[s] = comp(Slice(s, 0, ks[s]))
    return dr, lw


def test_sharded_forecast_single_process():
    S, ks, h, D, P = 3, [4, 2, 5], 3, 2, 4
    dr, lw = sharded_forecast(_fake_compute(h, D, P), S, ks, h, D, P)
    edr, elw = _expected(S, ks, h, D, P)
    for s in range(S):
        assert np.array_equal(dr[s], edr[s]) and np.array_equal(lw[s], elw[s])


def _worker(rank, world, port, S, ks, h, D, P, q, split=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def compute(sl):
            calls.append(sl)
            return _fake_compute(h, D, P)(sl)

        dr, lw = sharded_forecast(compute, S, ks, h, D, P, split_series=split)
        edr, elw = _expected(S, ks, h, D, P)
        ok = all(np.array_equal(dr[s], edr[s]) and np.array_equal(lw[s], elw[s]) for s in range(S))
        q.put((rank, ok, [(c.series, c.k0, c.k1) for c in calls]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S,ks,split", [(3, [4, 2, 5], False), (1, [7], False), (3, [4, 2, 5], True)])
def test_sharded_forecast_gloo_world2(S, ks, split):
    import torch.multiprocessing as mp
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    h, D, P, world = 3, 2, 4, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, ks, h, D, P, q, split)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)                      # every rank holds the full, correctly ordered result
    owned = sorted(c for _, _, calls in res for c in calls)
    want = sorted((sl.series, sl.k0, sl.k1) for part in partition(S, ks, world, split) for sl in part)
    assert owned == want                                     # each rank computed only its own slices


def _fit_worker(rank, world, port, S, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seen = []

        def fit_local(idx):
            seen.extend(idx)
            return [{"series": s, "payload": [s * 1.5, "z" * s]} for s in idx]

        out = sharded_fit(fit_local, S)
        q.put((rank, [d["series"] for d in out] == list(range(S)) and out[S - 1]["payload"] == [(S - 1) * 1.5, "z" * (S - 1)],
               seen))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S", [5, 2, 1])
def test_sharded_fit_gloo_world2(S):
    import torch.multiprocessing as mp
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_fit_worker, args=(r, world, port, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)                       # every rank holds all S models in series order
    owned = sorted(s for _, _, seen in res for s in seen)
    assert owned == list(range(S))                           # each series fitted by exactly one rank


def test_sharded_fit_single_process():
    out = sharded_fit(lambda idx: [{"s": s} for s in idx], 4)
    assert [d["s"] for d in out] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        sharded_fit(lambda idx: [], 2)


class _StubModel:
    """Just enough of a GPModel for forecast_with_nowcasts_sharded: a generator, a particle count, an engine."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def num_particles(self):
        return 2

    def _engine(self):
        return None


def _rng_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nowcastautogp_b200 import api

        def fake(base_model, nowcasts, dates, D, *, n_mcmc=0, n_hmc=0, ess_threshold=0.0, forecast_n_hmc=None, rng=None):
            kk, h = len(nowcasts), len(dates)
            return rng.standard_normal((h, kk * D)), np.zeros((kk, 2))      # pure noise: what the generator supplies

        api._forecast_with_nowcasts = fake
        models = [_StubModel(7)]                 # the same seed on every rank, as models rebuilt from gathered dicts have
        draws, _ = api.forecast_with_nowcasts_sharded(models, [[None] * 8], [0, 1, 2], 4)
        q.put((rank, draws[0]))
    finally:
        dist.destroy_process_group()


def test_sharded_forecast_slices_use_independent_generators():
    """One series split over two ranks: the ranks' column blocks must not repeat each other's normals (identically
    seeded model generators on every rank would make them identical), and every rank must hold the same full matrix."""
    import torch.multiprocessing as mp
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rng_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0], res[1]) and res[0].shape == (3, 32)
    first, second = res[0][:, :16], res[0][:, 16:]          # scenarios 0-3 (rank 0) and 4-7 (rank 1), D = 4
    assert not np.array_equal(first, second)
    assert abs(np.corrcoef(first.ravel(), second.ravel())[0, 1]) < 0.5
