"""Synthetic workloads of SURVEY.md §8(d): weekly-hospitalisation-shaped series, kernel trees and
hyperparameters sampled from AutoGP's default prior, and multiplicative nowcast scenarios.

Series: `/root/reference/docs/vignettes/setting-priors.jl:92-98`; scenarios:
`/root/reference/docs/vignettes/getting-started.jl:505`; structure prior:
`/root/reference/docs/src/vignettes/setting-priors.md:92-97,238-240`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import kernels as kn

# AutoGP default PCFG weights over node codes 1..8 (setting-priors.md:238-240)
NODE_DIST_LEAF = np.array([0.0, 1 / 3, 0.0, 1 / 3, 1 / 3])
NODE_DIST_NOCP = np.array([0, 6, 0, 6, 6, 5, 5], float) / 28
NODE_DIST_CP = np.array([0, 6, 0, 6, 6, 4, 4, 2], float) / 28
PRIOR_MU, PRIOR_SIGMA = -1.5, 1.0   # LogNormal for period [C] / wildcard [R]
CP_SCALE = 1e-3                      # [R] fixed ChangePoint sharpness


def weekly_series(n: int, seed: int, period: float = 52.0) -> Tuple[np.ndarray, np.ndarray]:
    """(day offsets, raw values): exp(log 50 + sin(2πt/52) + 0.02 t + 0.15 ε), weekly dates."""
    rng = np.random.default_rng(seed)
    w = np.arange(n, dtype=np.float64)
    raw = np.exp(np.log(50.0) + np.sin(2 * np.pi * w / period) + 0.02 * w * (52.0 / period)
                 + 0.15 * rng.standard_normal(n))
    return w * 7.0, raw


def _lognormal(rng) -> float:
    return float(np.exp(PRIOR_MU + PRIOR_SIGMA * rng.standard_normal()))


def sample_leaf(rng, code: int) -> kn.Node:
    if code == kn.OP_CONSTANT:
        return kn.Constant(_lognormal(rng))
    if code == kn.OP_LINEAR:
        return kn.Linear(float(rng.uniform(0.0, 1.0)), _lognormal(rng), _lognormal(rng))
    if code == kn.OP_SQEXP:
        return kn.SquaredExponential(_lognormal(rng), _lognormal(rng))
    if code == kn.OP_GAMMAEXP:
        gamma = 2.0 / (1.0 + np.exp(-rng.standard_normal()))
        return kn.GammaExponential(_lognormal(rng), float(gamma), _lognormal(rng))
    if code == kn.OP_PERIODIC:
        return kn.Periodic(_lognormal(rng), _lognormal(rng), _lognormal(rng))
    raise ValueError(code)


def sample_tree(rng, max_depth: int = 4, changepoints: bool = True, depth: int = 1,
                node_dist_leaf=NODE_DIST_LEAF) -> kn.Node:
    """One draw from the PCFG structure prior, depth-capped for benchmarking."""
    if depth >= max_depth:
        code = 1 + int(rng.choice(5, p=node_dist_leaf))
        return sample_leaf(rng, code)
    dist = NODE_DIST_CP if changepoints else NODE_DIST_NOCP
    code = 1 + int(rng.choice(len(dist), p=dist))
    if code <= kn.OP_PERIODIC:
        return sample_leaf(rng, code)
    left = sample_tree(rng, max_depth, changepoints, depth + 1, node_dist_leaf)
    right = sample_tree(rng, max_depth, changepoints, depth + 1, node_dist_leaf)
    if code == kn.OP_PLUS:
        return kn.Plus(left, right)
    if code == kn.OP_TIMES:
        return kn.Times(left, right)
    return kn.ChangePoint(left, right, float(rng.uniform(0.0, 1.0)), CP_SCALE)


def sample_ensemble(rng, P: int, max_depth: int = 4, changepoints: bool = True
                    ) -> Tuple[List[kn.Node], np.ndarray]:
    trees = [sample_tree(rng, max_depth, changepoints) for _ in range(P)]
    noise = np.array([_lognormal(rng) for _ in range(P)])
    return trees, noise


@dataclass
class Workload:
    """One series' forecast_with_nowcasts inputs in device-ready (scaled) form."""
    n: int
    k: int
    h: int
    t: np.ndarray          # [q] rescaled time (train window → [0,1])
    g: np.ndarray          # [q] int32 grid indices (weekly grid)
    step: float
    y1: np.ndarray         # [n] scaled training targets
    y2: np.ndarray         # [K,k] scaled scenario targets
    ya: float
    yb: float
    trees: List[kn.Node]
    noise: np.ndarray
    logw0: np.ndarray
    ens: kn.FlatEnsemble


def make_workload(n: int, k: int, h: int, K: int, P: int, seed: int, max_depth: int = 4,
                  period: float = 52.0) -> Workload:
    """Series + prior-sampled ensemble + K multiplicative nowcast scenarios, log-transformed
    ("positive" transformation) and scaled the way AutoGP scales its inputs."""
    rng = np.random.default_rng(seed)
    days, raw = weekly_series(n, seed + 1, period)
    q = n + k + h
    g = np.arange(q, dtype=np.int32)
    slope_t = 1.0 / (days[-1] - days[0])
    step = 7.0 * slope_t
    t = (np.arange(q) * 7.0 - days[0]) * slope_t
    ylog = np.log(raw)
    ya = 2.0 / (ylog.max() - ylog.min())          # LinearTransform(y, -1, 1) [R]
    yb = -1.0 - ya * ylog.min()
    y1 = ya * ylog + yb
    last = raw[-1]
    scen = last * np.exp(0.1 + 0.027 * rng.standard_normal((K, k)))
    y2 = ya * np.log(scen) + yb
    trees, noise = sample_ensemble(rng, P, max_depth)
    logw0 = np.log(rng.dirichlet(np.ones(P) * 2.0))
    return Workload(n, k, h, t, g, step, y1, y2, float(ya), float(yb), trees, noise, logw0,
                    kn.pack_ensemble(trees, noise))


def perturbed_theta(ens: kn.FlatEnsemble, K: int, seed: int, rel: float = 0.05
                    ) -> Tuple[np.ndarray, np.ndarray]:
    """Per-(scenario, particle) hyperparameters, as they look after per-scenario HMC rejuvenation
    (`n_hmc > 0`, `/root/reference/src/forecasting.jl:145-149`): multiplicative log-normal jitter on
    every positive slot; signs/zeros preserved. Returns theta [K, total] and noise [K, P]."""
    rng = np.random.default_rng(seed)
    names = kn.theta_slot_names(ens.prog.tobytes())
    fixed = np.array([nm in ("scale",) for nm in names])          # ChangePoint sharpness is not learned
    gamma = np.array([nm == "gamma" for nm in names])
    fac = np.exp(rel * rng.standard_normal((K, len(ens.theta))))
    fac[:, fixed] = 1.0
    th = ens.theta[None, :] * fac
    th[:, gamma] = np.minimum(th[:, gamma], 2.0)                    # GammaExponential is PSD only for γ ≤ 2
    nz = ens.noise[None, :] * np.exp(rel * rng.standard_normal((K, ens.size)))
    return np.ascontiguousarray(th), np.ascontiguousarray(nz)
