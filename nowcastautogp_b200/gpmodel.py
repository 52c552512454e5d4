"""`GPConfig` and `GPModel`: the particle-ensemble state the reference keeps inside `AutoGP.GPModel`.

The reference re-exports `AutoGP.GPModel` / `AutoGP.GP.GPConfig` (`/root/reference/src/NowcastAutoGP.jl:8-9`)
and drives it through 13 calls. This module is the host-side mirror of that object:
explicit arrays (kernel programs, unconstrained hyperparameters, noise, log-weights, time/target
transforms) instead of Gen traces, with the same method names as the AutoGP calls the reference
makes — `add_data`, `maybe_resample`, `mcmc_structure`, `mcmc_parameters`, `predict_mvn`,
`num_particles`, `to_dict` / `from_dict` (`Dict(model)` / `GPModel(dict)`). Kernel-structure and
parameter proposals stay on the CPU (BASELINE.json north_star); every likelihood, factorisation
and predictive moment is a batched call into libnagp.

[R] marks semantics recalled from AutoGP.jl that cannot be checked in this image (docs/KERNEL_SPEC.md).
"""
from __future__ import annotations

import copy
import functools
from dataclasses import dataclass, field
from functools import reduce
from math import gcd
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import kernels as kn
from .engine import Engine, PosDefError

_DEFAULT_ENGINES: Dict[int, Engine] = {}


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine for `device` (created on first use; raises without a GPU)."""
    if device not in _DEFAULT_ENGINES:
        _DEFAULT_ENGINES[device] = Engine(device)
    return _DEFAULT_ENGINES[device]


# ------------------------------------------------------------------------------------------------
@dataclass
class GPConfig:
    """Mirror of `AutoGP.GP.GPConfig` (fields and defaults as dumped at
    `/root/reference/docs/src/vignettes/setting-priors.md:228-245`)."""
    node_dist_leaf: Sequence[float] = (0.0, 1 / 3, 0.0, 1 / 3, 1 / 3)
    node_dist_nocp: Sequence[float] = tuple(np.array([0, 6, 0, 6, 6, 5, 5], float) / 28)
    node_dist_cp: Sequence[float] = tuple(np.array([0, 6, 0, 6, 6, 4, 4, 2], float) / 28)
    max_branch: int = 2
    max_depth: int = -1
    changepoints: bool = True
    noise: Optional[float] = None
    prior: dict = field(default_factory=lambda: {
        "gamma": {"mu": 0.0, "sigma": 1.0},        # gamma = 2 logistic(mu + sigma z) [R]
        "period": {"mu": -1.5, "sigma": 1.0},      # LogNormal [C: setting-priors.md:106-116]
        "wildcard": {"mu": -1.5, "sigma": 1.0},    # LogNormal for every other positive parameter [R]
    })
    cp_scale: float = 1e-3                          # fixed ChangePoint sharpness [R]
    # depth cap used when max_depth == -1 so that programs respect the wire-format limits
    hard_max_depth: int = 5

    Constant, Linear, SquaredExponential, GammaExponential, Periodic, Plus, Times, ChangePoint = range(1, 9)


@dataclass
class LinearTransform:
    """`x -> slope * x + intercept`; `fit` maps [min, max] of the data onto [lo, hi] [R]."""
    slope: float
    intercept: float

    @classmethod
    def fit(cls, x: np.ndarray, lo: float, hi: float) -> "LinearTransform":
        xmin, xmax = float(np.min(x)), float(np.max(x))
        if not xmax > xmin:
            # AutoGP divides by the range: a flat series gives a singular covariance (issue #51,
            # `/root/reference/src/make_and_fit_model.jl:6-8`)
            raise PosDefError(1)
        slope = (hi - lo) / (xmax - xmin)
        return cls(slope, lo - slope * xmin)

    def apply(self, x):
        return self.slope * np.asarray(x, np.float64) + self.intercept

    def unapply(self, x):
        return (np.asarray(x, np.float64) - self.intercept) / self.slope


def to_numeric(ds) -> np.ndarray:
    """Dates → float days (AutoGP uses Unix seconds [R]; the unit cancels in the [0,1] rescaling)."""
    arr = np.asarray(ds)
    if arr.dtype.kind == "M":
        return arr.astype("datetime64[D]").astype(np.int64).astype(np.float64)
    return arr.astype(np.float64)


def lag_grid(num: np.ndarray) -> Tuple[Optional[np.ndarray], float]:
    """Integer grid indices and grid step (in `num` units) when every time point is an integer
    multiple of a common step from the first point; (None, 0.0) otherwise."""
    if len(num) < 2 or not np.all(np.equal(np.mod(num, 1), 0)):
        return None, 0.0
    d = (num - num.min()).astype(np.int64)
    step = reduce(gcd, [int(v) for v in d if v > 0], 0)
    if step == 0:
        return None, 0.0
    g = d // step
    if g.max() > 8 * len(num) + 64:     # a table this sparse costs more than pairwise evaluation
        return None, 0.0
    return g.astype(np.int32), float(step)


# ------------------------------------------------------------------------------------------------
# hyperparameters live as standard-normal z; theta = transform(z)  [R]
def _logistic(x):
    return 1.0 / (1.0 + np.exp(-x))


def _norm_cdf(x):
    from math import erf, sqrt
    return 0.5 * (1.0 + erf(x / sqrt(2.0)))


def slot_transform(name: str, z: float, config: GPConfig) -> float:
    pr = config.prior
    z = float(min(max(z, -60.0), 60.0))      # a diverging HMC trajectory must not overflow exp(); it is rejected anyway
    if name == "period":
        return float(np.exp(pr["period"]["mu"] + pr["period"]["sigma"] * z))
    if name == "gamma":
        return float(2.0 * _logistic(pr["gamma"]["mu"] + pr["gamma"]["sigma"] * z))
    if name == "intercept":
        return float(z)
    if name == "location":
        return float(_norm_cdf(z))
    if name == "scale":
        return float(config.cp_scale)
    return float(np.exp(pr["wildcard"]["mu"] + pr["wildcard"]["sigma"] * z))


def slot_dtheta_dz(name: str, z: float, theta: float, config: GPConfig) -> float:
    """d theta / d z of `slot_transform` (chain rule from the device's constrained-space gradient)."""
    pr = config.prior
    if name == "period":
        return pr["period"]["sigma"] * theta
    if name == "gamma":
        return pr["gamma"]["sigma"] * theta * (1.0 - 0.5 * theta)
    if name == "intercept":
        return 1.0
    if name == "location":
        return float(np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi))
    if name == "scale":
        return 0.0
    return pr["wildcard"]["sigma"] * theta


# ---- the same two maps over whole slot vectors (every particle of a call at once) ---------------------------
_SLOT_CODE = {"period": 1, "gamma": 2, "intercept": 3, "location": 4, "scale": 5}      # 0: wildcard (log-normal)


@functools.lru_cache(maxsize=65536)
def slot_codes(prog: bytes) -> np.ndarray:
    """Slot kind of every theta slot of a program (cached per structure)."""
    out = np.array([_SLOT_CODE.get(nm, 0) for nm in kn.theta_slot_names(prog)], np.int8)
    out.setflags(write=False)
    return out


def transform_slots(codes: np.ndarray, z: np.ndarray, config: GPConfig) -> np.ndarray:
    """`slot_transform` over a vector of slots."""
    pr = config.prior
    z = np.clip(np.asarray(z, np.float64), -60.0, 60.0)
    th = np.exp(pr["wildcard"]["mu"] + pr["wildcard"]["sigma"] * z)
    m = codes == 1
    if m.any():
        th[m] = np.exp(pr["period"]["mu"] + pr["period"]["sigma"] * z[m])
    m = codes == 2
    if m.any():
        th[m] = 2.0 * _logistic(pr["gamma"]["mu"] + pr["gamma"]["sigma"] * z[m])
    m = codes == 3
    if m.any():
        th[m] = z[m]
    m = codes == 4
    if m.any():
        th[m] = [_norm_cdf(v) for v in z[m]]
    m = codes == 5
    if m.any():
        th[m] = config.cp_scale
    return th


def dtheta_dz_slots(codes: np.ndarray, z: np.ndarray, theta: np.ndarray, config: GPConfig) -> np.ndarray:
    """`slot_dtheta_dz` over a vector of slots."""
    pr = config.prior
    out = pr["wildcard"]["sigma"] * theta
    m = codes == 1
    if m.any():
        out[m] = pr["period"]["sigma"] * theta[m]
    m = codes == 2
    if m.any():
        out[m] = pr["gamma"]["sigma"] * theta[m] * (1.0 - 0.5 * theta[m])
    out[codes == 3] = 1.0
    m = codes == 4
    if m.any():
        out[m] = np.exp(-0.5 * z[m] * z[m]) / np.sqrt(2.0 * np.pi)
    out[codes == 5] = 0.0
    return out


def device_slot_spec(codes: np.ndarray, config: GPConfig):
    """The z -> theta map of every slot in the form `nagp_hmc` takes: (kind int32, a, b) per slot."""
    pr = config.prior
    kind = np.zeros(len(codes), np.int32)
    a = np.full(len(codes), float(pr["wildcard"]["mu"]))
    b = np.full(len(codes), float(pr["wildcard"]["sigma"]))
    m = codes == 1
    a[m], b[m] = pr["period"]["mu"], pr["period"]["sigma"]
    m = codes == 2
    kind[m], a[m], b[m] = 2, pr["gamma"]["mu"], pr["gamma"]["sigma"]
    kind[codes == 3] = 3
    kind[codes == 4] = 4
    m = codes == 5
    kind[m], a[m] = 5, float(config.cp_scale)
    return kind, a, b


def device_noise_spec(config: GPConfig):
    if config.noise is not None:
        return 5, float(config.noise), 0.0
    return 0, float(config.prior["wildcard"]["mu"]), float(config.prior["wildcard"]["sigma"])


HMC_DEFAULT = {"n_leapfrog": 10, "eps": 0.02}     # Gen.hmc's L = 10 [R]; step sized for N(0,1)-scaled z
GRAD_KERNEL_MAX_N = 232                            # nagp_logml_grad / nagp_hmc keep a chain's factor in shared memory


@dataclass
class Particle:
    """One SMC particle: kernel structure + unconstrained hyperparameters."""
    prog: bytes
    z: np.ndarray          # one entry per theta slot
    noise_z: float

    def theta(self, config: GPConfig) -> List[float]:
        return transform_slots(slot_codes(self.prog), self.z, config).tolist()

    def noise(self, config: GPConfig) -> float:
        if config.noise is not None:
            return float(config.noise)
        return slot_transform("noise", self.noise_z, config)

    def copy(self) -> "Particle":
        return Particle(self.prog, self.z.copy(), float(self.noise_z))

    @property
    def n_nodes(self) -> int:
        return len(self.prog)


def _draw_code(rng, probs) -> int:
    """1 + a categorical draw from unnormalised weights (inverse CDF on one uniform)."""
    cum = np.cumsum(np.asarray(probs, float))
    return 1 + int(min(np.searchsorted(cum, rng.random() * cum[-1], side="right"), len(cum) - 1))


def sample_structure(rng, config: GPConfig, depth: int = 1) -> List[int]:
    """Post-order opcode list drawn from the PCFG prior (`node_dist_*`)."""
    max_depth = config.max_depth if config.max_depth > 0 else config.hard_max_depth
    if depth >= max_depth:
        return [_draw_code(rng, config.node_dist_leaf)]
    code = _draw_code(rng, config.node_dist_cp if config.changepoints else config.node_dist_nocp)
    if code <= kn.OP_PERIODIC:
        return [code]
    return sample_structure(rng, config, depth + 1) + sample_structure(rng, config, depth + 1) + [code]


def sample_particle(rng, config: GPConfig) -> Particle:
    while True:
        prog = bytes(sample_structure(rng, config))
        if len(prog) <= kn.MAX_PROG:
            break
    nslots = len(kn.theta_slot_names(prog))
    return Particle(prog, rng.standard_normal(nslots), float(rng.standard_normal()))


def subtree_slices(prog: bytes) -> List[Tuple[int, int]]:
    """(start, end) of the sub-tree rooted at every node of a post-order program."""
    out, stack = [], []
    for i, op in enumerate(prog):
        if op <= kn.OP_PERIODIC:
            stack.append(i)
        else:
            stack.pop()
            start = stack.pop()
            stack.append(start)
        out.append((stack[-1], i + 1))
    return out


def pack_particles(particles: Sequence[Particle], config: GPConfig) -> kn.FlatEnsemble:
    """All particles of a call as one flat ensemble; the z -> theta map runs once over the concatenated slots."""
    progs = [p.prog for p in particles]
    prog_off = np.zeros(len(progs) + 1, np.int64)
    theta_off = np.zeros(len(progs) + 1, np.int64)
    np.cumsum([len(p) for p in progs], out=prog_off[1:])
    np.cumsum([len(p.z) for p in particles], out=theta_off[1:])
    codes = np.concatenate([slot_codes(p.prog) for p in particles]) if particles else np.zeros(0, np.int8)
    z = np.concatenate([p.z for p in particles]) if particles else np.zeros(0)
    assert len(codes) == len(z), "particle with a z vector that does not match its structure"
    if config.noise is not None:
        noise = np.full(len(progs), float(config.noise))
    else:
        noise = transform_slots(np.zeros(len(progs), np.int8), np.array([p.noise_z for p in particles]), config)
    return kn.FlatEnsemble(np.frombuffer(b"".join(progs), np.uint8).copy(), prog_off,
                           transform_slots(codes, z, config), theta_off, noise)


class _FlatChains:
    """The particles of one call laid out for vector arithmetic: concatenated slot vectors and their offsets."""

    def __init__(self, particles: Sequence[Particle], config: GPConfig):
        self.P = len(particles)
        progs = [p.prog for p in particles]
        self.prog = np.frombuffer(b"".join(progs), np.uint8).copy()
        self.prog_off = np.zeros(self.P + 1, np.int64)
        self.off = np.zeros(self.P + 1, np.int64)
        np.cumsum([len(p) for p in progs], out=self.prog_off[1:])
        np.cumsum([len(p.z) for p in particles], out=self.off[1:])
        self.codes = np.concatenate([slot_codes(p.prog) for p in particles]) if self.P else np.zeros(0, np.int8)
        self.Z = np.concatenate([p.z for p in particles]) if self.P else np.zeros(0)
        self.NZ = np.array([p.noise_z for p in particles], np.float64)
        self.noise_codes = np.zeros(self.P, np.int8)
        self._seg = np.repeat(np.arange(self.P), np.diff(self.off))

    def segsum(self, v: np.ndarray) -> np.ndarray:
        """Per-particle sums of a slot vector (particles without slots give 0)."""
        return np.bincount(self._seg, weights=v, minlength=self.P)


# ------------------------------------------------------------------------------------------------
class MixtureMVN:
    """What `AutoGP.predict_mvn` returns (`MixtureModel{MvNormal}`): per-particle mean and Cholesky
    factor in original units plus normalised weights. `rand` = `rand(dist, n)` / `rand(dist)`
    (`/root/reference/src/forecasting.jl:47,67`), drawn on the device from host-supplied normals."""

    def __init__(self, engine: Engine, logw: np.ndarray, mu: np.ndarray, L: np.ndarray):
        self.engine, self.logw, self.mu, self.L = engine, logw, mu, L

    @property
    def weights(self) -> np.ndarray:
        w = np.exp(self.logw - self.logw.max())
        return w / w.sum()

    def rand(self, n: Optional[int] = None, rng=None) -> np.ndarray:
        rng = np.random.default_rng() if rng is None else rng
        D = 1 if n is None else int(n)
        h = self.mu.shape[-1]
        zeta = rng.standard_normal((1, D, h))
        u = rng.uniform(size=(1, D))
        x, _, _ = self.engine.draw(self.logw[None, :], self.mu[None], self.L[None], zeta, u=u)
        x = np.ascontiguousarray(x)
        return x[:, 0] if n is None else x


class GPModel:
    """Host mirror of `AutoGP.GPModel(ds, y; n_particles, config)` (`src/make_and_fit_model.jl:84`)."""

    def __init__(self, ds, y, *, n_particles: int = 8, config: Optional[GPConfig] = None, rng=None,
                 engine: Optional[Engine] = None, _state: Optional[dict] = None):
        self.engine = engine
        if _state is not None:
            self.__dict__.update(_state)
            return
        self.config = GPConfig() if config is None else config
        self.rng = np.random.default_rng() if rng is None else rng
        self.ds = np.asarray(ds)
        self.y = np.asarray(y, np.float64)
        assert len(self.ds) == len(self.y)
        num = to_numeric(self.ds)
        self.ds_transform = LinearTransform.fit(num, 0.0, 1.0)        # [C] time window → [0,1]
        self.y_transform = LinearTransform.fit(self.y, -1.0, 1.0)     # [R] target range → [-1,1]
        self.particles = [sample_particle(self.rng, self.config) for _ in range(n_particles)]
        self.log_weights = np.zeros(n_particles)
        self.n_obs = 0                      # observations already absorbed into the weights (SMC)
        self.obs_order = np.arange(len(self.y))
        self._logml = np.zeros(n_particles)  # log marginal likelihood of the first n_obs observations

    # ---- plumbing --------------------------------------------------------------------------------
    def _engine(self) -> Engine:
        if self.engine is None:
            self.engine = default_engine()
        return self.engine

    def num_particles(self) -> int:
        return len(self.particles)

    def _times(self, ds_all) -> Tuple[np.ndarray, Optional[np.ndarray], float]:
        num = to_numeric(ds_all)
        g, step = lag_grid(num)
        return self.ds_transform.apply(num), g, step * self.ds_transform.slope

    def ensemble(self) -> kn.FlatEnsemble:
        return pack_particles(self.particles, self.config)

    def kernels(self) -> List[kn.Node]:
        """The particles' kernels as DSL trees (AutoGP: `covariance_kernels(model)`)."""
        return [kn.unflatten(p.prog, p.theta(self.config)) for p in self.particles]

    # ---- Dict(model) / GPModel(dict): src/forecasting.jl:128,133 -------------------------------------
    def to_dict(self) -> dict:
        return {
            "ds": self.ds.astype(str).tolist() if self.ds.dtype.kind == "M" else self.ds.tolist(),
            "ds_is_date": self.ds.dtype.kind == "M",
            "y": self.y.tolist(),
            "ds_transform": [self.ds_transform.slope, self.ds_transform.intercept],
            "y_transform": [self.y_transform.slope, self.y_transform.intercept],
            "config": copy.deepcopy(self.config.__dict__),
            "particles": [{"prog": list(p.prog), "z": p.z.tolist(), "noise_z": p.noise_z} for p in self.particles],
            "log_weights": self.log_weights.tolist(),
            "n_obs": int(self.n_obs),
            "obs_order": self.obs_order.tolist(),
            "logml": self._logml.tolist(),
        }

    @classmethod
    def from_dict(cls, d: dict, *, engine: Optional[Engine] = None, rng=None) -> "GPModel":
        d = copy.deepcopy(d)
        ds = np.asarray(d["ds"], "datetime64[D]") if d.get("ds_is_date") else np.asarray(d["ds"])
        state = dict(
            config=GPConfig(**d["config"]), rng=np.random.default_rng() if rng is None else rng,
            ds=ds, y=np.asarray(d["y"], np.float64),
            ds_transform=LinearTransform(*d["ds_transform"]), y_transform=LinearTransform(*d["y_transform"]),
            particles=[Particle(bytes(p["prog"]), np.asarray(p["z"], np.float64), float(p["noise_z"]))
                       for p in d["particles"]],
            log_weights=np.asarray(d["log_weights"], np.float64), n_obs=int(d["n_obs"]),
            obs_order=np.asarray(d["obs_order"], np.int64), _logml=np.asarray(d["logml"], np.float64))
        return cls(None, None, engine=engine, _state=state)

    # ---- batched likelihood: the primitive fit_smc! and the MCMC moves call ---------------------------
    def logml(self, particles: Sequence[Particle], idx: np.ndarray) -> np.ndarray:
        """log p(y[idx] | particle) for every particle in one device call; -inf where the Gram is
        not positive definite."""
        if len(idx) == 0:
            return np.zeros(len(particles))
        t, g, step = self._times(self.ds[idx])
        ens = pack_particles(particles, self.config)
        lm, info = self._engine().logml_batch(ens, t, self.y_transform.apply(self.y[idx]), g=g, step=step)
        return np.where(info == 0, lm, -np.inf)

    # ---- AutoGP.fit_smc!: src/make_and_fit_model.jl:91 --------------------------------------------------
    def fit_smc(self, *, schedule: Sequence[int], n_mcmc: int, n_hmc: int, shuffle: bool = True,
                biased: bool = False, adaptive_rejuvenation: bool = False, hmc_config=None,
                verbose: bool = False, ess_fraction: float = 0.5, obs_order=None) -> None:
        """Data-annealed SMC [R]: for each cumulative count in `schedule` absorb the next batch of
        observations (one batched device logML call for all particles), resample when the ESS
        drops below `ess_fraction`·P (AutoGP's adaptive default, `docs/vignettes/setting-priors.jl:
        174-175`), then rejuvenate with `n_mcmc` structure moves × `n_hmc` parameter steps.
        `n_mcmc` and `n_hmc` are required keywords, as in AutoGP (`test/test_gpconfig.jl:37-43`)."""
        n = len(self.y)
        self._hmc_config = hmc_config
        if obs_order is not None:           # series fitted in lockstep share one order, so their requests coalesce
            self.obs_order = np.asarray(obs_order, np.int64)
            assert sorted(self.obs_order.tolist()) == list(range(n)), "obs_order must be a permutation of the observations"
        else:
            self.obs_order = self.rng.permutation(n) if shuffle else np.arange(n)
        # Steps after which no particle changed (no resampling, no rejuvenation: n_mcmc == 0, or adaptive rejuvenation
        # that did not fire) extend the particles' Cholesky factors by the new observations instead of re-factoring:
        # the factors stay on the device in an appendable store (`nagp_factor_store_large` / `nagp_factor_append`,
        # rank-append of the new rows: the stored factor is read once). Observations are kept in arrival order
        # (`obs_order`, never sorted: the row order of a Gram matrix is arbitrary), on the lag grid of the whole series.
        use_store = (n_mcmc == 0 or adaptive_rejuvenation) and not getattr(self._engine(), "is_coalescing", False)
        store = None                        # (Factor, failed-particle mask): valid while the particles are unchanged
        t_all = g_all = None
        step_all = 0.0
        self.append_steps = 0               # schedule steps served by a rank-append (diagnostic, used by the tests)
        try:
            for step in schedule:
                step = int(min(step, n))
                if use_store and t_all is None:
                    t_all, g_all, step_all = self._times(self.ds)
                    y_all = self.y_transform.apply(self.y)
                if use_store and store is not None and step > self.n_obs and store[0].n == self.n_obs:
                    new_idx = self.obs_order[self.n_obs:step]
                    dl, _, info = self._engine().factor_append(store[0], t_all[new_idx], y_all[new_idx],
                                                               g_new=None if g_all is None else g_all[new_idx], check=False)
                    bad = store[1] | (info != 0)
                    store = (store[0], bad)
                    with np.errstate(invalid="ignore"):
                        new = np.where(bad, -np.inf, self._logml + dl)
                    self.append_steps += 1
                elif use_store:
                    if store is not None:
                        store[0].free()
                    idx = self.obs_order[:step]
                    f = self._engine().factor_store_large(pack_particles(self.particles, self.config), t_all[idx], y_all[idx],
                                                          capacity=n, g=None if g_all is None else g_all[idx], step=step_all,
                                                          check=False)
                    store = (f, f.info != 0)
                    new = np.where(f.info == 0, f.logml_n, -np.inf)
                else:
                    new = self.logml(self.particles, np.sort(self.obs_order[:step]))
                if not np.any(np.isfinite(new)):
                    raise PosDefError(1)
                with np.errstate(invalid="ignore"):
                    self.log_weights = self.log_weights + np.where(np.isfinite(self._logml), new - self._logml, 0.0)
                self._logml = new
                self.n_obs = step
                resampled = self.maybe_resample(ess_fraction * self.num_particles())
                rejuvenate = (not adaptive_rejuvenation or resampled) and n_mcmc > 0
                if not adaptive_rejuvenation or resampled:
                    self.mcmc_structure(n_mcmc, n_hmc)
                if (resampled or rejuvenate) and store is not None:     # the particles changed: their factors are stale
                    store[0].free()
                    store = None
                if verbose:
                    print(f"fit_smc: {step}/{n} observations, ESS {self.effective_sample_size():.2f}")
        finally:
            if store is not None:
                store[0].free()

    # ---- AutoGP.add_data!: src/forecasting.jl:135 -----------------------------------------------------
    def add_data(self, ds, y) -> None:
        ds, y = np.asarray(ds), np.asarray(y, np.float64)
        assert len(ds) == len(y)
        assert self.n_obs == len(self.y), "add_data! on a model that has not absorbed its own data"
        old = self._logml
        self.ds = np.concatenate([self.ds, ds.astype(self.ds.dtype)])
        self.y = np.concatenate([self.y, y])
        self.obs_order = np.concatenate([self.obs_order, np.arange(self.n_obs, len(self.y))])
        new = self.logml(self.particles, np.arange(len(self.y)))
        # a particle that left the fit with weight -inf (its Gram was not positive definite) carries no mass: only
        # the particles that still count must factor
        alive = np.isfinite(self.log_weights)
        if not np.all(np.isfinite(new[alive])) or not alive.any():
            raise PosDefError(1)
        with np.errstate(invalid="ignore"):
            self.log_weights = np.where(alive, self.log_weights + (new - old), -np.inf)     # log w += logML(m) - logML(n) [R]
        self._logml = new
        self.n_obs = len(self.y)

    # ---- AutoGP.maybe_resample!: src/forecasting.jl:138-141 ---------------------------------------------
    def effective_sample_size(self) -> float:
        ess, _ = self._engine().ess(self.log_weights[None, :])
        return float(ess[0])

    def maybe_resample(self, ess_threshold: float) -> bool:
        if not self.effective_sample_size() < ess_threshold:
            return False
        w = np.exp(self.log_weights - self.log_weights.max())
        parents = self.rng.choice(len(w), size=len(w), p=w / w.sum())   # multinomial [R]
        self.particles = [self.particles[a].copy() for a in parents]
        self._logml = self._logml[parents]
        self.log_weights = np.zeros(len(w))
        return True

    # ---- rejuvenation moves (proposals on the CPU, likelihoods batched on the device) -------------------
    def _obs_idx(self) -> np.ndarray:
        return np.sort(self.obs_order[:self.n_obs])

    def _logpost_grad(self, particles: Sequence[Particle], idx: np.ndarray):
        """log p(y[idx] | particle) + log N(z; 0, I) and its gradient in unconstrained space, all particles in
        one `nagp_logml_grad` call. Returns (lp [P], list of dz arrays, dnoise_z [P])."""
        flat = _FlatChains(particles, self.config)
        lp, dZ, dNZ = self._logpost_grad_flat(flat, flat.Z, flat.NZ, self._grid_of(idx))
        return lp, [dZ[flat.off[i]:flat.off[i + 1]] for i in range(len(particles))], dNZ

    def _grid_of(self, idx: np.ndarray):
        t, g, step = self._times(self.ds[idx])
        return t, g, step, self.y_transform.apply(self.y[idx])

    def _logpost_grad_flat(self, flat: "_FlatChains", Z: np.ndarray, NZ: np.ndarray, grid):
        """The same over concatenated slot vectors: Z [sum slots], NZ [P] -> (lp [P], dZ, dNZ)."""
        cfg = self.config
        t, g, step, y = grid
        theta = transform_slots(flat.codes, Z, cfg)
        noise = (np.full(flat.P, float(cfg.noise)) if cfg.noise is not None
                 else transform_slots(flat.noise_codes, NZ, cfg))
        ens = kn.FlatEnsemble(flat.prog, flat.prog_off, theta, flat.off, noise)
        if len(y) > GRAD_KERNEL_MAX_N or getattr(self, "_force_fd_gradient", False):
            lm, gth, gnz, info = self._logml_grad_by_differences(ens, flat, t, g, step, y)
        else:
            lm, gth, gnz, info = self._engine().logml_grad(ens, t, y, g=g, step=step)
            lm, gth, gnz, info = lm[0], gth[0], gnz[0], info[0]
        lp = np.where(info == 0, lm, -np.inf)
        dZ = np.where(np.isfinite(gth), gth, 0.0) * dtheta_dz_slots(flat.codes, Z, theta, cfg) - Z
        dZ = np.where(np.isfinite(dZ), dZ, 0.0)
        lp = lp - 0.5 * flat.segsum(Z * Z)
        dNZ = np.zeros(flat.P)
        if cfg.noise is None:
            dNZ = np.where(np.isfinite(gnz), gnz, 0.0) * dtheta_dz_slots(flat.noise_codes, NZ, noise, cfg) - NZ
            dNZ = np.where(np.isfinite(dNZ), dNZ, 0.0)
            lp = lp - 0.5 * NZ ** 2
        return lp, dZ, dNZ

    def _logml_grad_by_differences(self, ens, flat: "_FlatChains", t, g, step, y):
        """logML and its gradient in the constrained parameters for series beyond the gradient kernels' size
        (n > 232: they keep a chain's factor in shared memory): central differences of the DEVICE log marginal likelihood,
        every perturbation of every particle in one batched call through the large-path kernel (scenario s of the batch
        perturbs slot (s - 1) // 2 of each particle by ±h; the particles are independent, so one scenario moves them
        all). 2·(slots per particle) + 3 factorisations per particle instead of one factorisation and one inverse, but
        the leapfrog integrator only needs a deterministic force field to stay reversible and volume-preserving, and the
        accept step uses the exact logML: the chain targets the same posterior as with the analytic gradient."""
        P, off = flat.P, np.asarray(flat.off)
        ns = np.diff(off)
        D = int(ns.max()) if P else 0
        learn_noise = self.config.noise is None
        K = 1 + 2 * D + (2 if learn_noise else 0)
        theta, noise = np.asarray(ens.theta, np.float64), np.asarray(ens.noise, np.float64)
        TH, NZ = np.tile(theta, (K, 1)), np.tile(noise, (K, 1))
        h_th = 1e-6 * np.maximum(1.0, np.abs(theta))
        h_nz = 1e-6 * np.maximum(1.0, np.abs(noise))
        for j in range(D):
            sel = off[:-1][ns > j] + j
            TH[1 + 2 * j, sel] += h_th[sel]
            TH[2 + 2 * j, sel] -= h_th[sel]
        if learn_noise:
            NZ[1 + 2 * D] += h_nz
            NZ[2 + 2 * D] -= h_nz
        n = len(y)
        lm = np.empty((K, P))
        r = self._engine().forecast_instances(ens, n, 0, 0, t, y, np.empty((K, 0)), np.zeros(P), 1.0, 0.0, g=g, step=step,
                                              theta=TH, noise=NZ, K=K, logml_m=lm, want_moments=False)
        info = np.asarray(r["info"])
        lm = np.where(info == 0, lm, np.nan)
        gth = np.zeros_like(theta)
        for j in range(D):
            sel = off[:-1][ns > j] + j
            pj = np.nonzero(ns > j)[0]
            gth[sel] = (lm[1 + 2 * j, pj] - lm[2 + 2 * j, pj]) / ((TH[1 + 2 * j, sel] - TH[2 + 2 * j, sel]))
        gnz = np.zeros(P)
        if learn_noise:
            gnz = (lm[1 + 2 * D] - lm[2 + 2 * D]) / (NZ[1 + 2 * D] - NZ[2 + 2 * D])
        return lm[0], gth, gnz, info[0]

    def mcmc_parameters(self, n_hmc: int, hmc_config: Optional[dict] = None) -> float:
        """`AutoGP.mcmc_parameters!(model, n_hmc)` (`src/forecasting.jl:148,65`): `n_hmc` Hamiltonian Monte
        Carlo steps on the unconstrained hyperparameters of every particle (N(0,1) prior on z [R]), all
        particles advanced together in ONE device call (`nagp_hmc`): the leapfrog integrator, the z -> theta
        maps and the accept/reject step run on the device, the host only supplies the momenta and the uniforms
        (drawn in the order the host integrator `_mcmc_parameters_host` draws them, so both walk the same
        chain). Beyond the gradient kernels' size limit (n > 232) the integrator runs on the host over gradients taken as
        central differences of the device logML (`_logml_grad_by_differences`): same target, same move. Returns the
        acceptance rate."""
        from .engine import NagpError
        hc = dict(HMC_DEFAULT, **(hmc_config or getattr(self, "_hmc_config", None) or {}))
        if not hc.get("device", True) or self.n_obs > GRAD_KERNEL_MAX_N:
            return self._mcmc_parameters_host(n_hmc, hmc_config)
        idx = self._obs_idx()
        P = len(self.particles)
        if len(idx) == 0 or n_hmc <= 0:
            return 0.0
        cfg = self.config
        learn_noise = cfg.noise is None
        flat = _FlatChains(self.particles, cfg)
        t, g, step, y = self._grid_of(idx)
        total = len(flat.Z)
        mom, mnz, logu = np.empty((n_hmc, total)), np.empty((n_hmc, P)), np.empty((n_hmc, P))
        for it in range(n_hmc):
            mom[it] = self.rng.standard_normal(total)
            if learn_noise:
                mnz[it] = self.rng.standard_normal(P)
            logu[it] = np.log(self.rng.uniform(size=P))
        kind, sa, sb = device_slot_spec(flat.codes, cfg)
        try:
            Z, NZ, lm, nacc, info = self._engine().hmc(
                flat.prog, flat.prog_off, flat.off, kind, sa, sb, device_noise_spec(cfg), flat.Z[None, :].copy(),
                flat.NZ[None, :].copy(), t, y, g=g, step=step, n_leapfrog=int(hc["n_leapfrog"]), eps=float(hc["eps"]),
                momenta=mom, noise_momenta=mnz if learn_noise else None, log_u=logu)
        except NagpError as e:
            if e.code != -4:
                raise
            return self._metropolis_parameters(n_hmc)
        Z, NZ, lm, nacc = Z[0], NZ[0], lm[0], nacc[0]
        for i in np.nonzero(nacc > 0)[0]:
            self.particles[i] = Particle(self.particles[i].prog, Z[flat.off[i]:flat.off[i + 1]].copy(), float(NZ[i]))
            self._logml[i] = lm[i]
        return float(nacc.sum()) / max(1, n_hmc * P)

    def _mcmc_parameters_host(self, n_hmc: int, hmc_config: Optional[dict] = None) -> float:
        """The same move with the integrator on the host: each leapfrog stage is one `nagp_logml_grad` call.
        Kept as the cross-check of `nagp_hmc` (`hmc_config={"device": False}`)."""
        from .engine import NagpError
        hc = dict(HMC_DEFAULT, **(hmc_config or getattr(self, "_hmc_config", None) or {}))
        L, eps = int(hc["n_leapfrog"]), float(hc["eps"])
        idx = self._obs_idx()
        P = len(self.particles)
        if len(idx) == 0 or n_hmc <= 0:
            return 0.0
        learn_noise = self.config.noise is None
        flat = _FlatChains(self.particles, self.config)
        grid = self._grid_of(idx)
        Z, NZ = flat.Z, flat.NZ
        try:
            lp, dZ, dNZ = self._logpost_grad_flat(flat, Z, NZ, grid)
        except NagpError as e:
            if e.code != -4:
                raise
            return self._metropolis_parameters(n_hmc)
        acc = 0
        for _ in range(n_hmc):
            mom = self.rng.standard_normal(len(Z))
            mnz = self.rng.standard_normal(P) if learn_noise else np.zeros(P)
            h0 = -lp + 0.5 * flat.segsum(mom * mom) + 0.5 * mnz ** 2
            Zq, NZq, g_z, g_n, lp_new = Z, NZ, dZ, dNZ, lp
            for _l in range(L):
                mom = mom + 0.5 * eps * g_z
                Zq = Zq + eps * mom
                if learn_noise:
                    mnz = mnz + 0.5 * eps * g_n
                    NZq = NZq + eps * mnz
                lp_new, g_z, g_n = self._logpost_grad_flat(flat, Zq, NZq, grid)
                mom = mom + 0.5 * eps * g_z
                if learn_noise:
                    mnz = mnz + 0.5 * eps * g_n
            h1 = -lp_new + 0.5 * flat.segsum(mom * mom) + 0.5 * mnz ** 2
            accept = np.log(self.rng.uniform(size=P)) < np.where(np.isfinite(h1), h0 - h1, -np.inf)
            acc += int(accept.sum())
            slot_acc = np.repeat(accept, np.diff(flat.off))
            Z, dZ = np.where(slot_acc, Zq, Z), np.where(slot_acc, g_z, dZ)
            NZ, dNZ, lp = np.where(accept, NZq, NZ), np.where(accept, g_n, dNZ), np.where(accept, lp_new, lp)
        prior = -0.5 * flat.segsum(Z * Z) - (0.5 * NZ ** 2 if learn_noise else 0.0)
        moved = np.array([not np.array_equal(Z[flat.off[i]:flat.off[i + 1]], p.z) or NZ[i] != p.noise_z
                          for i, p in enumerate(self.particles)])
        for i in np.nonzero(moved)[0]:
            self.particles[i] = Particle(self.particles[i].prog, Z[flat.off[i]:flat.off[i + 1]].copy(), float(NZ[i]))
            self._logml[i] = lp[i] - prior[i]
        return acc / max(1, n_hmc * P)

    def _metropolis_parameters(self, n_steps: int, step_size: float = 0.15) -> float:
        """Gaussian random-walk Metropolis on all unconstrained hyperparameters (same target as the HMC move)."""
        idx = self._obs_idx()
        P = len(self.particles)
        acc = 0
        for _ in range(n_steps):
            props = []
            for p in self.particles:
                q = p.copy()
                q.z = p.z + step_size * self.rng.standard_normal(len(p.z))
                if self.config.noise is None:
                    q.noise_z = p.noise_z + step_size * self.rng.standard_normal()
                props.append(q)
            lm_new = self.logml(props, idx)
            for i, (p, q) in enumerate(zip(self.particles, props)):
                lp = -0.5 * (q.z @ q.z + q.noise_z ** 2) + 0.5 * (p.z @ p.z + p.noise_z ** 2)
                if np.log(self.rng.uniform()) < lm_new[i] - self._logml[i] + lp:
                    self.particles[i] = q
                    self._logml[i] = lm_new[i]
                    acc += 1
        return acc / max(1, n_steps * P)

    def mcmc_structure(self, n_mcmc: int, n_hmc: int) -> float:
        """`AutoGP.mcmc_structure!(model, n_mcmc, n_hmc)` (`src/forecasting.jl:146`): `n_mcmc` rounds of
        a sub-tree regeneration Metropolis move (replace a uniformly chosen sub-tree by a fresh prior
        draw; the prior terms cancel, leaving the likelihood ratio times the node-count ratio), each
        followed by `n_hmc` parameter steps. AutoGP's own involutive move set differs in detail [R];
        like it, the proposals run on the CPU and only the likelihoods go to the device."""
        idx = self._obs_idx()
        acc = 0
        for _ in range(n_mcmc):
            props = []
            for p in self.particles:
                slices = subtree_slices(p.prog)
                node = int(self.rng.integers(len(slices)))
                s0, s1 = slices[node]
                names = kn.theta_slot_names(p.prog)
                z0 = len(kn.theta_slot_names(p.prog[:s0]))
                z1 = len(kn.theta_slot_names(p.prog[:s1]))
                depth = 1 + sum(1 for (a, b) in slices if a <= s0 and b >= s1 and (a, b) != (s0, s1))
                sub = bytes(sample_structure(self.rng, self.config, depth))
                prog = p.prog[:s0] + sub + p.prog[s1:]
                if len(prog) > kn.MAX_PROG:
                    props.append(p.copy())
                    continue
                zsub = self.rng.standard_normal(len(kn.theta_slot_names(sub)))
                props.append(Particle(prog, np.concatenate([p.z[:z0], zsub, p.z[z1:]]), p.noise_z))
                assert len(props[-1].z) == len(kn.theta_slot_names(prog)) and len(names) == len(p.z)
            lm_new = self.logml(props, idx)
            for i, (p, q) in enumerate(zip(self.particles, props)):
                log_alpha = lm_new[i] - self._logml[i] + np.log(p.n_nodes) - np.log(q.n_nodes)
                if np.log(self.rng.uniform()) < log_alpha:
                    self.particles[i] = q
                    self._logml[i] = lm_new[i]
                    acc += 1
            if n_hmc > 0:
                self.mcmc_parameters(n_hmc)
        return acc / max(1, n_mcmc * len(self.particles))

    # ---- AutoGP.predict_mvn: src/forecasting.jl:46,66 ---------------------------------------------------
    def predict_mvn(self, forecast_dates, noise_pred: Optional[float] = None) -> MixtureMVN:
        fd = np.asarray(forecast_dates).astype(self.ds.dtype)
        idx = self._obs_idx()
        t, g, step = self._times(np.concatenate([self.ds[idx], fd]))
        ens = self.ensemble()
        eng = self._engine()
        f = eng.factor_store(ens, len(idx), 0, len(fd), t, self.y_transform.apply(self.y[idx]), self.log_weights,
                             self.y_transform.slope, self.y_transform.intercept, g=g, step=step,
                             noise_pred=-1.0 if noise_pred is None else float(noise_pred), check=False)
        # PosDefException only for particles that carry weight (MvNormal's constructor would never see the others)
        alive = np.isfinite(self.log_weights)
        bad = np.asarray(f.info) != 0
        if (bad & alive).any() or not alive.any():
            f.free()
            raise PosDefError(int(np.asarray(f.info)[bad & alive][0]) if (bad & alive).any() else 1)
        mu, L = eng.predict(f)
        f.free()
        if bad.any():       # zero-weight particles: finite placeholders, never drawn from
            mu = np.where(bad[:, None], 0.0, mu)
            L = np.where(bad[:, None, None], np.eye(L.shape[-1])[None], L)
        return MixtureMVN(eng, self.log_weights.copy(), mu, L)
