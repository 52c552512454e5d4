#!/usr/bin/env python
"""Generates tests/golden/*.json — the fixtures that pin the CPU oracle (oracle/) and, through it,
the CUDA path. Run from the repo root: `python tests/golden/make_golden.py`.

The reference's arithmetic lives in AutoGP.jl, which is neither vendored under /root/reference nor
runnable here (no Julia) — so these vectors are NOT outputs of the reference ("parity unpinned",
DESIGN.md §3). They are produced by implementations independent of oracle/nagp_oracle.c:

  kernel_values.json   every DSL node evaluated with mpmath at 50 digits from the formulas of
                       docs/KERNEL_SPEC.md §3, plus scikit-learn's ExpSineSquared / RBF /
                       DotProduct / ConstantKernel for the primitives whose published form coincides
  logml_closed.json    log marginal likelihoods with a closed form: Constant kernel (rank-1 +
                       sigma^2 I, Sherman-Morrison) and Linear kernel (rank-2, Woodbury) in mpmath
  logml_sklearn.json   scikit-learn GaussianProcessRegressor.log_marginal_likelihood on composite
                       kernels (Sum/Product of the coinciding primitives), LAPACK dpotrf underneath
  reference_fixtures.json  the series and scenarios the reference's own tests use
                       (/root/reference/test/test_nowcast_functions.jl:29-41,285,
                       /root/reference/test/test_model_fitting.jl:7-9) with mpmath log marginal
                       likelihoods / predictive moments of a fixed kernel on them
  draws.json           mixture draws mu_c + L_c zeta evaluated with exact rational arithmetic rounded
                       per fused-multiply-add (the bit-exact contract of KERNEL_SPEC §7)
"""
import json
import os
from fractions import Fraction

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 50


# ---- mpmath restatement of KERNEL_SPEC §3 (independent of the C oracle and of gram_np) ----------
def k_mp(node, t, u):
    kind = node[0]
    t, u = mp.mpf(t), mp.mpf(u)
    d = abs(t - u)
    if kind == "const":
        return mp.mpf(node[1])
    if kind == "lin":
        c, b, a = map(mp.mpf, node[1:])
        return b + a * (t - c) * (u - c)
    if kind == "se":
        l, a = map(mp.mpf, node[1:])
        return a * mp.e ** (-(d / l) ** 2 / 2)
    if kind == "ge":
        l, g, a = map(mp.mpf, node[1:])
        return a * mp.e ** (-((d / l) ** g)) if d > 0 else a
    if kind == "per":
        l, p, a = map(mp.mpf, node[1:])
        return a * mp.e ** (-2 * mp.sin(mp.pi * d / p) ** 2 / l ** 2)
    if kind == "plus":
        return k_mp(node[1], t, u) + k_mp(node[2], t, u)
    if kind == "times":
        return k_mp(node[1], t, u) * k_mp(node[2], t, u)
    if kind == "cp":
        loc, sc = mp.mpf(node[3]), mp.mpf(node[4])
        st = (1 + mp.tanh((t - loc) / sc)) / 2
        su = (1 + mp.tanh((u - loc) / sc)) / 2
        return (1 - st) * (1 - su) * k_mp(node[1], t, u) + st * su * k_mp(node[2], t, u)
    raise ValueError(kind)


OPS = {"const": 1, "lin": 2, "se": 3, "ge": 4, "per": 5, "plus": 6, "times": 7, "cp": 8}


def flatten(node):
    kind = node[0]
    if kind in ("plus", "times"):
        pl, tl = flatten(node[1]); pr, tr = flatten(node[2])
        return pl + pr + [OPS[kind]], tl + tr
    if kind == "cp":
        pl, tl = flatten(node[1]); pr, tr = flatten(node[2])
        return pl + pr + [8], tl + tr + [float(node[3]), float(node[4])]
    return [OPS[kind]], [float(x) for x in node[1:]]


def chol_mp(K):
    n = K.rows
    L = mp.zeros(n)
    for j in range(n):
        s = K[j, j] - sum(L[j, p] ** 2 for p in range(j))
        L[j, j] = mp.sqrt(s)
        for i in range(j + 1, n):
            L[i, j] = (K[i, j] - sum(L[i, p] * L[j, p] for p in range(j))) / L[j, j]
    return L


def gp_mp(node, noise, jitter, t_train, y, t_star=None, noise_pred=None):
    n = len(t_train)
    ts = list(t_train) + (list(t_star) if t_star is not None else [])
    q = len(ts)
    K = mp.matrix(q, q)
    for i in range(q):
        for j in range(q):
            K[i, j] = k_mp(node, ts[i], ts[j])
        K[i, i] += mp.mpf(noise if (i < n or noise_pred is None) else noise_pred) + mp.mpf(jitter)
    L = chol_mp(K)
    yv = mp.matrix([mp.mpf(v) for v in y])
    z = mp.lu_solve(L[:n, :n], yv)
    logml = -(n * mp.log(2 * mp.pi) + 2 * sum(mp.log(L[i, i]) for i in range(n)) + sum(v ** 2 for v in z)) / 2
    out = {"logml": float(logml)}
    if t_star is not None:
        h = q - n
        mu = [float(sum(L[n + r, c] * z[c] for c in range(n))) for r in range(h)]
        L33 = [[float(L[n + r, n + c]) if c <= r else 0.0 for c in range(h)] for r in range(h)]
        out.update(mu=mu, L=L33)
    return out


def kernel_values():
    trees = {
        "constant": ("const", 0.7),
        "linear": ("lin", 0.3, 0.2, 1.7),
        "sqexp": ("se", 0.4, 1.3),
        "gammaexp": ("ge", 0.35, 1.4, 0.9),
        "gammaexp2": ("ge", 0.4, 2.0, 1.3),      # == SquaredExponential only up to the 1/2: pins the form
        "periodic": ("per", 0.8, 0.33, 1.1),
        "plus": ("plus", ("lin", 0.1, 0.3, 0.5), ("per", 1.2, 0.25, 0.6)),
        "times": ("times", ("ge", 0.5, 0.8, 1.0), ("per", 0.9, 0.5, 2.0)),
        "changepoint": ("cp", ("lin", 0.0, 0.1, 1.0), ("per", 0.7, 0.2, 0.8), 0.45, 0.05),
        "deep": ("plus", ("times", ("lin", 0.2, 0.4, 0.9), ("ge", 0.3, 1.1, 0.7)),
                 ("cp", ("const", 0.3), ("plus", ("se", 0.2, 0.5), ("per", 1.0, 0.125, 0.4)), 0.6, 0.01)),
    }
    pts = [(0.0, 0.0), (0.1, 0.1), (0.0, 1.0), (0.25, 0.75), (0.44, 0.46), (0.46, 0.44), (0.9, 1.3), (1.02, 0.013),
           (0.6, 0.6000001), (0.3333333333333333, 0.7142857142857143)]
    cases = []
    for name, tr in trees.items():
        prog, theta = flatten(tr)
        cases.append({"name": name, "prog": prog, "theta": theta,
                      "points": [list(p) for p in pts],
                      "values": [float(k_mp(tr, *p)) for p in pts]})
    # scikit-learn cross-check for the primitives whose published forms coincide
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, DotProduct, ExpSineSquared
    X = np.array([[0.0], [0.13], [0.5], [0.77], [1.0], [1.21]])
    sk = []
    l, p, a = 0.8, 0.33, 1.1
    sk.append({"name": "periodic~ExpSineSquared", "prog": [5], "theta": [l, p, a], "x": X[:, 0].tolist(),
               "gram": (a * ExpSineSquared(length_scale=l, periodicity=p)(X)).tolist()})
    l, a = 0.4, 1.3
    sk.append({"name": "sqexp~RBF", "prog": [3], "theta": [l, a], "x": X[:, 0].tolist(),
               "gram": (a * RBF(length_scale=l)(X)).tolist()})
    # DotProduct(sigma0)(x, x') = sigma0^2 + x.x'  == Linear(intercept 0, bias sigma0^2, amplitude 1)
    s0 = 0.6
    sk.append({"name": "linear~DotProduct", "prog": [2], "theta": [0.0, s0 * s0, 1.0], "x": X[:, 0].tolist(),
               "gram": DotProduct(sigma_0=s0)(X).tolist()})
    sk.append({"name": "constant~ConstantKernel", "prog": [1], "theta": [0.7], "x": X[:, 0].tolist(),
               "gram": ConstantKernel(0.7)(X).tolist()})
    json.dump({"mpmath": cases, "sklearn": sk}, open(os.path.join(HERE, "kernel_values.json"), "w"), indent=1)


def logml_closed():
    """Constant v: K = v 11^T + s I  =>  logdet = (n-1) log s + log(s + n v);
    y^T K^-1 y = (y.y - v (1.y)^2 / (s + n v)) / s. Linear (c,b,a): K = U diag(b,a) U^T + s I, U = [1, t-c]
    via Woodbury / matrix determinant lemma on the 2x2 capacitance matrix."""
    rng = np.random.default_rng(20261018)
    cases = []
    for n in (1, 2, 7, 30, 150):
        t = np.linspace(0.0, 1.0, n) if n > 1 else np.array([0.0])
        y = rng.standard_normal(n)
        v, noise, jit = 0.37, 0.05, 1e-5
        s = mp.mpf(noise) + mp.mpf(jit)
        ym = [mp.mpf(float(x)) for x in y]
        yy, sy = sum(x * x for x in ym), sum(ym)
        logdet = (n - 1) * mp.log(s) + mp.log(s + n * mp.mpf(v))
        quad = (yy - mp.mpf(v) * sy ** 2 / (s + n * mp.mpf(v))) / s
        cases.append({"name": f"constant_n{n}", "prog": [1], "theta": [v], "noise": noise, "jitter": jit,
                      "t": t.tolist(), "y": y.tolist(),
                      "logml": float(-(n * mp.log(2 * mp.pi) + logdet + quad) / 2)})
        if n >= 2:
            c, b, a = 0.3, 0.2, 1.7
            U = mp.matrix(n, 2)
            for i in range(n):
                U[i, 0] = 1; U[i, 1] = mp.mpf(float(t[i])) - mp.mpf(c)
            Dinv = mp.matrix([[1 / mp.mpf(b), 0], [0, 1 / mp.mpf(a)]])
            C = Dinv + U.T * U / s                      # capacitance
            yv = mp.matrix(ym)
            Uty = U.T * yv
            quad = (yy - (Uty.T * mp.lu_solve(C, Uty))[0] / s) / s
            logdet = n * mp.log(s) + mp.log(mp.det(C)) + mp.log(mp.mpf(b) * mp.mpf(a))
            cases.append({"name": f"linear_n{n}", "prog": [2], "theta": [c, b, a], "noise": noise, "jitter": jit,
                          "t": t.tolist(), "y": y.tolist(),
                          "logml": float(-(n * mp.log(2 * mp.pi) + logdet + quad) / 2)})
    json.dump(cases, open(os.path.join(HERE, "logml_closed.json"), "w"), indent=1)


def logml_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, DotProduct, ExpSineSquared
    rng = np.random.default_rng(7)
    cases = []
    for n in (12, 60, 150):
        t = np.sort(rng.uniform(0, 1, n))
        y = np.sin(9 * t) + 0.3 * rng.standard_normal(n)
        noise, jit = 0.04, 1e-5
        # Plus(Times(Linear, Periodic), SquaredExponential)
        s0, lp, pp, ap, ls, as_ = 0.6, 0.9, 0.31, 0.8, 0.25, 0.5
        kern = DotProduct(sigma_0=s0) * (ConstantKernel(ap) * ExpSineSquared(length_scale=lp, periodicity=pp)) \
            + ConstantKernel(as_) * RBF(length_scale=ls)
        gpr = GaussianProcessRegressor(kernel=kern, alpha=noise + jit, optimizer=None).fit(t[:, None], y)
        lm = gpr.log_marginal_likelihood(gpr.kernel_.theta)
        cases.append({"name": f"lin*per+se_n{n}", "prog": [2, 5, 7, 3, 6],
                      "theta": [0.0, s0 * s0, 1.0, lp, pp, ap, ls, as_], "noise": noise, "jitter": jit,
                      "t": t.tolist(), "y": y.tolist(), "logml": float(lm)})
    json.dump(cases, open(os.path.join(HERE, "logml_sklearn.json"), "w"), indent=1)


def reference_fixtures():
    """Inputs copied from the reference's test snippets; expected values from mpmath."""
    values10 = [10.0, 15.0, 12.0, 18.0, 22.0, 25.0, 20.0, 16.0, 14.0, 11.0]   # test_nowcast_functions.jl:29-30
    scen = [[12.0, 13.0], [11.5, 12.8]]                                       # :37-41
    scen_matrix = [[12.0, 11.8], [13.0, 12.5]]                                # :285 (columns = scenarios)
    n, k, h = 10, 2, 2
    days = np.arange(n + k + h, dtype=float)
    t = (days - days[0]) / (days[n - 1] - days[0])
    y = np.array(values10)
    ya = 2.0 / (y.max() - y.min()); yb = -1.0 - ya * y.min()
    tree = ("plus", ("lin", 0.2, 0.3, 0.8), ("per", 0.9, 0.45, 0.6))
    prog, theta = flatten(tree)
    noise, jit = 0.1, 1e-5
    out = {"values10": values10, "scenarios": scen, "scenario_matrix": scen_matrix, "n": n, "k": k, "h": h,
           "t": t.tolist(), "ya": ya, "yb": yb, "prog": prog, "theta": theta, "noise": noise, "jitter": jit,
           "cases": []}
    base = gp_mp(tree, noise, jit, t[:n], ya * y + yb)
    for sc in scen + [list(c) for c in zip(*scen_matrix)]:
        yy = np.concatenate([ya * y + yb, ya * np.array(sc) + yb])
        r = gp_mp(tree, noise, jit, t[:n + k], yy, t[n + k:])
        out["cases"].append({"scenario": sc, "logml_n": base["logml"], "logml_m": r["logml"],
                             "mu": [(m - yb) / ya for m in r["mu"]],
                             "L": [[v / ya for v in row] for row in r["L"]]})
    # 30-point trend + noise series of test_model_fitting.jl:7-9 shape (values regenerated: Julia RNG not available)
    rng = np.random.default_rng(42)
    y30 = 10.0 + 0.5 * np.arange(30) + rng.standard_normal(30)
    t30 = np.arange(30) / 29.0
    a30 = 2.0 / (y30.max() - y30.min()); b30 = -1.0 - a30 * y30.min()
    tree2 = ("plus", ("lin", 0.0, 0.1, 1.2), ("ge", 0.3, 1.5, 0.2))
    p2, th2 = flatten(tree2)
    r2 = gp_mp(tree2, 0.05, jit, t30, a30 * y30 + b30, [30 / 29.0, 31 / 29.0, 32 / 29.0, 33 / 29.0])
    out["trend30"] = {"y": y30.tolist(), "t": t30.tolist(), "t_star": [30 / 29.0, 31 / 29.0, 32 / 29.0, 33 / 29.0],
                      "ya": a30, "yb": b30, "prog": p2, "theta": th2, "noise": 0.05, "jitter": jit,
                      "logml": r2["logml"], "mu": [(m - b30) / a30 for m in r2["mu"]],
                      "L": [[v / a30 for v in row] for row in r2["L"]]}
    json.dump(out, open(os.path.join(HERE, "reference_fixtures.json"), "w"), indent=1)


def fma_exact(a, b, c):
    """round-to-nearest-even of a*b + c computed exactly."""
    r = Fraction(a) * Fraction(b) + Fraction(c)
    return float(r)   # Fraction -> float is correctly rounded


def draws():
    rng = np.random.default_rng(11)
    K, P, h, D = 3, 4, 5, 6
    logw = rng.standard_normal((K, P))
    mu = rng.standard_normal((K, P, h)) * 10
    L = np.tril(rng.standard_normal((K, P, h, h)))
    zeta = rng.standard_normal((K, D, h))
    comp = rng.integers(0, P, (K, D))
    x = np.empty((h, K * D))
    for s in range(K):
        for d in range(D):
            c = comp[s, d]
            for i in range(h):
                acc = float(mu[s, c, i])
                for j in range(i + 1):
                    acc = fma_exact(float(L[s, c, i, j]), float(zeta[s, d, j]), acc)
                x[i, s * D + d] = acc
    # weights / ESS in mpmath
    ess, w = [], []
    for s in range(K):
        mx = max(logw[s])
        e = [mp.e ** (mp.mpf(float(v)) - mp.mpf(float(mx))) for v in logw[s]]
        tot = sum(e)
        ws = [v / tot for v in e]
        w.append([float(v) for v in ws]); ess.append(float(1 / sum(v * v for v in ws)))
    json.dump({"logw": logw.tolist(), "mu": mu.tolist(), "L": L.tolist(), "zeta": zeta.tolist(),
               "comp": comp.tolist(), "x": x.tolist(), "ess": ess, "w": w},
              open(os.path.join(HERE, "draws.json"), "w"), indent=1)


if __name__ == "__main__":
    kernel_values(); logml_closed(); logml_sklearn(); reference_fixtures(); draws()
    print("golden fixtures written to", HERE)
