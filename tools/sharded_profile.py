"""cProfile of `forecast_with_nowcasts_sharded` (default schedule) on seven C4-shaped series — one rank's share at 8 GPUs:
`python tools/sharded_profile.py`."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nowcastautogp_b200 as nag
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
from nowcastautogp_b200.gpmodel import GPModel
S, n4, k4, h4, P4, K4, D4 = 7, 150, 1, 4, 64, 1000, 20
eng = Engine(0)
dates = np.datetime64("2022-10-01") + 7 * np.arange(n4 + k4 + h4)
models, nowcasts = [], []
for s_ in range(S):
    _, raw = syn.weekly_series(n4, 1001 + s_)
    rg = np.random.default_rng([2026, s_])
    m_ = GPModel(dates[:n4], np.log(raw), n_particles=P4, rng=rg, engine=eng)
    m_.fit_smc(schedule=[n4], n_mcmc=0, n_hmc=0, shuffle=False)
    models.append(m_)
    scen = raw[-1] * np.exp(0.1 + 0.027 * rg.standard_normal((k4, K4)))
    nowcasts.append(nag.create_nowcast_data(scen, dates[n4:n4 + k4], transformation=np.log))
fdates = dates[n4 + k4:]
import torch
for _ in range(2):
    nag.forecast_with_nowcasts_sharded(models, nowcasts, fdates, D4, device="cuda:0")
t0 = time.perf_counter()
for _ in range(5):
    nag.forecast_with_nowcasts_sharded(models, nowcasts, fdates, D4, device="cuda:0")
print(f"sharded, 7 series: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call")
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    nag.forecast_with_nowcasts_sharded(models, nowcasts, fdates, D4, device="cuda:0")
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:4500])
