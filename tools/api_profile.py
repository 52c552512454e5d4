"""cProfile of the public `forecast_with_nowcasts` (default schedule, n_hmc = 0) on one C4-shaped series:
where the host time goes once the device work is 0.1 ms. `python tools/api_profile.py [n_hmc]`."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nowcastautogp_b200 as nag
from nowcastautogp_b200 import synthetic as syn
from nowcastautogp_b200.engine import Engine
from nowcastautogp_b200.gpmodel import GPModel
n_hmc = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n4, k4, h4, P4, K4, D4 = 150, 1, 4, 64, 1000, 20
eng = Engine(0)
d0 = np.datetime64("2022-10-01")
dates = d0 + 7 * np.arange(n4 + k4 + h4)
_, raw = syn.weekly_series(n4, 1001)
rg = np.random.default_rng([2026, 0])
m_ = GPModel(dates[:n4], np.log(raw), n_particles=P4, rng=rg, engine=eng)
m_.fit_smc(schedule=[n4], n_mcmc=0, n_hmc=0, shuffle=False)
scen = raw[-1] * np.exp(0.1 + 0.027 * rg.standard_normal((k4, K4)))
nowcasts = nag.create_nowcast_data(scen, dates[n4:n4 + k4], transformation=np.log)
fdates = dates[n4 + k4:]
for _ in range(2):
    nag.forecast_with_nowcasts(m_, nowcasts, fdates, D4, n_hmc=n_hmc)
t0 = time.perf_counter()
for _ in range(5):
    out = nag.forecast_with_nowcasts(m_, nowcasts, fdates, D4, n_hmc=n_hmc)
print(f"forecast_with_nowcasts(n_hmc={n_hmc}): {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call, out {np.shape(out)}")
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    nag.forecast_with_nowcasts(m_, nowcasts, fdates, D4, n_hmc=n_hmc)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
