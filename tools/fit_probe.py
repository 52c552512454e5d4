"""Times make_and_fit_models (lockstep, coalesced) against a loop of make_and_fit_model on S weekly series."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nowcastautogp_b200 as ng
from nowcastautogp_b200.api import make_and_fit_models
from nowcastautogp_b200.engine import Engine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
P = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = 150
eng = Engine(0)
dates = np.datetime64("2022-01-01") + 7 * np.arange(n)
datas = []
for s in range(S):
    rng = np.random.default_rng(1000 + s)
    tt = np.arange(n)
    y = np.exp(np.log(50) + np.sin(2 * np.pi * tt / 52) + 0.02 * tt + 0.15 * rng.standard_normal(n))
    datas.append(ng.TData(dates, y, transformation=np.log))
kw = dict(n_particles=P, smc_data_proportion=0.2, n_mcmc=4, n_hmc=2)
ng.make_and_fit_model(datas[0], rng=np.random.default_rng(0), engine=eng, **kw)      # warm-up
l0 = eng.launch_count; t0 = time.perf_counter()
ms = make_and_fit_models(datas, rng=np.random.default_rng(1), engine=eng, **kw)
t1 = time.perf_counter(); l1 = eng.launch_count
print(f"lockstep : S={S} P={P}: {t1 - t0:.2f} s, {l1 - l0} launches, {make_and_fit_models.last_stats}")
Sseq = min(S, 4)
t0 = time.perf_counter(); l0 = eng.launch_count
for s in range(Sseq):
    ng.make_and_fit_model(datas[s], rng=np.random.default_rng(s), engine=eng, **kw)
t1 = time.perf_counter(); l1 = eng.launch_count
print(f"one by one: {Sseq} series: {(t1 - t0) / Sseq:.2f} s/series -> {S} series = {(t1 - t0) / Sseq * S:.2f} s, {(l1 - l0) // Sseq} launches/series")
