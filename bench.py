#!/usr/bin/env python
"""bench.py — forecast draws/sec (and batched logML evals/sec) of the GP hot path on B200.

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A step = one pass of the hot path over one batch: BASELINE.json configs[1] (vignette shape: weekly
series n=150, 32 particles x 1000 nowcast scenarios, 1 nowcast point, 9 forecast dates, 20 draws
per scenario) in the per-(scenario, particle) regime — every one of the K*P = 32000 instances has
its own hyperparameters (what `n_hmc > 0` rejuvenation produces, /root/reference/src/forecasting.jl:
145-149), so nothing is shared between scenarios: one fused Gram -> Cholesky -> solve per instance,
then weights/ESS and the mixture draws. The default-API fast path (`n_hmc == 0`, one factorisation
per particle) is timed as well and reported under "fast_path".

`value`  : device-resident inputs, CUDA-event timed, max over ranks (weak scaling: one series/rank).
`e2e`    : same step through the C ABI with pinned HOST buffers, H2D/D2H inside the timed region.
`--impl reference`: the CPU restatement of the reference schedule (oracle/), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n=150, k=1, h=9, P=32, K=1000, D=20, max_depth=4)
METRIC = "forecast draws/sec"
UNIT = "draws/s"
WORKLOAD = ("BASELINE configs[1] vignette shape: weekly series n=150, P=32 particles x K=1000 nowcast "
            "scenarios, k=1, h=9, D=20 draws/scenario; per-(scenario,particle) hyperparameters "
            "(n_hmc>0 regime, K*P=32000 fused Gram+Cholesky instances per step)")


def make_inputs(rank: int):
    from nowcastautogp_b200 import synthetic as syn
    c = CFG
    w = syn.make_workload(c["n"], c["k"], c["h"], c["K"], c["P"], seed=20261018 + 2 + 1000 * rank,
                          max_depth=c["max_depth"])
    theta_k, noise_k = syn.perturbed_theta(w.ens, c["K"], seed=77 + rank)
    rng = np.random.default_rng(4242 + rank)
    zeta = rng.standard_normal((c["K"], c["D"], c["h"]))
    u = rng.uniform(size=(c["K"], c["D"]))
    return w, theta_k, noise_k, zeta, u


def flops_per_instance(n, k, h):
    """factor + forward solve + reductions of the joint q x q problem."""
    m, q = n + k, n + k + h
    return q ** 3 / 3.0 + 2.0 * m * m + 2.0 * h * m


# ------------------------------------------------------------------------------------------------
def clocks_sampler_start():
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=f, stderr=subprocess.DEVNULL)
    except OSError:
        return None, f.name
    return p, f.name


def _sample_lines(path):
    try:
        with open(path) as fh:
            return sum(1 for _ in fh)
    except OSError:
        return 0


def clocks_sampler_stop(p, path, dev_index, skip_lines=0):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    if p is None:
        return out
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()
    sm, reasons, mx = [], set(), None
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
        for ln, line in enumerate(open(path)):
            if ln < skip_lines:
                continue
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9 or parts[0] != str(dev_index):
                continue
            try:
                sm.append(float(parts[1])); mx = float(parts[2])
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(path)
    except OSError:
        pass
    if sm:
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = mx
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
    return out


# ------------------------------------------------------------------------------------------------
def cpu_reference(w, theta_k, noise_k, zeta, u, steps, warmup, target_s=12.0):
    """Reference schedule on the host cores (oracle port), bounded sample of the same workload."""
    from oracle.oracle import Oracle
    o = Oracle()
    try:     # all host cores this process may use, whatever OMP_NUM_THREADS the launcher exported (torchrun sets 1)
        o.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        o.set_num_threads(os.cpu_count() or 1)
    c = CFG

    def run(Ks):
        t0 = time.perf_counter()
        r = o.forecast_instances(w.ens, c["n"], c["k"], c["h"], w.t, w.y1, w.y2[:Ks], w.logw0, w.ya, w.yb,
                                 g=w.g, step=w.step, use_joint=False,
                                 theta_per_scenario=theta_k[:Ks], noise_per_scenario=noise_k[:Ks])
        o.draws(r["logw"], r["mu"], r["L"], zeta[:Ks], u=u[:Ks])
        return time.perf_counter() - t0

    probe_K = 4
    t_probe = run(probe_K)
    Ks = int(max(probe_K, min(c["K"], probe_K * target_s / max(t_probe, 1e-6) / max(steps + warmup, 1))))
    for _ in range(warmup):
        run(Ks)
    times = [run(Ks) for _ in range(steps)]
    t = float(np.mean(times))
    return dict(value=Ks * c["D"] / t, unit=UNIT, cores=o.num_threads(), kind="port",
                sample=f"{Ks} of {c['K']} scenarios x {c['P']} particles per step, reference schedule "
                       f"(rebuild+add_data!+predict_mvn LU+chol+draws), OpenMP over instances"), t * 1e3, Ks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-sanity", action="store_true", help="timing experiments with deliberately wrong kernels")
    ap.add_argument("--only-value", action="store_true", help="time only the device-resident step")
    ap.add_argument("--no-micro", action="store_true", help="skip the configs[2]/configs[4] micro-benchmarks")
    ap.add_argument("--variant", type=int, default=0, help="factorisation kernel: 0 auto, 2 tile kernel, 3 slot kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = CFG

    if args.impl == "reference":
        if rank != 0:
            return 0
        w, theta_k, noise_k, zeta, u = make_inputs(0)
        cb, ms, Ks = cpu_reference(w, theta_k, noise_k, zeta, u, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "sample_scenarios_per_step": Ks},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created: keep stdout to the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    from nowcastautogp_b200.engine import Engine
    from nowcastautogp_b200.kernels import FlatEnsemble

    dev = torch.device("cuda", local_rank)
    eng = Engine(local_rank)
    # a dedicated stream: kernels, copies, events and the NCCL gather are all ordered on it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.set_variant(args.variant)

    w, theta_k, noise_k, zeta, u = make_inputs(rank)
    n, k, h, P, K, D = c["n"], c["k"], c["h"], c["P"], c["K"], c["D"]

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def pinned(a):
        t_ = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t_

    # ---- device-resident buffers (bulk data; the <2 KB of programs/offsets/time grid stay host) ----
    d_theta, d_noise = to_dev(theta_k), to_dev(noise_k)
    d_y1, d_y2, d_logw0 = to_dev(w.y1), to_dev(w.y2), to_dev(w.logw0)
    d_zeta, d_u = to_dev(zeta), to_dev(u)
    d_logw = torch.empty((K, P), dtype=torch.float64, device=dev)
    d_mu = torch.empty((K, P, h), dtype=torch.float64, device=dev)
    d_L = torch.empty((K, P, h, h), dtype=torch.float64, device=dev)
    d_info = torch.zeros((K, P), dtype=torch.int32, device=dev)
    d_x = torch.empty((K * D, h), dtype=torch.float64, device=dev)
    d_ens_theta, d_ens_noise = to_dev(w.ens.theta), to_dev(w.ens.noise)
    ens_dev = FlatEnsemble(w.ens.prog, w.ens.prog_off, d_ens_theta, w.ens.theta_off, d_ens_noise)
    gather_x = gather_lw = None
    if world > 1:
        gather_x = torch.empty((world,) + tuple(d_x.shape), dtype=torch.float64, device=dev)
        gather_lw = torch.empty((world, K, P), dtype=torch.float64, device=dev)

    def step_device():
        eng.forecast_instances(ens_dev, n, k, h, w.t, d_y1, d_y2, d_logw0, w.ya, w.yb, g=w.g, step=w.step,
                               theta=d_theta, noise=d_noise, K=K, logw=d_logw, mu=d_mu, L=d_L, info=d_info)
        eng.draw(d_logw, d_mu, d_L, d_zeta, u=d_u, x=d_x, want_aux=False)
        if world > 1:   # the only collective on the path: gather log-weights and draws (NCCL/NVLink)
            dist.all_gather_into_tensor(gather_x, d_x)
            dist.all_gather_into_tensor(gather_lw, d_logw)

    def kernel_only():
        eng.forecast_instances(ens_dev, n, k, h, w.t, d_y1, d_y2, d_logw0, w.ya, w.yb, g=w.g, step=w.step,
                               theta=d_theta, noise=d_noise, K=K, logw=d_logw, mu=d_mu, L=d_L, info=d_info)

    def step_fast_device():
        eng.forecast_with_nowcasts(ens_dev, n, k, h, w.t, d_y1, d_y2, d_logw0, d_zeta, w.ya, w.yb, g=w.g,
                                   step=w.step, u=d_u, x=d_x, logw=d_logw, info=d_info[0], K=K, D=D)

    # ---- host (pinned) buffers for the end-to-end path ------------------------------------------------
    h_theta, h_noise = pinned(theta_k), pinned(noise_k)
    h_y1, h_y2, h_logw0 = pinned(w.y1), pinned(w.y2), pinned(w.logw0)
    h_zeta, h_u = pinned(zeta), pinned(u)
    h_x = torch.empty((K * D, h), dtype=torch.float64).pin_memory()
    h_logw = torch.empty((K, P), dtype=torch.float64).pin_memory()
    h_info = torch.zeros((K, P), dtype=torch.int32).pin_memory()
    h_ens = FlatEnsemble(w.ens.prog, w.ens.prog_off, w.ens.theta, w.ens.theta_off, w.ens.noise)
    h2d = sum(t_.numel() * t_.element_size() for t_ in (h_theta, h_noise, h_y1, h_y2, h_logw0, h_zeta, h_u))
    h2d += w.ens.prog.nbytes + w.ens.prog_off.nbytes + w.ens.theta_off.nbytes + w.t.nbytes + w.g.nbytes
    h2d += h_logw.numel() * 8    # log-weights go back in for the draw call
    d2h = sum(t_.numel() * t_.element_size() for t_ in (h_x, h_logw, h_info))

    def step_e2e():
        # C ABI with host buffers: H2D of every input, D2H of draws, log-weights and info
        eng.forecast_instances(h_ens, n, k, h, w.t, h_y1, h_y2, h_logw0, w.ya, w.yb, g=w.g, step=w.step,
                               theta=h_theta, noise=h_noise, K=K, logw=h_logw, mu=d_mu, L=d_L, info=h_info)
        eng.draw(h_logw, d_mu, d_L, h_zeta, u=h_u, x=h_x, want_aux=False)

    h2d_fast = sum(t_.numel() * t_.element_size() for t_ in (h_y1, h_y2, h_logw0, h_zeta, h_u)) + \
        w.ens.prog.nbytes + w.ens.theta.nbytes + w.ens.noise.nbytes + w.t.nbytes + w.g.nbytes

    def step_fast_e2e():
        eng.forecast_with_nowcasts(h_ens, n, k, h, w.t, h_y1, h_y2, h_logw0, h_zeta, w.ya, w.yb, g=w.g,
                                   step=w.step, u=h_u, x=h_x, logw=h_logw, info=h_info[0], K=K, D=D)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed(fn, steps, warmup, do_flush=True):
        """CUDA-event time per step on the launching stream; L2 flushed between steps (untimed)."""
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total = 0.0
        for _ in range(steps):
            if do_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = total / steps
        if world > 1:
            t_ = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms = float(t_.item())
        return ms

    # clocks: nvidia-smi samples every 100 ms and needs a moment to start, the timed region is tens of ms. Wait for
    # its first line (idle), time the steps, then keep rank 0's GPU under the same kernel until a few samples have
    # been taken under load; only samples from the start of the timed region on are used.
    sampler, spath = clocks_sampler_start() if rank == 0 else (None, None)
    skip = 0
    if sampler is not None:
        t_wait = time.time() + 3.0
        while _sample_lines(spath) == 0 and time.time() < t_wait:
            time.sleep(0.02)
        skip = _sample_lines(spath)
    l0 = eng.launch_count
    ms_step = timed(step_device, args.steps, args.warmup)
    launches = (eng.launch_count - l0) // (args.steps + args.warmup) * args.steps
    if sampler is not None:
        n_dev = max(1, torch.cuda.device_count())
        t_wait = time.time() + 1.5
        while _sample_lines(spath) < skip + 3 * n_dev and time.time() < t_wait:
            kernel_only()
            torch.cuda.synchronize()
    clocks = clocks_sampler_stop(sampler, spath, local_rank, skip) if rank == 0 else None
    ms_kernel = timed(kernel_only, args.steps, 1)
    if args.only_value:
        if rank == 0:
            print(json.dumps({"ms_per_step": ms_step, "kernel_ms": ms_kernel, "kernel": eng.last_kernel}))
        eng.close()
        return 0
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    ms_fast = timed(step_fast_device, args.steps, args.warmup)
    ms_fast_e2e = timed(step_fast_e2e, args.steps, args.warmup)

    # ---- secondary metric (BASELINE configs[2]): batched logML evals/sec, 1024 particles x n=512 ----------
    def micro_logml():
        from nowcastautogp_b200 import synthetic as syn
        Bm, nm = 1024, 512
        wm = syn.make_workload(nm, 0, 0, 1, Bm, seed=20261018 + 3 + 1000 * rank, max_depth=4, period=365.0)
        ens_m = FlatEnsemble(wm.ens.prog, wm.ens.prog_off, to_dev(wm.ens.theta), wm.ens.theta_off, to_dev(wm.ens.noise))
        d_y, d_lm = to_dev(wm.y1), torch.empty(Bm, dtype=torch.float64, device=dev)
        d_inf = torch.zeros(Bm, dtype=torch.int32, device=dev)
        tm, gm = wm.t[:nm], wm.g[:nm]

        def run():
            eng.logml_batch(ens_m, tm, d_y, g=gm, step=wm.step, logml=d_lm, info=d_inf)
        ms = timed(run, max(3, args.steps // 2), 3)
        ok_ = int(d_inf.abs().max().item()) == 0 and bool(torch.isfinite(d_lm).all().item())
        fl_ = Bm * (nm ** 3 / 3.0 + 2.0 * nm * nm)
        return {"what": "BASELINE configs[2]: 1024 particles, compositional kernels, n=512, Gram+Cholesky+logdet "
                        "(chol_large_kernel, factor in HBM/L2, Gram never stored)",
                "value": Bm * world / (ms * 1e-3), "unit": "logML evals/s", "ms_per_step": ms, "ok": ok_,
                "roofline": {"bound": "tensor", "achieved": fl_ / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                             "flops_per_launch": fl_}}

    # ---- BASELINE configs[4]: daily series n=2048, 256 particles, incremental add_data! (rank-append) -------
    def micro_append():
        from nowcastautogp_b200 import synthetic as syn
        Pa, na, ka = 256, 2048, 1
        wa = syn.make_workload(na + 8 * ka, 0, 0, 1, Pa, seed=20261018 + 5 + 1000 * rank, max_depth=4, period=365.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the factorisation itself at this size, steady state (workspace from the context's arena)
        ens_a = FlatEnsemble(wa.ens.prog, wa.ens.prog_off, to_dev(wa.ens.theta), wa.ens.theta_off, to_dev(wa.ens.noise))
        d_ya, d_lma = to_dev(wa.y1[:na]), torch.empty(Pa, dtype=torch.float64, device=dev)
        d_infa = torch.zeros(Pa, dtype=torch.int32, device=dev)
        ta, ga = wa.t[:na], wa.g[:na]
        ms_store = timed(lambda: eng.logml_batch(ens_a, ta, d_ya, g=ga, step=wa.step, logml=d_lma, info=d_infa), 2, 3)
        # the appendable store (allocates 4.3 GB of factor storage: not timed), then k = 1 appends
        f = eng.factor_store_large(wa.ens, wa.t[:na], wa.y1[:na], capacity=na + 8 * ka, g=wa.g[:na], step=wa.step,
                                   check=False)
        ok_ = bool((f.info == 0).all())
        tms = []
        cur = na
        for i in range(8):
            flush.fill_(1)
            e0.record(stream)
            eng.factor_append(f, wa.t[cur:cur + ka], wa.y1[cur:cur + ka], g_new=wa.g[cur:cur + ka], check=False)
            e1.record(stream); e1.synchronize()
            tms.append(e0.elapsed_time(e1)); cur += ka
        f.free()
        ms_app = float(np.median(tms[1:]))
        bytes_ = Pa * 4.0 * na * (na + 1)          # the stored factor read once
        fl_ = Pa * (na ** 3 / 3.0 + 2.0 * na * na)
        return {"what": "BASELINE configs[4]: daily series n=2048, 256 particles; from-scratch factorisation "
                        "(nagp_logml_batch, device-resident) and in-place rank-append of k=1 point on the stored factor "
                        "(append time includes the blocking C-ABI call, host->device of the new point and "
                        "device->host of dlogml/logml/info)",
                "ok": ok_, "factor_ms": ms_store,
                "factor_roofline": {"bound": "tensor", "achieved": fl_ / (ms_store * 1e-3) / 1e12, "unit": "TFLOP/s"},
                "append_ms": ms_app, "appends_per_s": Pa * world / (ms_app * 1e-3),
                "append_roofline": {"bound": "hbm", "achieved": bytes_ / (ms_app * 1e-3) / 1e9, "unit": "GB/s",
                                    "bytes_per_launch": bytes_}}

    # ---- the HMC primitive — logML + gradient for every (scenario, particle) chain of the workload ----
    def micro_grad():
        m = n + k
        total = int(w.ens.theta_off[-1])
        outs = (torch.empty((K, P), dtype=torch.float64, device=dev), torch.empty((K, total), dtype=torch.float64, device=dev),
                torch.empty((K, P), dtype=torch.float64, device=dev), torch.zeros((K, P), dtype=torch.int32, device=dev))

        def run():
            eng.logml_grad(ens_dev, w.t[:m], d_y1, y2=d_y2, g=w.g[:m], step=w.step, theta=d_theta, noise=d_noise, out=outs)
        ms = timed(run, max(3, args.steps // 2), 3)
        ok_ = int(outs[3].abs().max().item()) == 0 and bool(torch.isfinite(outs[1]).all().item())
        fl_ = K * P * (m ** 3 / 3.0 + 2.0 * m ** 3 / 3.0)          # factorisation + inverse from the factor
        return {"what": "logML + d logML/d(theta, noise) for K*P = 32000 per-scenario HMC chains at n+k=151 "
                        "(one leapfrog stage of mcmc_parameters! on every chain): tile kernel keeps L, gradient kernel "
                        "forms K^-1 in place on the FP64 tensor pipe and differentiates the tree per lag",
                "value": K * P * world / (ms * 1e-3), "unit": "gradient evals/s", "ms_per_step": ms, "ok": ok_,
                "roofline": {"bound": "tensor", "achieved": fl_ / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                             "flops_per_launch": fl_, "note": "m^3/3 (Cholesky) + 2m^3/3 (inverse) per instance; "
                             "the per-entry reverse-mode work is not counted"}}

    # ---- device-resident HMC — one iteration (10 leapfrog stages) of all K*P chains, one C-ABI call -------
    def micro_hmc():
        from nowcastautogp_b200 import kernels as kn_
        m = n + k
        names = kn_.theta_slot_names(w.ens.prog.tobytes())
        base = w.ens.theta
        ident = np.array([nm in ("intercept", "location") for nm in names]) | (base <= 0)
        fixed = np.array([nm == "scale" for nm in names])
        kind = np.where(fixed, 5, np.where(ident, 3, 0)).astype(np.int32)
        sa = np.where(fixed, base, np.where(ident, 0.0, np.log(np.where(base > 0, base, 1.0))))
        sb = np.ones(len(base))
        with np.errstate(divide="ignore", invalid="ignore"):
            z0 = np.where(kind[None, :] == 0, np.log(theta_k / np.where(base > 0, base, 1.0)[None, :]), theta_k)
        z0 = np.ascontiguousarray(np.where(np.isfinite(z0), z0, 0.0))
        nz0 = np.ascontiguousarray(np.log(noise_k / w.ens.noise[None, :]))
        rng_ = np.random.default_rng(99 + rank)
        Lf, n_it = 10, 1
        mom = rng_.standard_normal((n_it, K, len(base)))
        mnz = rng_.standard_normal((n_it, K, P))
        logu = np.log(rng_.uniform(size=(n_it, K, P)))
        res = {}

        def run():
            res["out"] = eng.hmc(w.ens.prog, w.ens.prog_off, w.ens.theta_off, kind, sa, sb, (0, 0.0, 1.0), z0.copy(),
                                 nz0.copy() + np.log(w.ens.noise)[None, :], w.t[:m], w.y1, y2=w.y2, g=w.g[:m], step=w.step,
                                 n_leapfrog=Lf, eps=0.002, momenta=mom, noise_momenta=mnz, log_u=logu)
        ms = timed(run, 2, 1)
        _, _, lm_, nacc_, info_ = res["out"]
        evals = K * P * (n_it * Lf + 1)
        return {"what": "nagp_hmc: 1 HMC iteration (10 leapfrog stages + the initial evaluation = 11 logML+gradient "
                        "evaluations) of all K*P = 32000 per-scenario chains in ONE C-ABI call with host buffers; integrator, "
                        "z->theta maps and accept/reject on the device",
                "value": evals * world / (ms * 1e-3), "unit": "gradient evals/s", "ms_per_step": ms,
                "accept_rate": float(nacc_.mean()) / n_it, "ok": bool(np.isfinite(lm_[info_ == 0]).all())}

    # ---- inverse transformation + 25/50/75 % bands of the (h, K*D) draw matrix ---------------------
    def micro_summary():
        d_q = torch.empty((h, 3), dtype=torch.float64, device=dev)
        d_p = to_dev(np.array([0.25, 0.5, 0.75]))
        lib, ctx = eng._lib, eng._ctx

        def run():
            eng._check(lib.nagp_forecast_summary(ctx, 1, 0.0, 0.0, 0.0, h, K * D, d_x.data_ptr(), d_x.data_ptr(), 3,
                                                 d_p.data_ptr(), d_q.data_ptr()))
        step_device()
        ms = timed(run, max(3, args.steps // 2), 3)
        return {"what": "inverse 'positive' transformation in place + 25/50/75 % row quantiles of the (9, 20000) draws",
                "ms_per_step": ms, "bytes_per_launch": int(h * K * D * (8 + 16 + 2 * 3 * 8 * 8))}

    micro = None if args.no_micro else {"logml": micro_logml(), "append": micro_append(), "grad": micro_grad(),
                                        "hmc": micro_hmc(), "summary": micro_summary()}

    # sanity: the step produced finite draws and no factorisation failed
    step_device()
    torch.cuda.synchronize()
    ok = bool(torch.isfinite(d_x).all().item()) and int(d_info.abs().max().item()) == 0
    if not ok and not args.skip_sanity:
        raise SystemExit("bench step produced non-finite draws or a failed factorisation")

    if rank == 0:
        draws = K * D * world
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))
        except OSError:
            pass
        peak_tf = float(peaks.get("dmma_m8n8k4_tflops", 37.1))
        fl = K * P * flops_per_instance(n, k, h)
        achieved = fl / (ms_kernel * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": draws / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "series_per_gpu": 1, "instances_per_step_per_gpu": K * P,
                       "l2": "flushed between timed steps (256 MB write)", "parallelism": f"series-sharded x{world}"},
            "clocks": clocks,
            "e2e": {"value": draws / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA; tcgen05 has no f64 kind)",
                         "kernel": "fused Gram+Cholesky+solve (nagp_fused)", "achieved": achieved,
                         "peak": peak_tf, "peak_source": "profiles/r01_fp64_peak.json (DMMA m8n8k4 measured on this "
                         "pool's B200; MEASURED_PEAKS.json has no FP64 entry)", "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         "traffic": 2722304, "traffic_source": "profiles/r01_v2i_final_summary.csv: dram__bytes_read.sum + "
                         "dram__bytes_write.sum of one launch (bytes; the Gram and the factor never leave shared memory, and the 23 MB of "
                         "outputs were still in L2 when the counter was read)",
                         "kernel_ms": ms_kernel,
                         "flops_per_launch": fl},
            "fast_path": {"what": "default API path n_hmc==0: one factorisation per particle, O(k^2+hk) per scenario",
                          "value": draws / (ms_fast * 1e-3), "unit": UNIT, "ms_per_step": ms_fast,
                          "e2e": {"value": draws / (ms_fast_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_fast_e2e,
                                  "h2d_bytes_per_step": int(h2d_fast), "d2h_bytes_per_step": int(d2h)}},
        }
        if micro is not None:
            hbm_peak = 6541.8
            try:
                hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except (OSError, KeyError, ValueError):
                pass
            micro["logml"]["roofline"].update(peak=peak_tf, frac=micro["logml"]["roofline"]["achieved"] / peak_tf)
            micro["append"]["factor_roofline"].update(peak=peak_tf, frac=micro["append"]["factor_roofline"]["achieved"] / peak_tf)
            micro["append"]["append_roofline"].update(peak=hbm_peak, peak_source="MEASURED_PEAKS.json hbm_gbs",
                                                      frac=micro["append"]["append_roofline"]["achieved"] / hbm_peak)
            micro["grad"]["roofline"].update(peak=peak_tf, frac=micro["grad"]["roofline"]["achieved"] / peak_tf)
            line["hmc_gradient_microbench"] = micro["grad"]
            line["hmc_microbench"] = micro["hmc"]
            line["summary_microbench"] = micro["summary"]
            line["logml_microbench"] = micro["logml"]
            line["append_microbench"] = micro["append"]
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_reference(w, theta_k, noise_k, zeta, u, steps=1, warmup=0)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
