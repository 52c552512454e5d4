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

`value`  : device-resident inputs, CUDA-event timed, max over ranks (weak scaling: one series/rank), through the
           product multi-GPU function `sharding.sharded_forecast` at every N (one packed all-gather of draws and
           log-weights over NCCL when N > 1).
`e2e`    : same step through the C ABI with pinned HOST buffers, H2D/D2H inside the timed region.
`parity_max_rel`: after the timed region 64 random instances of the timed batch are recomputed by the CPU oracle
           (checker only) and compared with what the device wrote; the run fails above 1e-9.
`c4`     : (N = 8, or `--c4`) BASELINE configs[3]: 53 series x 64 particles x 1000 nowcasts through
           `sharding.partition` / `sharded_forecast` (device-resident) and through the public
           `forecast_with_nowcasts_sharded` API, both regimes.
`--impl reference`: the reference schedule on the host cores (oracle/nagp_cpu_blocked.c: blocked AVX2 Cholesky/LU,
           -O3, checked against the oracle), all host threads.
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
def host_threads():
    try:     # all host cores this process may use, whatever OMP_NUM_THREADS the launcher exported (torchrun sets 1)
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def cpu_reference(w, theta_k, noise_k, zeta, u, steps, warmup, target_s=12.0, cfg=None):
    """Reference schedule on the host cores, bounded sample of the same workload. The timed code is
    oracle/nagp_cpu_blocked.c (blocked, AVX2/FMA Cholesky and LU, -O3; tests/test_cpu_baseline.py checks it against the
    scalar oracle); the draws come from the oracle's bit-exact draw routine."""
    from oracle.oracle import BlockedCpu, Oracle
    o, f = Oracle(), BlockedCpu()
    nth = host_threads()
    o.set_num_threads(nth)
    f.set_num_threads(nth)
    c = CFG if cfg is None else cfg

    def run(Ks):
        t0 = time.perf_counter()
        r = f.forecast_instances(w.ens, c["n"], c["k"], c["h"], w.t, w.y1, w.y2[:Ks], w.logw0, w.ya, w.yb,
                                 g=w.g, step=w.step, theta_per_scenario=None if theta_k is None else theta_k[:Ks],
                                 noise_per_scenario=None if noise_k is None else noise_k[:Ks])
        o.draws(r["logw"], r["mu"], r["L"], zeta[:Ks], u=u[:Ks])
        return time.perf_counter() - t0

    probe_K = 4
    t_probe = run(probe_K)
    Ks = int(max(probe_K, min(c["K"], probe_K * target_s / max(t_probe, 1e-6) / max(steps + warmup, 1))))
    for _ in range(warmup):
        run(Ks)
    times = [run(Ks) for _ in range(steps)]
    t = float(np.mean(times))
    n, m, h = c["n"], c["n"] + c["k"], c["h"]
    fl = Ks * c["P"] * (n ** 3 / 3 + m ** 3 / 3 + 2 * m ** 3 / 3 + 2 * h * m * m + 2 * n * n + 2 * m * m)
    pc, pl = f.factor_rates(m, 30)
    return dict(value=Ks * c["D"] / t, unit=UNIT, cores=f.num_threads(), kind="port",
                sample=f"{Ks} of {c['K']} scenarios x {c['P']} particles per step, reference schedule "
                       f"(rebuild+add_data!+predict_mvn LU+chol+draws), OpenMP over instances; blocked AVX2 build: "
                       f"{fl / t / 1e9 / f.num_threads():.2f} GFLOP/s/core over the whole schedule (per-entry kernel "
                       f"evaluation included, as in the reference), factorisations alone {pc:.1f} (Cholesky) / {pl:.1f} (LU) "
                       f"GFLOP/s single-thread at order {m}"), t * 1e3, Ks


def read_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed ncu summary
    named in profiles/current.json (None when absent)."""
    try:
        cur = json.load(open(os.path.join(ROOT, "profiles", "current.json")))["fused_kernel"]
        rd = wr = None
        import csv
        for row in csv.reader(open(os.path.join(ROOT, "profiles", cur["summary"]))):
            if len(row) < 4:
                continue
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row[2])
            if mult is None:
                continue
            if row[1] == "dram__bytes_read.sum":
                rd = float(row[3]) * mult
            elif row[1] == "dram__bytes_write.sum":
                wr = float(row[3]) * mult
        if rd is None or wr is None:
            return None, None
        return int(rd + wr), f"profiles/{cur['summary']} ({cur.get('note', 'ncu --set full, one launch')})"
    except (OSError, KeyError, ValueError):
        return None, None


def cpu_logml_baseline(rank=0):
    """BASELINE configs[2] on the host cores: a bounded sample of the 1024 x (n = 512) batched logML (blocked AVX2
    Cholesky, per-entry kernel evaluation), all host threads."""
    from nowcastautogp_b200 import synthetic as syn
    from nowcastautogp_b200 import kernels as kn_
    from oracle.oracle import BlockedCpu
    f = BlockedCpu()
    nth = host_threads()
    f.set_num_threads(nth)
    nm = 512
    Bs = max(2 * nth, 16)
    wm = syn.make_workload(nm, 0, 0, 1, 1024, seed=20261018 + 3 + 1000 * rank, max_depth=4, period=365.0)
    ens = kn_.pack_ensemble(wm.trees[:Bs], np.asarray(wm.noise)[:Bs])
    f.logml_batch(ens, wm.t[:nm], wm.y1, g=wm.g[:nm], step=wm.step)
    t0 = time.perf_counter()
    _, info = f.logml_batch(ens, wm.t[:nm], wm.y1, g=wm.g[:nm], step=wm.step)
    dt = time.perf_counter() - t0
    return {"value": Bs / dt, "unit": "logML evals/s", "cores": f.num_threads(), "kind": "port",
            "sample": f"first {Bs} of the 1024 instances (n = 512), Gram + blocked Cholesky + solve + logdet, OpenMP over instances",
            "ok": bool((info == 0).all())}


def cpu_append_baseline(rank=0):
    """BASELINE configs[4] on the host cores: factor a sample of the 256 particles at n = 2048 (blocked AVX2 Cholesky) and
    append k = 1 point to every stored factor (forward substitution of the new row: reads the factor once)."""
    from nowcastautogp_b200 import synthetic as syn
    from nowcastautogp_b200 import kernels as kn_
    from oracle.oracle import BlockedCpu
    f = BlockedCpu()
    nth = host_threads()
    f.set_num_threads(nth)
    na, ka = 2048, 1
    Bs = int(min(max(nth, 4), 32))                      # 34 MB of factor per instance
    wa = syn.make_workload(na + 8 * ka, 0, 0, 1, 256, seed=20261018 + 5 + 1000 * rank, max_depth=4, period=365.0)
    ens = kn_.pack_ensemble(wa.trees[:Bs], np.asarray(wa.noise)[:Bs])
    t0 = time.perf_counter()
    st = f.factor_store(ens, wa.t[:na], wa.y1[:na], n_cap=na + 8 * ka, g=wa.g[:na], step=wa.step)
    t_factor = time.perf_counter() - t0
    ts = []
    for i in range(4):
        m_ = na + (i + 1) * ka
        t0 = time.perf_counter()
        f.append(ens, st, wa.t[:m_], wa.y1[:m_], ka, g=wa.g[:m_], step=wa.step)
        ts.append(time.perf_counter() - t0)
    t_app = float(np.median(ts))
    return {"factor_per_s": Bs / t_factor, "appends_per_s": Bs / t_app, "unit": "particles/s", "cores": f.num_threads(),
            "kind": "port", "append_GBps": Bs * 4.0 * na * (na + 1) / t_app / 1e9,
            "sample": f"first {Bs} of the 256 particles (n = 2048): from-scratch factorisation {t_factor * 1e3:.0f} ms, "
                      f"k = 1 rank-append {t_app * 1e3:.1f} ms per {Bs} particles, OpenMP over particles"}


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
    ap.add_argument("--variant", type=int, default=0, help="factorisation kernel: 0 auto, 2 tile kernel, 3 slot kernel, 4 large-path kernel")
    ap.add_argument("--c4", action="store_true", help="run the BASELINE configs[3] block at any N (default: only at N = 8)")
    ap.add_argument("--no-c4", action="store_true")
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
    d_logw = torch.empty((K, P), dtype=torch.float64, device=dev)     # kernel-only / fast-path runs
    d_mu = torch.empty((K, P, h), dtype=torch.float64, device=dev)
    d_L = torch.empty((K, P, h, h), dtype=torch.float64, device=dev)
    d_info = torch.zeros((K, P), dtype=torch.int32, device=dev)
    d_x = torch.empty((K * D, h), dtype=torch.float64, device=dev)
    d_ens_theta, d_ens_noise = to_dev(w.ens.theta), to_dev(w.ens.noise)
    ens_dev = FlatEnsemble(w.ens.prog, w.ens.prog_off, d_ens_theta, w.ens.theta_off, d_ens_noise)
    from nowcastautogp_b200.sharding import sharded_forecast

    def compute_series(sl, x_out, lw_out):
        # this rank's series, all K scenarios: outputs straight into the packed send buffer of the gather
        eng.forecast_instances(ens_dev, n, k, h, w.t, d_y1, d_y2, d_logw0, w.ya, w.yb, g=w.g, step=w.step,
                               theta=d_theta, noise=d_noise, K=K, logw=lw_out, mu=d_mu, L=d_L, info=d_info)
        eng.draw(lw_out, d_mu, d_L, d_zeta, u=d_u, x=x_out, want_aux=False)

    res = {}

    def step_device():
        # the product multi-GPU path: series-first partition (one series per rank here), per-rank compute, ONE packed
        # all-gather of draws and log-weights (NCCL over NVLink when world > 1; no collective at world = 1)
        res["x"], res["lw"] = sharded_forecast(compute_series, world, K, h, D, P, device=dev, in_place=True)

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
    d2h = sum(t_.numel() * t_.element_size() for t_ in (h_x, h_logw, h_info))

    def step_e2e():
        # ONE C-ABI call with host buffers (nagp_forecast_with_nowcasts_theta): H2D of every input, fused instances, ESS +
        # draws, D2H of draws, log-weights and info
        eng.forecast_with_nowcasts_theta(h_ens, n, k, h, w.t, h_y1, h_y2, h_logw0, h_zeta, h_theta, h_noise, w.ya, w.yb,
                                         g=w.g, step=w.step, u=h_u, x=h_x, logw=h_logw, info=h_info, K=K, D=D)

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

    def timed_local(fn, steps, warmup):
        """CUDA-event time of this rank alone (no barrier, no max over ranks)."""
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total / steps

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
        # fit_smc!-shaped schedule (make_and_fit_model.jl:88-91): 20 cumulative data counts up to n = 2048, un-rejuvenated —
        # every step after the first extends the stored factors by the new block of rows (what GPModel.fit_smc does with
        # n_mcmc = 0), against re-factoring from scratch at every step
        sched = [int(round(na * (i + 1) / 20)) for i in range(20)]
        # best of three for both arms: the appended arm allocates and frees its 4.3 GB store inside the timed region and
        # makes 20 blocking calls, so a single run moves by tens of ms with the state of the allocator and the host
        ms_sched_app, ms_sched_scratch, lm_app = float("inf"), float("inf"), None
        for _rep in range(3):
            e0.record(stream)
            f2 = eng.factor_store_large(wa.ens, wa.t[:sched[0]], wa.y1[:sched[0]], capacity=na, g=wa.g[:sched[0]], step=wa.step, check=False)
            prev = sched[0]
            for c_ in sched[1:]:
                eng.factor_append(f2, wa.t[prev:c_], wa.y1[prev:c_], g_new=wa.g[prev:c_], check=False)
                prev = c_
            e1.record(stream); e1.synchronize()
            ms_sched_app = min(ms_sched_app, e0.elapsed_time(e1))
            lm_app = np.array(f2.logml_n, copy=True)
            f2.free()
            e0.record(stream)
            for c_ in sched:
                eng.logml_batch(ens_a, wa.t[:c_], d_ya[:c_], g=wa.g[:c_], step=wa.step, logml=d_lma, info=d_infa)
            e1.record(stream); e1.synchronize()
            ms_sched_scratch = min(ms_sched_scratch, e0.elapsed_time(e1))
        sched_rel = float(np.abs(lm_app - d_lma.cpu().numpy()).max() / np.abs(lm_app).max())
        ms_app = float(np.median(tms[1:]))
        bytes_ = Pa * 4.0 * na * (na + 1)          # the stored factor read once
        fl_ = Pa * (na ** 3 / 3.0 + 2.0 * na * na)
        return {"what": "BASELINE configs[4]: daily series n=2048, 256 particles; from-scratch factorisation "
                        "(nagp_logml_batch, device-resident) and in-place rank-append of k=1 point on the stored factor "
                        "(append time includes the blocking C-ABI call, host->device of the new point and "
                        "device->host of dlogml/logml/info)",
                "ok": ok_, "factor_ms": ms_store,
                "factor_roofline": {"bound": "tensor", "achieved": fl_ / (ms_store * 1e-3) / 1e12, "unit": "TFLOP/s"},
                "smc_schedule": {"what": "20-step linear_schedule to n = 2048, 256 particles, no rejuvenation: factor store + 19 "
                                 "rank-appends of about 102 rows each (includes the 4.3 GB allocation of the store) vs 20 "
                                 "from-scratch factorisations", "appended_ms": ms_sched_app, "from_scratch_ms": ms_sched_scratch,
                                 "final_logml_rel_diff": sched_rel},
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
        # nominal FLOPs of the per-entry differentiation pass, from the SOURCE programs (what a direct reverse-mode sweep of
        # every entry of the triangle would execute; the kernel does less: stationary sub-trees are folded into lag tables
        # and additive leaves are summarised per lag): forward + backward per op, plus 4 for the entry's weight
        per_op = {1: 1, 2: 12, 3: 14, 4: 30, 5: 34, 6: 2, 7: 3, 8: 26}
        rev_ = sum(4 + sum(per_op[int(o)] for o in w.ens.prog[w.ens.prog_off[p_]:w.ens.prog_off[p_ + 1]]) for p_ in range(P))
        fl_rev = K * rev_ * (m * (m + 1) / 2.0)
        return {"what": "logML + d logML/d(theta, noise) for K*P = 32000 per-scenario HMC chains at n+k=151 "
                        "(one leapfrog stage of mcmc_parameters! on every chain): tile kernel keeps L, gradient kernel "
                        "forms K^-1 in place on the FP64 tensor pipe and differentiates the tree per lag",
                "value": K * P * world / (ms * 1e-3), "unit": "gradient evals/s", "ms_per_step": ms, "ok": ok_,
                "roofline": {"bound": "tensor", "achieved": fl_ / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                             "flops_per_launch": fl_, "note": "m^3/3 (Cholesky) + 2m^3/3 (inverse) per instance; "
                             "the per-entry reverse-mode work is not counted",
                             "reverse_mode_flops_nominal": fl_rev,
                             "achieved_with_reverse_mode": (fl_ + fl_rev) / (ms * 1e-3) / 1e12}}

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
        kernel_only()
        eng.draw(d_logw, d_mu, d_L, d_zeta, u=d_u, x=d_x, want_aux=False)
        ms = timed(run, max(3, args.steps // 2), 3)
        return {"what": "inverse 'positive' transformation in place + 25/50/75 % row quantiles of the (9, 20000) draws",
                "ms_per_step": ms, "bytes_per_launch": int(h * K * D * (8 + 16 + 2 * 3 * 8 * 8))}

    # ---- BASELINE configs[3] (the north_star target): 53 series x 64 particles x 1000 nowcasts, sharded over the ranks --------
    def run_c4():
        from nowcastautogp_b200 import synthetic as syn
        from nowcastautogp_b200.sharding import partition
        S4, n4, k4, h4, P4, K4, D4 = 53, 150, 1, 4, 64, 1000, 20
        parts = partition(S4, K4, world)                           # scenario-shared regime: whole series per rank
        parts_split = partition(S4, K4, world, split_series=True)  # per-scenario regime: the pair list cut evenly
        mine = sorted({sl.series for sl in parts[rank]} | {sl.series for sl in parts_split[rank]})
        ser = {}
        for s_ in mine:
            ws = syn.make_workload(n4, k4, h4, K4, P4, seed=1000 + s_, max_depth=4)
            th_, nz_ = syn.perturbed_theta(ws.ens, K4, seed=5000 + s_)
            rg = np.random.default_rng(9000 + s_)
            ser[s_] = dict(w=ws, th=to_dev(th_), nz=to_dev(nz_), y1=to_dev(ws.y1), y2=to_dev(ws.y2), lw0=to_dev(ws.logw0),
                           zeta=to_dev(rg.standard_normal((K4, D4, h4))), u=to_dev(rg.uniform(size=(K4, D4))),
                           ens=FlatEnsemble(ws.ens.prog, ws.ens.prog_off, to_dev(ws.ens.theta), ws.ens.theta_off,
                                            to_dev(ws.ens.noise)), th_host=th_, nz_host=nz_)
        mu4 = torch.empty((K4, P4, h4), dtype=torch.float64, device=dev)
        L4 = torch.empty((K4, P4, h4, h4), dtype=torch.float64, device=dev)
        inf4 = torch.zeros((K4, P4), dtype=torch.int32, device=dev)

        def comp_per_scenario(sl, x_out, lw_out):
            d_ = ser[sl.series]; ws = d_["w"]
            a_, b_ = sl.k0, sl.k1
            eng.forecast_instances(d_["ens"], n4, k4, h4, ws.t, d_["y1"], d_["y2"][a_:b_], d_["lw0"], ws.ya, ws.yb, g=ws.g,
                                   step=ws.step, theta=d_["th"][a_:b_], noise=d_["nz"][a_:b_], K=b_ - a_, logw=lw_out,
                                   mu=mu4[:b_ - a_], L=L4[:b_ - a_], info=inf4[:b_ - a_])
            eng.draw(lw_out, mu4[:b_ - a_], L4[:b_ - a_], d_["zeta"][a_:b_], u=d_["u"][a_:b_], x=x_out, want_aux=False)

        def comp_shared(sl, x_out, lw_out):
            d_ = ser[sl.series]; ws = d_["w"]
            eng.forecast_with_nowcasts(d_["ens"], n4, k4, h4, ws.t, d_["y1"], d_["y2"], d_["lw0"], d_["zeta"], ws.ya, ws.yb,
                                       g=ws.g, step=ws.step, u=d_["u"], x=x_out, logw=lw_out, info=inf4[0], K=K4, D=D4)

        def local_only(comp, parts_):
            # this rank's compute without the gather: per-rank time, for the load-imbalance figure
            xs = torch.empty((K4 * D4, h4), dtype=torch.float64, device=dev)
            ls = torch.empty((K4, P4), dtype=torch.float64, device=dev)
            def run():
                for sl in parts_[rank]:
                    kk = sl.k1 - sl.k0
                    comp(sl, xs[:kk * D4], ls[:kk])
            return run

        def per_rank(ms_local):
            if world == 1:
                return [ms_local]
            t_ = torch.tensor([ms_local], dtype=torch.float64, device=dev)
            out_ = torch.empty(world, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out_, t_)
            return [float(v) for v in out_.cpu()]

        out = {"what": "BASELINE configs[3]: 53 series x 64 particles x 1000 nowcast scenarios (n=150, k=1, h=4, D=20), series-first "
                       "partition over the ranks (sharding.partition), one packed all-gather per step (sharding.sharded_forecast)",
               "series_per_rank": [len(p_) for p_ in parts],
               "pairs_per_rank_split": [sum(sl.k1 - sl.k0 for sl in p_) for p_ in parts_split], "draws_per_step": S4 * K4 * D4}
        steps4 = max(2, min(args.steps, 5))
        fl_inst = flops_per_instance(n4, k4, h4)
        for name, comp, split in (("per_scenario_theta", comp_per_scenario, True), ("scenario_shared", comp_shared, False)):
            parts_ = parts_split if split else parts
            ms = timed(lambda: sharded_forecast(comp, S4, K4, h4, D4, P4, device=dev, in_place=True, split_series=split),
                       steps4, 3)
            ms_loc = per_rank(timed_local(local_only(comp, parts_), steps4, 2))
            pairs_ = [sum(sl.k1 - sl.k0 for sl in p_) for p_ in parts_]
            blk = {"ms_per_step": ms, "value": S4 * K4 * D4 / (ms * 1e-3), "unit": UNIT, "per_rank_compute_ms": ms_loc,
                   "imbalance_max_over_mean": max(ms_loc) / (sum(ms_loc) / len(ms_loc)),
                   "partition": "pairs cut evenly (split_series)" if split else "whole series per rank",
                   "partition_efficiency_cap": (S4 * K4 / world) / max(pairs_)}
            if name == "per_scenario_theta":
                fl_busy = max(pairs_) * P4 * fl_inst
                blk["roofline"] = {"bound": "tensor", "unit": "TFLOP/s", "peak": None,
                                   "achieved_busiest_rank": fl_busy / (max(ms_loc) * 1e-3) / 1e12,
                                   "achieved_aggregate": S4 * K4 * P4 * fl_inst / (ms * 1e-3) / 1e12 / world,
                                   "flops_per_instance": fl_inst}
            else:
                m4, q4 = n4 + k4, n4 + k4 + h4
                fl_s = P4 * (q4 ** 3 / 3.0 + h4 * m4 * m4 + h4 * h4 * m4) + K4 * P4 * (2 * k4 * m4 + 2 * h4 * k4) + K4 * D4 * h4 * h4
                blk["roofline"] = {"bound": "launch latency", "unit": "GFLOP/s", "flops_per_series": fl_s,
                                   "achieved_aggregate": S4 * fl_s / (ms * 1e-3) / 1e9,
                                   "note": "SURVEY 8(d) algorithmic work of the n_hmc == 0 schedule: 64 factorisations and O(k m) per "
                                           "scenario, three launches per series: bounded by launch latency, not by a pipe"}
            inf_ok = int(inf4.abs().max().item()) == 0
            blk["ok"] = inf_ok
            out[name] = blk
        # parity of the C4 batch: 16 instances of this rank's first series against the oracle
        if mine:
            from oracle.oracle import Oracle
            o = Oracle()
            s0 = mine[0]; d_ = ser[s0]; ws = d_["w"]
            lw_t = torch.empty((K4, P4), dtype=torch.float64, device=dev)
            eng.forecast_instances(d_["ens"], n4, k4, h4, ws.t, d_["y1"], d_["y2"], d_["lw0"], ws.ya, ws.yb, g=ws.g, step=ws.step,
                                   theta=d_["th"], noise=d_["nz"], K=K4, logw=lw_t, mu=mu4, L=L4, info=inf4)
            torch.cuda.synchronize()
            lw_h, mu_h, L_h = lw_t.cpu().numpy(), mu4.cpu().numpy(), L4.cpu().numpy()
            rg = np.random.default_rng(77 + rank)
            off = ws.ens.theta_off
            worst = 0.0
            refs, gots = [], []
            for _ in range(16):
                a_, p_ = int(rg.integers(K4)), int(rg.integers(P4))
                prog = bytes(ws.ens.prog[ws.ens.prog_off[p_]:ws.ens.prog_off[p_ + 1]])
                r_ = o.instance_joint(prog, d_["th_host"][a_, off[p_]:off[p_ + 1]], d_["nz_host"][a_, p_], n4, k4, h4, ws.t,
                                      np.concatenate([ws.y1, ws.y2[a_]]), ws.ya, ws.yb, g=ws.g, step=ws.step)
                refs.append(np.concatenate([[ws.logw0[p_] + r_["logml_m"] - r_["logml_n"]], r_["mu"], r_["L"].ravel()]))
                gots.append(np.concatenate([[lw_h[a_, p_]], mu_h[a_, p_], L_h[a_, p_].ravel()]))
            refs, gots = np.asarray(refs), np.asarray(gots)
            for lo, hi in ((0, 1), (1, 1 + h4), (1 + h4, refs.shape[1])):
                worst = max(worst, float(np.abs(gots[:, lo:hi] - refs[:, lo:hi]).max() / np.abs(refs[:, lo:hi]).max()))
            pr_ = worst
        else:
            pr_ = 0.0
        if world > 1:
            pt_ = torch.tensor([pr_], dtype=torch.float64, device=dev)
            dist.all_reduce(pt_, op=dist.ReduceOp.MAX)
            pr_ = float(pt_.item())
        out["parity_max_rel"] = pr_
        if not pr_ < 1e-9:
            raise SystemExit(f"bench c4: device disagrees with the CPU oracle ({pr_:.3e})")
        try:
            out["api"] = run_c4_api(S4, n4, k4, h4, P4, K4, D4, set(mine))
        except Exception as e:      # noqa: BLE001 - the device-level block above stands on its own
            if world > 1:
                raise               # a rank that drops out of the collectives would hang the others
            out["api"] = {"error": f"{type(e).__name__}: {e}"}
        return out

    def run_c4_api(S4, n4, k4, h4, P4, K4, D4, mine_set):
        """The same configuration through the PUBLIC API: GPModel objects (prior-sampled particle sets that absorbed their
        series in one SMC step), TData scenarios, `forecast_with_nowcasts_sharded` with n_hmc = 0 and n_hmc = 1."""
        import nowcastautogp_b200 as nag
        from nowcastautogp_b200 import synthetic as syn
        from nowcastautogp_b200.gpmodel import GPModel
        t_build = time.perf_counter()
        d0 = np.datetime64("2022-10-01")
        dates = d0 + 7 * np.arange(n4 + k4 + h4)
        models, nowcasts = [], []
        for s_ in range(S4):
            _, raw = syn.weekly_series(n4, 1000 + s_ + 1)
            rg = np.random.default_rng([2026, s_])
            m_ = GPModel(dates[:n4], np.log(raw), n_particles=P4, rng=rg, engine=eng)
            m_.fit_smc(schedule=[n4], n_mcmc=0, n_hmc=0, shuffle=False)
            models.append(m_)
            scen = raw[-1] * np.exp(0.1 + 0.027 * rg.standard_normal((k4, K4)))
            nowcasts.append(nag.create_nowcast_data(scen, dates[n4:n4 + k4], transformation=np.log)
                            if s_ in mine_set else [None] * K4)
        build_s = time.perf_counter() - t_build
        fdates = dates[n4 + k4:]
        out = {"what": "forecast_with_nowcasts_sharded (public API, host objects in, NumPy matrices out on every rank)",
               "build_s": build_s}
        for name, kw in (("n_hmc_0", dict(n_hmc=0)), ("n_hmc_1", dict(n_hmc=1))):
            resd = {}
            def call():
                resd["d"], resd["lw"] = nag.forecast_with_nowcasts_sharded(models, nowcasts, fdates, D4, device=dev, **kw)
            ms = timed(call, 2 if name == "n_hmc_1" else 3, 1, do_flush=False)
            ok_ = all(np.isfinite(v).all() and v.shape == (h4, K4 * D4) for v in resd["d"].values()) and len(resd["d"]) == S4
            out[name] = {"ms_per_step": ms, "value": S4 * K4 * D4 / (ms * 1e-3), "unit": UNIT, "ok": bool(ok_)}
        return out

    micro = None if args.no_micro else {"logml": micro_logml(), "append": micro_append(), "grad": micro_grad(),
                                        "hmc": micro_hmc(), "summary": micro_summary()}

    c4 = run_c4() if (args.c4 or (world == 8 and not args.no_c4)) else None
    if c4 is not None and rank == 0:
        pk = 37.1
        try:
            pk = float(json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json"))).get("dmma_m8n8k4_tflops", 37.1))
        except OSError:
            pass
        rl = c4["per_scenario_theta"]["roofline"]
        rl["peak"] = pk
        rl["frac_busiest_rank"] = rl["achieved_busiest_rank"] / pk
        rl["frac_aggregate"] = rl["achieved_aggregate"] / pk

    # sanity + parity of the timed batch: the step produced finite draws and no factorisation failed, and 64 random
    # (scenario, particle) instances of THIS batch agree with the CPU oracle (checker only) within 1e-9 relative
    step_device()
    torch.cuda.synchronize()
    x_mine = res["x"][rank]
    ok = bool(torch.isfinite(x_mine).all().item()) and int(d_info.abs().max().item()) == 0
    if not ok and not args.skip_sanity:
        raise SystemExit("bench step produced non-finite draws or a failed factorisation")

    def parity_spot_check(n_check=64):
        from oracle.oracle import Oracle
        from nowcastautogp_b200 import kernels as kn_
        o = Oracle()
        rng_ = np.random.default_rng(12345 + rank)
        lw_dev = res["lw"][rank].cpu().numpy()
        mu_dev, L_dev = d_mu.cpu().numpy(), d_L.cpu().numpy()
        off = w.ens.theta_off
        worst = 0.0
        picks = [(int(rng_.integers(K)), int(rng_.integers(P))) for _ in range(n_check)]
        ref = {"logw": [], "mu": [], "L": []}
        got = {"logw": [], "mu": [], "L": []}
        for s_, p_ in picks:
            prog = bytes(w.ens.prog[w.ens.prog_off[p_]:w.ens.prog_off[p_ + 1]])
            y = np.concatenate([w.y1, w.y2[s_]])
            r_ = o.instance_joint(prog, theta_k[s_, off[p_]:off[p_ + 1]], noise_k[s_, p_], n, k, h, w.t, y, w.ya, w.yb,
                                  g=w.g, step=w.step)
            if r_["info"] != 0:
                return float("inf")
            ref["logw"].append(w.logw0[p_] + r_["logml_m"] - r_["logml_n"]); got["logw"].append(lw_dev[s_, p_])
            ref["mu"].append(r_["mu"]); got["mu"].append(mu_dev[s_, p_])
            ref["L"].append(r_["L"]); got["L"].append(L_dev[s_, p_])
        for key in ref:
            a_, b_ = np.asarray(got[key], float), np.asarray(ref[key], float)
            worst = max(worst, float(np.abs(a_ - b_).max() / max(np.abs(b_).max(), 1e-300)))
        return worst

    parity = parity_spot_check() if not args.skip_sanity else None
    if world > 1:
        pt = torch.tensor([parity if parity is not None else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        parity = float(pt.item()) if parity is not None else None
    if parity is not None and not parity < 1e-9:
        raise SystemExit(f"bench: the timed batch disagrees with the CPU oracle (max relative error {parity:.3e} > 1e-9)")

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
        traffic, traffic_src = read_traffic()
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
            "parity_max_rel": parity,
            "parity_check": "64 random (scenario, particle) instances per rank of the timed batch: log-weight, mu and L33 vs the "
                            "CPU oracle's joint factorisation, max relative error (run fails above 1e-9)",
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA; tcgen05 has no f64 kind)",
                         "kernel": "fused Gram+Cholesky+solve (nagp_fused)", "achieved": achieved,
                         "peak": peak_tf, "peak_source": "profiles/r01_fp64_peak.json (DMMA m8n8k4 measured on this "
                         "pool's B200; MEASURED_PEAKS.json has no FP64 entry)", "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch; the Gram and the factor never "
                                         "leave shared memory, and the 23 MB of outputs are still in L2 when the counter is read",
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
            micro["grad"]["roofline"].update(peak=peak_tf, frac=micro["grad"]["roofline"]["achieved"] / peak_tf,
                                             frac_with_reverse_mode=micro["grad"]["roofline"]["achieved_with_reverse_mode"] / peak_tf)
            line["hmc_gradient_microbench"] = micro["grad"]
            line["hmc_microbench"] = micro["hmc"]
            line["summary_microbench"] = micro["summary"]
            line["logml_microbench"] = micro["logml"]
            line["append_microbench"] = micro["append"]
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_reference(w, theta_k, noise_k, zeta, u, steps=1, warmup=0)
            line["cpu_baseline"] = cb
            if micro is not None:
                line["logml_microbench"]["cpu_baseline"] = cpu_logml_baseline(rank)
                line["append_microbench"]["cpu_baseline"] = cpu_append_baseline(rank)
        if c4 is not None:
            if not args.no_cpu_baseline:
                from nowcastautogp_b200 import synthetic as syn
                c4cfg = dict(n=150, k=1, h=4, P=64, K=1000, D=20)
                w4 = syn.make_workload(150, 1, 4, 1000, 64, seed=1000, max_depth=4)
                th4, nz4 = syn.perturbed_theta(w4.ens, 1000, seed=5000)
                rg4 = np.random.default_rng(9000)
                cb4, _, _ = cpu_reference(w4, th4, nz4, rg4.standard_normal((1000, 20, 4)), rg4.uniform(size=(1000, 20)),
                                          steps=1, warmup=0, target_s=8.0, cfg=c4cfg)
                c4["cpu_baseline"] = cb4
            line["c4"] = c4
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
