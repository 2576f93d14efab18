#!/usr/bin/env python
"""Headline benchmark: bounded fits solved per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3]
    python bench.py --impl reference ...     # CPU arm (oracle port, all cores)

A *step* is one complete batched solve of the workload: every problem is taken
from x0 to termination by the lock-step driver (user callbacks in PyTorch +
the two CUDA kernels per round behind include/blsq.h).

  value  fits/s with the data (y, x0, bounds) already resident in HBM
  e2e    fits/s through the public API with HOST buffers: pinned y/x0 are
         copied to the device and x/status/cost copied back inside the timing
  roofline  the HBM-bound linearisation kernel (QR of [J | f] per problem):
         algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (NumPy port of the reference) on the host cores,
         bounded sample of the same workload

With N > 1 (torchrun) the problems are split by index, no collective on the
data path ("weak": per-GPU batch fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(desc="batched TRF: 1M bounded 4-param exponential-decay fits, "
                    "m=64, analytic Jacobian", model="ExpDecay2", B=1_000_000,
               method="trf", jac="exact", n=4, m=64),
    # BASELINE.json configs[2] (10M problems; 61 GB of residual evaluations
    # per FD sweep, so the per-GPU batch is processed in chunks)
    "c3": dict(desc="batched dogbox: 10M bounded 6-param Gaussian-peak fits, "
                    "m=128, 2-point Jacobian", model="GaussPeak", B=10_000_000,
               method="dogbox", jac="2-point", n=6, m=128, chunk=1_000_000),
}


def _model(name):
    from bounded_lsq_b200 import synthetic
    return getattr(synthetic, name)()


# ------------------------------------------------------------ CPU arm -----

def _cpu_worker(job):
    name, method, jac, seed, count = job
    from oracle import blsq_oracle as orc
    model = _model(name)
    _, y = model.make_data(count, seed=seed)
    t0 = time.perf_counter()
    nfev = 0
    for b in range(count):
        if jac == "exact":
            r = orc.least_squares(model.fun_np, model.x0, jac=model.jac_np,
                                  bounds=(model.lb, model.ub), method=method,
                                  args=(y[b],))
        else:
            r = orc.least_squares(model.fun_np, model.x0, jac="2-point",
                                  bounds=(model.lb, model.ub), method=method,
                                  args=(y[b],))
        nfev += r.nfev
    return count, nfev, time.perf_counter() - t0


def cpu_fits_per_second(w, per_core, cores=None, seed0=1000):
    """Oracle (NumPy port of the reference) over all host cores."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    jobs = [(w["model"], w["method"], w["jac"], seed0 + c, per_core)
            for c in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        out = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    fits = sum(o[0] for o in out)
    return fits / wall, cores, fits, sum(o[1] for o in out) / fits


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # a step is a bounded sample: ~6 s of CPU work per core, so that pool
    # start-up (fork + imports) stays a small part of the timed region
    per_core = 256 if w["jac"] == "exact" else 96
    for _ in range(min(args.warmup, 1)):
        cpu_fits_per_second(w, 4, cores)
    t0 = time.perf_counter()
    fits = 0
    nfev = 0.0
    for k in range(args.steps):
        v, c, f, nf = cpu_fits_per_second(w, per_core, cores, seed0=2000 + 97 * k)
        fits += f
        nfev += nf * f
    wall = time.perf_counter() - t0
    value = fits / wall
    line = {
        "impl": "reference", "metric": "bounded fits solved/sec (batched)",
        "value": value, "unit": "fits/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w["desc"], "sample_fits_per_step": per_core * cores},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": cores,
                         "kind": "port",
                         "sample": f"{per_core} fits per core x {cores} cores "
                                   f"per step, oracle/blsq_oracle.py "
                                   f"(NumPy/SciPy restatement of the "
                                   f"reference), mean nfev {nfev / fits:.1f}"},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------ clocks ------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index),
                     "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------ GPU arm -----

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None,
                    help="problems per GPU (default: the workload's size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--callbacks", default="fused", choices=["fused", "torch"],
                    help="residual/Jacobian callbacks: the fused CUDA model "
                         "op shipped for the synthetic workloads, or plain "
                         "torch elementwise ops")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, w)

    import torch
    import torch.distributed as dist
    from bounded_lsq_b200 import least_squares_batched, PerProblem
    from bounded_lsq_b200 import batched as drv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = _model(w["model"])
    B = args.batch or w["B"]            # per GPU ("weak")
    chunk = min(B, w.get("chunk", B))
    n, m = w["n"], w["m"]
    method = w["method"]

    # synthetic data of the named shape, generated on the host once
    # (seeded per rank: different problems on every GPU)
    t_gen = time.perf_counter()
    rng_seed = 10_000 + rank
    ngen = min(B, 262_144)              # tile a 256k-problem pool up to B
    _, ypool = model.make_data(ngen, seed=rng_seed)
    reps = (B + ngen - 1) // ngen
    y_host = torch.from_numpy(np.tile(ypool, (reps, 1))[:B].copy()).pin_memory()
    x0_host = torch.from_numpy(np.tile(model.x0, (B, 1))).pin_memory()
    lb = torch.as_tensor(model.lb, device=dev)
    ub = torch.as_tensor(model.ub, device=dev)
    gen_s = time.perf_counter() - t_gen

    fun, jac = model.fun_t, (model.jac_t if w["jac"] == "exact" else "2-point")
    if args.callbacks == "fused":
        try:
            from bounded_lsq_b200 import models as fused
            fun, jac = fused.callbacks(w["model"], w["jac"])
        except Exception as e:          # the fused op is optional sugar
            if rank == 0:
                print(f"# fused callbacks unavailable ({e!r}); using torch",
                      file=sys.stderr)
            args.callbacks = "torch"

    timers = {}

    def solve(y_dev, x0_dev, collect=None):
        outs = []
        for c0 in range(0, B, chunk):
            yc = y_dev[c0:c0 + chunk]
            res = least_squares_batched(
                fun, x0_dev[c0:c0 + chunk], jac=jac, bounds=(lb, ub),
                method=method, args=(PerProblem(yc),),
                options=dict(timers=collect) if collect is not None else {})
            outs.append(res)
        return outs

    y_dev = y_host.to(dev)
    x0_dev = x0_host.to(dev)
    x_out = torch.empty((B, n), dtype=torch.float64).pin_memory()
    st_out = torch.empty((B,), dtype=torch.int64).pin_memory()
    obj_out = torch.empty((B,), dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing -----------------------------------------
    for _ in range(args.warmup):
        solve(y_dev, x0_dev)
    barrier()
    launches = 0
    rounds = 0
    with ClockSampler(local) as clk:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            outs = solve(y_dev, x0_dev)
            launches += sum(o.kernel_launches for o in outs)
            rounds += sum(o.rounds for o in outs)
        e1.record()
        barrier()
    dev_ms = reduce_max(e0.elapsed_time(e1))
    value = world * B * args.steps / (dev_ms * 1e-3)
    status = torch.cat([o.status for o in outs])
    nfev_mean = float(torch.cat([o.nfev for o in outs]).double().mean())
    njev_mean = float(torch.cat([o.njev for o in outs]).double().mean())
    converged = float((status > 0).double().mean())

    # ---- per-kernel timing for the roofline (separate instrumented step) --
    collect = {}
    torch.cuda.synchronize()
    solve(y_dev, x0_dev, collect)
    torch.cuda.synchronize()
    ksum = drv.summarize_timers(collect)

    # ---- end to end: host buffers in, host results out -------------------
    barrier()
    e2 = torch.cuda.Event(enable_timing=True)
    e3 = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e2.record()
    for _ in range(args.steps):
        yd = y_host.to(dev, non_blocking=True)
        xd = x0_host.to(dev, non_blocking=True)
        outs = solve(yd, xd)
        x_out.copy_(torch.cat([o.x for o in outs]), non_blocking=True)
        st_out.copy_(torch.cat([o.status for o in outs]), non_blocking=True)
        obj_out.copy_(torch.cat([o.obj_value for o in outs]), non_blocking=True)
        torch.cuda.synchronize()
    e3.record()
    barrier()
    e2e_ms = reduce_max(e2.elapsed_time(e3))
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    h2d = y_host.numel() * 8 + x0_host.numel() * 8
    d2h = x_out.numel() * 8 + st_out.numel() * 8 + obj_out.numel() * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else \
        "fallback 6650 GB/s (B200_PROFILING.md)"
    lin = ksum.get("linearise", {})
    rnd = ksum.get("round", {})
    roofline = {
        "bound": "hbm", "kernel": "lin_kernel (blsq_linearise_batched)",
        "achieved": lin.get("gbs"), "peak": peak, "unit": "GB/s",
        "frac": (lin.get("gbs") / peak) if lin.get("gbs") else None,
        "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_problem": lin.get("bytes_per_problem"),
        "avg_launch_ms": lin.get("avg_ms"), "launches": lin.get("launches"),
        "share_of_kernel_time": lin.get("share"),
        "round_kernel": {"avg_launch_ms": rnd.get("avg_ms"),
                         "share_of_kernel_time": rnd.get("share"),
                         "gbs": rnd.get("gbs")},
        "callbacks_ms_per_step": ksum.get("callbacks_ms"),
        "kernels_ms_per_step": ksum.get("kernels_ms"),
    }

    cpu = None
    if not args.no_cpu_baseline:
        per_core = 48 if w["jac"] == "exact" else 24
        v, cores, fits, nf = cpu_fits_per_second(w, per_core)
        cpu = {"value": v, "unit": "fits/s", "cores": cores, "kind": "port",
               "sample": f"{fits} fits ({per_core}/core), oracle/blsq_oracle.py"
                         f" NumPy/SciPy restatement, mean nfev {nf:.1f}"}

    line = {
        "metric": "bounded fits solved/sec (batched)", "value": value,
        "unit": "fits/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "problems_per_gpu": B,
                   "chunk": chunk, "method": method, "jac": w["jac"],
                   "callbacks": args.callbacks, "n": n, "m": m,
                   "l2": "inputs (J+f per round: %.1f GB) exceed the 126 MB L2"
                         % (B * m * (n + 1) * 8 / 1e9),
                   "mean_nfev": nfev_mean, "mean_njev": njev_mean,
                   "converged_frac": converged,
                   "rounds_per_step": rounds / args.steps},
        "e2e": {"value": e2e_value, "unit": "fits/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clk.summary(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
