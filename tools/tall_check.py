#!/usr/bin/env python
"""GPU check + timing of the tall-mode CholeskyQR2 kernels (blsq_tall_gram /
blsq_tall_factor) against torch.linalg.qr.  Development tool, run under gpurun.

    python tools/tall_check.py [--time]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bounded_lsq_b200 import get_lib  # noqa: E402


sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import tall_factor  # noqa: E402


def qr_tall(lib, J, f, nranks=1, sstride=1):
    return tall_factor(lib, J, f, nranks, sstride)


def make(m, n, cond, dev, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    J = torch.randn((m, n), dtype=torch.float64, device=dev, generator=g)
    # column mixing with a prescribed condition number
    s = torch.logspace(0, -torch.log10(torch.tensor(float(cond))).item(), n,
                       dtype=torch.float64, device=dev)
    Q, _ = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device=dev, generator=g))
    J = J @ (Q * s) @ Q.T
    f = torch.randn(m, dtype=torch.float64, device=dev, generator=g)
    return J.contiguous(), f


def check(lib, m, n, cond, nranks=1, sstride=1):
    dev = torch.device("cuda:0")
    J, f = make(m, n, cond, dev)
    out = qr_tall(lib, J, f, nranks, sstride)
    torch.cuda.synchronize()
    Q, R = torch.linalg.qr(J)
    sg = torch.sign(torch.diagonal(R))
    R = R * sg[:, None]
    qtf = (Q * sg[None, :]).T @ f
    g = J.T @ f
    rel = lambda a, b: float((a - b).norm() / b.norm())  # noqa: E731
    # quantity the solver consumes: the Gauss-Newton step R^-1 Q^T f
    p_ref = torch.linalg.solve_triangular(R, qtf[:, None], upper=True)[:, 0]
    p_got = torch.linalg.solve_triangular(out["R"], out["qtf"][:, None], upper=True)[:, 0]
    res = dict(m=m, n=n, cond=cond, nranks=nranks, info=float(out["info"]),
               refined=out["refined"],
               R=rel(out["R"], R), qtf=rel(out["qtf"], qtf), g=rel(out["g"], g),
               obj=abs(out["obj"] - float(f @ f)) / float(f @ f),
               gn_step=rel(p_got, p_ref),
               lower_zero=float(out["R"].tril(-1).abs().max()))
    print(json.dumps(res))
    return res


def timing(lib, m, n, reps=5):
    dev = torch.device("cuda:0")
    f64 = torch.float64
    J = torch.randn((m, n), dtype=f64, device=dev)
    f = torch.randn(m, dtype=f64, device=dev)
    lay = lib.tall_layout(n)
    gs = lay["record"]
    work = torch.empty(lay["gram_work"], dtype=f64, device=dev)
    fac = torch.zeros(lay["fac_size"], dtype=f64, device=dev)
    rec = torch.empty((1, gs), dtype=f64, device=dev)
    st = lib.stream(J)
    sstride = int(lib._fn["blsq_tall_sample_stride"](m, n))

    def one(p, ss=1):
        lib.call("blsq_tall_gram", p, m, n, J.data_ptr(), f.data_ptr(),
                 fac[lay["rinvp"]:].data_ptr(), ss, work.data_ptr(), rec.data_ptr(), st)

    def fact(p):
        lib.call("blsq_tall_factor", p, n, 1, gs, rec.data_ptr(), fac.data_ptr(), st)

    for p in (1, 2):
        one(p)
        fact(p)
    torch.cuda.synchronize()
    t = {}
    for name, fn in (("gram1_full", lambda: one(1)), ("gram1_sampled", lambda: one(1, sstride)),
                     ("factor1", lambda: fact(1)),
                     ("gram2", lambda: one(2)), ("factor2", lambda: fact(2))):
        best = 1e9
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        t[name] = best
    flops = 2.0 * m * n * n
    for kind in ("full", "sampled"):
        tot = t["gram1_" + kind] + t["gram2"]
        print(json.dumps(dict(m=m, n=n, pass1=kind, sstride=sstride if kind == "sampled" else 1,
                              ms=t, qr_tflops_algorithmic=flops / tot / 1e9,
                              frac_of_37=flops / tot / 1e9 / 37.0)))


def round_timing(lib, n, m=200000, reps=20):
    """Latency of the n x n tail kernel (blsq_tall_round phase 2) on a C4-like
    factor record: with the SVD (new_lin = 1) and without (rejected trial)."""
    from bounded_lsq_b200.synthetic import TallLinExp
    dev = torch.device("cuda:0")
    f64 = torch.float64
    wl = TallLinExp(m, n, seed=0, x0_tail=(0.8, 1.5, 0.3, 4.0)).to_device(dev)
    x0 = torch.as_tensor(wl.x0, device=dev)
    lb = torch.as_tensor(wl.lb, device=dev)
    ub = torch.as_tensor(wl.ub, device=dev)
    sc = torch.ones(n, dtype=f64, device=dev)
    lay = lib.tall_layout(n)
    J = wl.jac_t(x0).clone()
    f = wl.fun_t(x0)
    # factor record through the public pieces
    gs = lay["record"]
    work = torch.empty(lay["gram_work"], dtype=f64, device=dev)
    fac = torch.zeros(lay["fac_size"], dtype=f64, device=dev)
    rec = torch.empty((1, gs), dtype=f64, device=dev)
    st = lib.stream(J)
    for p in (1, 2):
        lib.call("blsq_tall_gram", p, m, n, J.data_ptr(), f.data_ptr(),
                 fac[lay["rinvp"]:].data_ptr(), 1, work.data_ptr(), rec.data_ptr(), st)
        lib.call("blsq_tall_factor", p, n, 1, gs, rec.data_ptr(), fac.data_ptr(), st)
    state = torch.zeros(lay["state_size"], dtype=f64, device=dev)
    istate = torch.zeros(lay["istate_size"], dtype=torch.int32, device=dev)
    rwork = torch.empty(n * n, dtype=f64, device=dev)
    ssq = (f @ f).reshape(1)

    def rnd(phase, first, new_lin):
        lib.call("blsq_tall_round", 0, phase, n, m, 1, ssq.data_ptr(), fac.data_ptr(),
                 x0.data_ptr(), lb.data_ptr(), ub.data_ptr(), sc.data_ptr(), 1.5e-8, 1.5e-8,
                 1.5e-8, 100 * n, first, new_lin, state.data_ptr(), istate.data_ptr(),
                 rwork.data_ptr(), st)

    rnd(0, 1, 0)
    rnd(1, 1, 0)
    rnd(2, 1, 1)
    torch.cuda.synchronize()
    out = {}
    for name, nl in (("propose_new_lin", 1), ("propose_cached", 0)):
        best = 1e9
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rnd(2, 0, nl)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = best
    print(json.dumps(dict(n=n, status=int(istate[0]), ms=out)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--round", action="store_true", help="latency of the tail kernel")
    ap.add_argument("--round-n", type=int, default=0)
    ap.add_argument("--n256", action="store_true", help="n = 256 / 200 / 130 only")
    ap.add_argument("--time256", action="store_true", help="n = 256 timing only")
    ap.add_argument("--quick", action="store_true",
                    help="one correctness case + the n = 64 timing (variant sweeps)")
    a = ap.parse_args()
    lib = get_lib()
    if a.time256:
        timing(lib, 1 << 20, 256, reps=3)
        sys.exit(0)
    if a.n256:
        check(lib, 30000, 256, 1e2, 1)
        check(lib, 9000, 200, 10.0, 2)
        check(lib, 5003, 130, 10.0, 1)
        timing(lib, 1 << 20, 256, reps=3)
        sys.exit(0)
    if a.quick:
        print("lib", lib.path)
        check(lib, 1 << 20, 64, 1e3, 2, 4)
        timing(lib, 1 << 24, 64)
        sys.exit(0)
    if a.round:
        for n in ([a.round_n] if a.round_n else (16, 64, 128, 256)):
            round_timing(lib, n)
        sys.exit(0)
    worst = 0.0
    for (m, n, cond, nr) in ((4096, 16, 10.0, 1), (20000, 64, 10.0, 1), (20001 * 2, 64, 1e3, 3),
                             (100000, 64, 1e5, 1), (1 << 20, 64, 30.0, 2), (65536 + 38, 32, 1e2, 1),
                             (50000, 128, 1e2, 1), (30000, 256, 1e2, 1), (9000, 200, 10.0, 2),
                             (640, 64, 2.0, 1), (100, 10, 2.0, 1)):
        r = check(lib, m, n, cond, nr)
        worst = max(worst, r["gn_step"])
    for (m, n, cond, nr, ss) in ((1 << 20, 64, 30.0, 2, 4), (1 << 21, 64, 1e4, 1, 8),
                                 (1 << 19, 16, 1e2, 1, 8)):
        r = check(lib, m, n, cond, nr, ss)
        worst = max(worst, r["gn_step"])
    print("worst gn_step rel err", worst)
    if a.time:
        timing(lib, 1 << 22, 64)
        timing(lib, 1 << 24, 64)
        timing(lib, 1 << 22, 16)
        timing(lib, 1 << 21, 128)
        timing(lib, 1 << 20, 256)
