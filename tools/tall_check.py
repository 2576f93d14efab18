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


def rinvp_off(n):
    """Offset of the fragment-ordered R1^-1 inside `fac` (FacLayout)."""
    return (3 * n * n + 2 * n + 2 + 1) // 2 * 2


def qr_tall(lib, J, f, nranks=1):
    """CholeskyQR2 through the C ABI; rows split into `nranks` shards to
    exercise the rank-partial path on one GPU."""
    m, n = J.shape
    dev = J.device
    f64 = torch.float64
    gs = n * n + n + 1
    work = torch.empty(lib._dll.blsq_tall_gram_work_size(n), dtype=f64, device=dev)
    fac = torch.zeros(lib._dll.blsq_tall_fac_size(n), dtype=f64, device=dev)
    recs = torch.empty((nranks, gs), dtype=f64, device=dev)
    st = lib.stream(J)
    bounds = [(m * r // nranks) // 2 * 2 for r in range(nranks)] + [m]
    for p in (1, 2):
        for r in range(nranks):
            a, b = bounds[r], bounds[r + 1]
            lib.call("blsq_tall_gram", p, b - a, n, J[a:b].data_ptr(), f[a:b].data_ptr(),
                     fac[rinvp_off(n):].data_ptr(), work.data_ptr(), recs[r].data_ptr(), st)
        lib.call("blsq_tall_factor", p, n, nranks, gs, recs.data_ptr(), fac.data_ptr(), st)
    n2 = n * n
    return dict(R=fac[n2:2 * n2].view(n, n), qtf=fac[3 * n2:3 * n2 + n],
                g=fac[3 * n2 + n:3 * n2 + 2 * n], obj=fac[3 * n2 + 2 * n],
                info=fac[3 * n2 + 2 * n + 1], R1=fac[:n2].view(n, n))


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


def check(lib, m, n, cond, nranks=1):
    dev = torch.device("cuda:0")
    J, f = make(m, n, cond, dev)
    out = qr_tall(lib, J, f, nranks)
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
               R=rel(out["R"], R), qtf=rel(out["qtf"], qtf), g=rel(out["g"], g),
               obj=abs(float(out["obj"]) - float(f @ f)) / float(f @ f),
               gn_step=rel(p_got, p_ref),
               lower_zero=float(out["R"].tril(-1).abs().max()))
    print(json.dumps(res))
    return res


def timing(lib, m, n, reps=5):
    dev = torch.device("cuda:0")
    f64 = torch.float64
    J = torch.randn((m, n), dtype=f64, device=dev)
    f = torch.randn(m, dtype=f64, device=dev)
    gs = n * n + n + 1
    work = torch.empty(lib._dll.blsq_tall_gram_work_size(n), dtype=f64, device=dev)
    fac = torch.zeros(lib._dll.blsq_tall_fac_size(n), dtype=f64, device=dev)
    rec = torch.empty((1, gs), dtype=f64, device=dev)
    st = lib.stream(J)

    def one(p):
        lib.call("blsq_tall_gram", p, m, n, J.data_ptr(), f.data_ptr(),
                 fac[rinvp_off(n):].data_ptr(), work.data_ptr(), rec.data_ptr(), st)

    def fact(p):
        lib.call("blsq_tall_factor", p, n, 1, gs, rec.data_ptr(), fac.data_ptr(), st)

    for p in (1, 2):
        one(p)
        fact(p)
    torch.cuda.synchronize()
    t = {}
    for name, fn in (("gram1", lambda: one(1)), ("factor1", lambda: fact(1)),
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
    tot = t["gram1"] + t["gram2"]
    print(json.dumps(dict(m=m, n=n, ms=t, qr_tflops_algorithmic=flops / tot / 1e9,
                          frac_of_37=flops / tot / 1e9 / 37.0,
                          gram1_gbs=8.0 * m * (n + 1) / t["gram1"] / 1e6,
                          gram2_gbs=8.0 * m * (n + 1) / t["gram2"] / 1e6)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--big", action="store_true")
    a = ap.parse_args()
    lib = get_lib()
    worst = 0.0
    for (m, n, cond, nr) in ((4096, 16, 10.0, 1), (20000, 64, 10.0, 1), (20001 * 2, 64, 1e3, 3),
                             (100000, 64, 1e5, 1), (1 << 20, 64, 30.0, 2), (65536 + 38, 32, 1e2, 1),
                             (50000, 128, 1e2, 1), (30000, 256, 1e2, 1), (9000, 200, 10.0, 2),
                             (640, 64, 2.0, 1), (100, 10, 2.0, 1)):
        r = check(lib, m, n, cond, nr)
        worst = max(worst, r["gn_step"])
    print("worst gn_step rel err", worst)
    if a.time:
        timing(lib, 1 << 22, 64)
        timing(lib, 1 << 24, 64)
        timing(lib, 1 << 22, 16)
        timing(lib, 1 << 21, 128)
        timing(lib, 1 << 20, 256)
