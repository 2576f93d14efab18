#!/usr/bin/env python
"""Kernel micro-benchmark: times blsq_linearise_batched and blsq_round_batched
alone on config-C2/C3 shaped data (CUDA events, L2-exceeding inputs).

    python tools/kbench.py [--lib path/to/lib.so] [--B 1000000] [--workload c2]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bounded_lsq_b200 import _lib as L          # noqa: E402
from bounded_lsq_b200 import synthetic          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=L.LIB_PATH)
    ap.add_argument("--B", type=int, default=1_000_000)
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    lib = L.Lib(a.lib)
    dev = torch.device("cuda:0")
    model = synthetic.ExpDecay2() if a.workload == "c2" else synthetic.GaussPeak()
    method = L.METHOD_TRF if a.workload == "c2" else L.METHOD_DOGBOX
    B, n, m = a.B, model.n, model.m
    _, ypool = model.make_data(min(B, 65536), seed=1)
    y = torch.from_numpy(np.tile(ypool, ((B + len(ypool) - 1) // len(ypool), 1))[:B].copy()).to(dev)
    X0 = torch.from_numpy(np.tile(model.x0, (B, 1))).to(dev)
    lb = torch.as_tensor(model.lb, device=dev)
    ub = torch.as_tensor(model.ub, device=dev)
    sc = torch.ones(n, dtype=torch.float64, device=dev)
    lay = lib.state_layout(method, n)
    S, LS = lay["size"], lib.lin_record_size(n)
    state = torch.zeros((B, S), dtype=torch.float64, device=dev)
    istate = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    Xnew = torch.empty((B, n), dtype=torch.float64, device=dev)
    lin = torch.empty((B, LS), dtype=torch.float64, device=dev)
    st = lib.stream(X0)
    lib.call("blsq_init_batched", method, B, n, X0.data_ptr(), lb.data_ptr(),
             ub.data_ptr(), 0, state.data_ptr(), istate.data_ptr(),
             Xnew.data_ptr(), st)
    F = model.fun_t(Xnew, y).contiguous()
    if a.workload == "c2":
        J = model.jac_t(Xnew, y).contiguous()
        mode, plist, dxp = 0, None, None
    else:
        Xp = torch.empty((n, B, n), dtype=torch.float64, device=dev)
        dx = torch.empty((B, n), dtype=torch.float64, device=dev)
        lib.call("blsq_fd2_points", B, None, n, Xnew.data_ptr(), lb.data_ptr(),
                 ub.data_ptr(), 0, float("nan"), Xp.data_ptr(), dx.data_ptr(), st)
        Fp = [model.fun_t(Xp[i], y).contiguous() for i in range(n)]
        arr = (C.c_void_p * n)(*[t.data_ptr() for t in Fp])
        mode, plist, dxp, J = 1, C.cast(arr, C.c_void_p), dx.data_ptr(), None
    RW = torch.empty(B + 1, dtype=torch.int32, device=dev).data_ptr() \
        if os.environ.get('BLSQ_TRF_TWO_KERNELS', '1') != '0' else None
    torch.cuda.synchronize()

    def t_lin():
        lib.call("blsq_linearise_batched", B, None, m, n, F.data_ptr(),
                 None if J is None else J.data_ptr(), plist, dxp, mode,
                 istate.data_ptr(), lin.data_ptr(), st)

    CNT = torch.zeros(4, dtype=torch.int32, device=dev)

    def call_round(first):
        lib.call("blsq_round_batched", method, B, None, m, n, lin.data_ptr(),
                 X0.data_ptr(), lb.data_ptr(), ub.data_ptr(), 0, sc.data_ptr(),
                 1.5e-8, 1.5e-8, 1.5e-8, 100 * n, first, state.data_ptr(),
                 istate.data_ptr(), Xnew.data_ptr(),
                 None if Xjac is None else Xjac.data_ptr(), RW, CNT.data_ptr(), st)

    def timeit(fn, pre=None):
        ts = []
        for r in range(a.reps + 3):
            if pre:
                pre()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(np.min(ts))

    Xjac = torch.empty_like(Xnew) if a.workload != "c2" else None
    lin_ms, lin_min = timeit(t_lin)
    # round 1 (first = 1), then the steady-state round: callbacks at the trial
    # points, linearise, and the judge + propose round timed from a restored
    # copy of the state
    call_round(1)
    Xj = Xnew if Xjac is None else Xjac
    F = model.fun_t(Xnew, y).contiguous()
    if a.workload == "c2":
        J = model.jac_t(Xj, y).contiguous()
    else:
        lib.call("blsq_fd2_points", B, None, n, Xj.data_ptr(), lb.data_ptr(),
                 ub.data_ptr(), 0, float("nan"), Xp.data_ptr(), dx.data_ptr(), st)
        Fp = [model.fun_t(Xp[i], y).contiguous() for i in range(n)]
        arr = (C.c_void_p * n)(*[t.data_ptr() for t in Fp])
        plist = C.cast(arr, C.c_void_p)
    t_lin()
    state0, istate0 = state.clone(), istate.clone()

    def restore():
        state.copy_(state0)
        istate.copy_(istate0)

    rnd_ms, rnd_min = timeit(lambda: call_round(0), restore)
    restore()
    call_round(0)
    torch.cuda.synchronize()
    running = int(CNT[2].item())
    lin_bytes = B * (8 * m * (n + 1) + 8 * LS)
    rnd_bytes = B * (8 * LS + 16 * S + 64 + 8 * n)
    print(json.dumps(dict(
        lib=os.path.basename(a.lib), workload=a.workload, B=B,
        lin_ms=lin_ms, lin_min_ms=lin_min, lin_gbs=lin_bytes / lin_ms / 1e6,
        lin_frac=lin_bytes / lin_ms / 1e6 / 6535.1,
        round_ms=rnd_ms, round_min_ms=rnd_min,
        round_gbs=rnd_bytes / rnd_ms / 1e6, running_after=running)))


if __name__ == "__main__":
    main()
