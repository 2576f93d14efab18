#!/usr/bin/env python
"""Cycles per phase of trf_round_impl (one thread per problem) from a variant
of the library built with -DBLSQ_PHASE_CLOCKS (tools/build_variants.sh phase
"-DBLSQ_PHASE_CLOCKS"): B = 1 solves (no warp divergence: pure single-thread
latency) and one B = 4096 batch (what a tail round of the C2 batch sees).

    python tools/phase_probe.py [--lib tools/variants/lib_phase.so]
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
from bounded_lsq_b200 import _lib as L, least_squares_batched, PerProblem   # noqa: E402
from bounded_lsq_b200 import synthetic                                      # noqa: E402

PHASES = ["load+judge+adopt", "scaling/CL", "hat fold", "GN shortcut taken",
          "SVD route: fold copy", "SVD route: Jacobi + finish", "SVD route: LM solve",
          "step selection", "store"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "tools/variants/lib_phase.so"))
    a = ap.parse_args()
    lib = L.Lib(a.lib)
    dev = torch.device("cuda:0")
    model = synthetic.ExpDecay2()
    read = lib._dll.blsq_debug_phase_read
    buf = (C.c_ulonglong * 32)()

    def phases(reset=True):
        torch.cuda.synchronize()
        assert read(buf, 1 if reset else 0) == 0
        v = np.array(list(buf), dtype=np.float64)
        return {PHASES[k]: {"visits": int(v[16 + k]),
                            "cycles_per_visit": round(v[k] / max(v[16 + k], 1), 1)}
                for k in range(len(PHASES))}

    def solve(B, seed):
        _, y = model.make_data(B, seed=seed)
        yt = torch.from_numpy(y).to(dev)
        X0 = torch.from_numpy(np.tile(model.x0, (B, 1))).to(dev)
        return least_squares_batched(
            model.fun_t, X0, jac=model.jac_t, bounds=(model.lb, model.ub), method="trf",
            args=(PerProblem(yt),), options=dict(graph_tail_rounds=0), _lib=lib)

    solve(8, 0)
    phases()
    nf = 0
    for seed in range(48):
        nf += int(solve(1, 100 + seed).nfev[0])
    print(json.dumps({"case": "48 solves with B = 1 (single-thread latency)", "nfev": nf,
                      "phases": phases()}))
    r = solve(4096, 7)
    print(json.dumps({"case": "B = 4096 (warps of 32 problems: divergence included)",
                      "mean_nfev": float(r.nfev.double().mean()), "phases": phases()}))


if __name__ == "__main__":
    main()
