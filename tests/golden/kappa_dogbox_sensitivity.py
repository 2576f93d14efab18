#!/usr/bin/env python
"""How far does the reference algorithm's own answer move under rounding noise
on the kappa = 1e6 dogbox case of check_tall_edge_cases_vs_oracle?

dogbox walks hundreds of iterations along an ill-conditioned valley there and
stops when one step happens to reduce the cost by less than ftol: 1-ulp noise
on f (and J) moves that point anywhere between 195 and 2400 evaluations.  The
ensemble written to kappa_dogbox_sensitivity.json is the gate of the test (the
tall kernels must end no worse than the worst run of the reference itself,
with a 10 % margin).  Oracle = oracle/blsq_oracle.py, bit-identical to the
reference on the golden files.

    python tests/golden/kappa_dogbox_sensitivity.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import blsq_oracle as orc          # noqa: E402


def problem():
    # the generator of check_tall_edge_cases_vs_oracle, same draws in the same order
    rng = np.random.default_rng(12)
    rng.standard_normal((7, 12)); rng.standard_normal((7, 12)); rng.standard_normal(7)
    rng.standard_normal((30, 10))
    m, n, kappa = 3000, 24, 1e6
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    V, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (U * np.logspace(0, -np.log10(kappa), n)) @ V.T
    xt = rng.uniform(-1, 1, n)
    y = A @ xt + 0.1 * np.sin(A @ xt) + 1e-3 * rng.standard_normal(m)
    return A, y, np.full(n, -0.6), np.full(n, 0.7), np.full(n, 0.05)


def main():
    A, y, lb, ub, x0 = problem()

    def run(seed, on_jac):
        nr = None if seed is None else np.random.default_rng(1000 * on_jac + seed)

        def noisy(v):
            return v if nr is None else v * (1 + nr.integers(-1, 2, v.shape) * 1.1e-16)

        def fun(x):
            return noisy(A @ x + 0.1 * np.sin(A @ x) - y)

        def jac(x):
            J = (1 + 0.1 * np.cos(A @ x))[:, None] * A
            return noisy(J) if on_jac else J
        r = orc.least_squares(fun, x0, jac=jac, bounds=(lb, ub), method="dogbox")
        return int(r.status), int(r.nfev), float(r.obj_value)

    base = run(None, 0)
    runs = [run(s, 0) for s in range(16)] + [run(s, 1) for s in range(40)]
    ratios = [r[2] / base[2] for r in runs]
    out = {"case": "m=3000, n=24, kappa=1e6, dogbox, analytic J",
           "unperturbed": {"status": base[0], "nfev": base[1], "cost": base[2]},
           "noise": "1 ulp on f (16 runs), 1 ulp on f and J (40 runs)",
           "nfev_min": min(r[1] for r in runs), "nfev_max": max(r[1] for r in runs),
           "cost_ratio_min": min(ratios), "cost_ratio_max": max(ratios),
           "runs": [[r[0], r[1], round(q, 5)] for r, q in zip(runs, ratios)]}
    with open(os.path.join(HERE, "kappa_dogbox_sensitivity.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "runs"}))


if __name__ == "__main__":
    main()
