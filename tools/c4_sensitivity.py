#!/usr/bin/env python
"""How reproducible is the reference on the C4 start?  Re-runs the oracle (bit
identical to the reference, tests/test_oracle_golden.py) with 1-ulp noise on
the residuals and prints how far nfev, x and the per-iteration trial points
move, plus cond(J) along the path.  DESIGN.md section 5 quotes its output.

    python tools/c4_sensitivity.py 4096 16 3 trf
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bounded_lsq_b200.synthetic import TallLinExp
from oracle import blsq_oracle as orc
m, n, seed, method = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
wl = TallLinExp(m, n, seed=seed)
def run(noise, s=0):
    rng = np.random.default_rng(s)
    xs=[]
    def fun(x):
        xs.append(x.copy())
        f = wl.fun_np(x)
        if noise: f = f*(1+noise*rng.standard_normal(f.shape))
        return f
    r = orc.least_squares(fun, wl.x0, jac=wl.jac_np, bounds=(wl.lb, wl.ub), method=method)
    return r, xs[1:]
r0, x0s = run(0)
print("base", r0.status, r0.nfev, r0.njev, repr(r0.obj_value))
for s in range(3):
    r1, x1s = run(2.2e-16, s)
    print("noisy", r1.status, r1.nfev, r1.njev, repr(r1.obj_value), "x_rel %.2e"%(np.abs(r1.x-r0.x).max()/np.abs(r0.x).max()))
    prev=wl.x0; out=[]
    for k in range(min(len(x0s),len(x1s),12)):
        out.append("%.1e"%(np.abs(x0s[k]-x1s[k]).max()/max(np.abs(x0s[k]-prev).max(),1e-300))); prev=x0s[k]
    print("  rel trial diffs:", " ".join(out))
# conditioning along the path
for k in (0,3,5,6,7,10,20):
    if k < len(x0s):
        J = wl.jac_np(x0s[k]); sv = np.linalg.svd(J, compute_uv=False); print(k, "cond %.2e"%(sv[0]/sv[-1]))
