"""Synthetic workloads named by BASELINE.json `configs`.

Each workload has a NumPy single-problem form (``fun(x)``/``jac(x)`` on one
parameter vector; what the reference solver consumes) and a torch batched form
(``fun(X)``/``jac(X)`` on ``(B, n)`` device tensors; what the B200 path
consumes).  Both evaluate the same formula in the same operation order, so a
GPU fit and a CPU fit of the same problem see residuals that agree to the last
few ulps (``exp`` differs between libm and CUDA by <= 1 ulp).

C2  batched TRF     y = a e^{-b t} + c e^{-d t}                      n=4 m=64
C3  batched dogbox  y = A e^{-((t-mu)/sigma)^2/2} + c0 + c1 t + c2 t^2 n=6 m=128
C4  tall            f = A x[:k] + x_k e^{-x_{k+1} t} + x_{k+2} e^{-x_{k+3} t} - y
"""
from __future__ import annotations

import numpy as np

try:  # torch is only needed for the batched/device forms
    import torch
except Exception:  # pragma: no cover
    torch = None


def _dev_const(model, X):
    """model.t on X's device, uploaded once (a host-to-device copy per
    callback would also make the callback impossible to capture in a CUDA
    graph)."""
    cache = model.__dict__.setdefault("_t_dev", {})
    key = (X.device, X.dtype)
    t = cache.get(key)
    if t is None:
        t = cache[key] = torch.as_tensor(model.t, dtype=X.dtype, device=X.device)
    return t


# ------------------------------------------------------------------ C2 ----

class ExpDecay2:
    """Config C2: double exponential decay, 4 parameters, bounded."""

    n = 4
    x0 = np.array([2.0, 1.0, 1.0, 3.0])
    lb = np.array([0.0, 0.0, 0.0, 0.0])
    ub = np.array([5.0, 2.0, 5.0, 5.0])

    def __init__(self, m=64):
        self.m = m
        self.t = np.linspace(0.0, 4.0, m)

    def make_data(self, B, seed=0, noise=0.01):
        rng = np.random.default_rng(seed)
        a = rng.uniform(1.0, 3.0, B)
        b = rng.uniform(0.5, 1.5, B)
        c = rng.uniform(0.5, 2.0, B)
        d = rng.uniform(2.0, 4.0, B)
        truth = np.stack([a, b, c, d], axis=1)
        t = self.t
        y = (a[:, None] * np.exp(-b[:, None] * t) +
             c[:, None] * np.exp(-d[:, None] * t))
        y = y + noise * rng.standard_normal(y.shape)
        return truth, y

    # single problem, NumPy
    def fun_np(self, x, y):
        t = self.t
        return x[0] * np.exp(-x[1] * t) + x[2] * np.exp(-x[3] * t) - y

    def jac_np(self, x, y=None):
        t = self.t
        e1 = np.exp(-x[1] * t)
        e2 = np.exp(-x[3] * t)
        J = np.empty((self.m, 4))
        J[:, 0] = e1
        J[:, 1] = -x[0] * t * e1
        J[:, 2] = e2
        J[:, 3] = -x[2] * t * e2
        return J

    # batched, torch
    def fun_t(self, X, y):
        t = _dev_const(self, X)
        return (X[:, 0:1] * torch.exp(-X[:, 1:2] * t) +
                X[:, 2:3] * torch.exp(-X[:, 3:4] * t) - y)

    def jac_t(self, X, y=None):
        t = _dev_const(self, X)
        e1 = torch.exp(-X[:, 1:2] * t)
        e2 = torch.exp(-X[:, 3:4] * t)
        J = torch.empty((X.shape[0], self.m, 4), dtype=X.dtype,
                        device=X.device)
        J[:, :, 0] = e1
        J[:, :, 1] = -X[:, 0:1] * t * e1
        J[:, :, 2] = e2
        J[:, :, 3] = -X[:, 2:3] * t * e2
        return J


# ------------------------------------------------------------------ C3 ----

class GaussPeak:
    """Config C3: Gaussian peak on a quadratic background, 6 parameters."""

    n = 6
    x0 = np.array([2.0, 0.0, 1.0, 0.5, 0.0, 0.0])
    lb = np.array([0.0, -2.0, 0.1, -1.0, -1.0, -1.0])
    ub = np.array([5.0, 2.0, 3.0, 2.0, 1.0, 1.0])

    def __init__(self, m=128):
        self.m = m
        self.t = np.linspace(-5.0, 5.0, m)

    def make_data(self, B, seed=0, noise=0.01):
        rng = np.random.default_rng(seed)
        A = rng.uniform(1.0, 3.0, B)
        mu = rng.uniform(-1.0, 1.0, B)
        sg = rng.uniform(0.5, 1.5, B)
        c0 = rng.uniform(0.0, 1.0, B)
        c1 = rng.uniform(-0.1, 0.1, B)
        c2 = rng.uniform(-0.01, 0.01, B)
        truth = np.stack([A, mu, sg, c0, c1, c2], axis=1)
        t = self.t
        z = (t - mu[:, None]) / sg[:, None]
        y = (A[:, None] * np.exp(-0.5 * z * z) + c0[:, None] +
             c1[:, None] * t + c2[:, None] * t * t)
        y = y + noise * rng.standard_normal(y.shape)
        return truth, y

    def fun_np(self, x, y):
        t = self.t
        z = (t - x[1]) / x[2]
        return x[0] * np.exp(-0.5 * z * z) + x[3] + x[4] * t + x[5] * t * t - y

    def fun_t(self, X, y):
        t = _dev_const(self, X)
        z = (t - X[:, 1:2]) / X[:, 2:3]
        return (X[:, 0:1] * torch.exp(-0.5 * z * z) + X[:, 3:4] +
                X[:, 4:5] * t + X[:, 5:6] * t * t - y)


# ---------------------------------------------- transcendental-free (FD) ----

class RatPoly5:
    """Rational model y = (a + b t + c t^2) / (1 + d t + e t^2), 5 parameters.

    Only + - * / : NumPy and torch (one rounding per eager op, same order)
    return bit-identical residuals, so a finite-difference Jacobian formed on
    the device can be compared with scipy's to the last bit.  This is the
    parity model for the 2-point / 3-point paths: with
    ExpDecay2 / GaussPeak the 1-ulp difference between libm's and CUDA's exp
    is amplified by 1/h ~ 7e7 and hides what the kernels themselves do."""

    n = 5
    x0 = np.array([1.0, 0.5, 0.2, 0.3, 0.1])
    lb = np.array([0.0, -1.0, 0.0, 0.0, 0.05])
    ub = np.array([3.0, 2.0, 0.25, 1.0, 1.0])

    def __init__(self, m=40):
        self.m = m
        self.t = np.linspace(0.0, 3.0, m)

    def make_data(self, B, seed=0, noise=0.01):
        rng = np.random.default_rng(seed)
        a = rng.uniform(0.5, 2.5, B)
        b = rng.uniform(-0.5, 1.5, B)
        c = rng.uniform(0.0, 0.5, B)        # > ub for about half the problems
        d = rng.uniform(0.1, 0.8, B)
        e = rng.uniform(0.0, 0.5, B)        # < lb for a tenth of them
        truth = np.stack([a, b, c, d, e], axis=1)
        y = np.stack([self.fun_np(truth[i], 0.0) for i in range(B)])
        y = y + noise * rng.standard_normal(y.shape)
        return truth, y

    def fun_np(self, x, y):
        t = self.t
        num = x[0] + x[1] * t + x[2] * t * t
        den = 1.0 + x[3] * t + x[4] * t * t
        return num / den - y

    def fun_t(self, X, y):
        t = _dev_const(self, X)
        num = X[:, 0:1] + X[:, 1:2] * t + X[:, 2:3] * t * t
        den = 1.0 + X[:, 3:4] * t + X[:, 4:5] * t * t
        return num / den - y


# ------------------------------------------------------------------ C4 ----

class TallLinExp:
    """Configs C4/C5: k linear columns + two exponentials, one tall problem.

    Rows are generated per shard from ``seed`` so a row-sharded run and a
    single-process run over the concatenated shards see identical data.
    """

    def __init__(self, m, n=64, seed=0, noise=0.01, dtype=np.float64,
                 x0_tail=(0.5, 1.0, 0.5, 1.0), lb=-0.5, ub=5.0):
        # the default start has the two exponentials
        # identical: J(x0) is exactly rank deficient and the symmetry is only
        # broken by rounding, so TRF iterates are ulp-chaotic from the ~5th
        # step on (the reference does not reproduce itself under 1-ulp noise
        # either); parity tests that gate x and nfev use an asymmetric tail
        assert n >= 5
        self.m, self.n, self.k = m, n, n - 4
        rng = np.random.default_rng(seed)
        self.A = rng.standard_normal((m, self.k))
        self.t = rng.uniform(0.0, 1.0, m)
        self.x_true = np.concatenate([rng.uniform(-1.0, 1.0, self.k),
                                      [1.0, 2.0, 0.5, 3.0]])
        self.y = self._model(self.x_true) + noise * rng.standard_normal(m)
        self.x0 = np.concatenate([np.full(self.k, 0.1), list(x0_tail)])
        # C5: lb = 0 puts about half of the linear truth values outside the box
        self.lb = np.full(n, float(lb))
        self.ub = np.full(n, float(ub))

    def _model(self, x):
        k, t = self.k, self.t
        return (self.A.dot(x[:k]) + x[k] * np.exp(-x[k + 1] * t) +
                x[k + 2] * np.exp(-x[k + 3] * t))

    def fun_np(self, x):
        return self._model(x) - self.y

    def jac_np(self, x):
        k, t = self.k, self.t
        J = np.empty((self.m, self.n))
        J[:, :k] = self.A
        e1 = np.exp(-x[k + 1] * t)
        e2 = np.exp(-x[k + 3] * t)
        J[:, k] = e1
        J[:, k + 1] = -x[k] * t * e1
        J[:, k + 2] = e2
        J[:, k + 3] = -x[k + 2] * t * e2
        return J

    # device form: data uploaded once, callbacks return local rows
    def to_device(self, device):
        self.A_t = torch.as_tensor(self.A, device=device)
        self.t_t = torch.as_tensor(self.t, device=device)
        self.y_t = torch.as_tensor(self.y, device=device)
        self.J_t = torch.empty((self.m, self.n), dtype=torch.float64,
                               device=device)
        self.J_t[:, :self.k] = self.A_t
        return self

    def fun_t(self, x):
        k, t = self.k, self.t_t
        return (self.A_t @ x[:k] + x[k] * torch.exp(-x[k + 1] * t) +
                x[k + 2] * torch.exp(-x[k + 3] * t) - self.y_t)

    def jac_t(self, x):
        k, t = self.k, self.t_t
        J = self.J_t
        e1 = torch.exp(-x[k + 1] * t)
        e2 = torch.exp(-x[k + 3] * t)
        J[:, k] = e1
        J[:, k + 1] = -x[k] * t * e1
        J[:, k + 2] = e2
        J[:, k + 3] = -x[k + 2] * t * e2
        return J


class TallLinExpDevice:
    """The C4/C5 workload generated ON the device (bench.py): same model family
    as TallLinExp, rows drawn with a torch generator so that 16M x 64 need not
    pass through host memory.  ``rows`` is this rank's share; ``seed`` should
    differ per rank."""

    def __init__(self, rows, n, device, seed=0, noise=0.01,
                 x0_tail=(0.5, 1.0, 0.5, 1.0), lb=-0.5, ub=5.0):
        assert n >= 6 and n % 2 == 0
        self.m, self.n, self.k = rows, n, n - 4
        k = self.k
        g = torch.Generator(device=device).manual_seed(1000 + seed)
        f64 = torch.float64
        self.A_t = torch.randn((rows, k), dtype=f64, device=device, generator=g)
        self.t_t = torch.rand(rows, dtype=f64, device=device, generator=g)
        rng = np.random.default_rng(12345)              # shared by all ranks
        self.x_true = np.concatenate([rng.uniform(-1.0, 1.0, k), [1.0, 2.0, 0.5, 3.0]])
        xt = torch.as_tensor(self.x_true, device=device)
        self.y_t = torch.zeros(rows, dtype=f64, device=device)
        self.y_t = self.fun_t(xt) + noise * torch.randn(rows, dtype=f64, device=device,
                                                        generator=g)
        self.x0 = np.concatenate([np.full(k, 0.1), list(x0_tail)])
        self.lb = np.full(n, float(lb))
        self.ub = np.full(n, float(ub))
        self.J_t = torch.empty((rows, n), dtype=f64, device=device)
        self.J_t[:, :k] = self.A_t

    def load(self, A, t, y):
        """Replace the data (e2e: fresh copies from pinned host buffers)."""
        self.A_t, self.t_t, self.y_t = A, t, y
        self.J_t[:, :self.k] = A

    fun_t = TallLinExp.fun_t
    jac_t = TallLinExp.jac_t
