"""CPU oracle for the bounded trust-region least-squares hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package imports this file;
it is loaded by ``tests/``, by ``__graft_entry__.smoke()`` as the checker and
by ``bench.py`` for the ``cpu_baseline`` / ``--impl reference`` legs.  The
product path (``bounded_lsq_b200``) has no CPU fallback and never routes here.

What it is: a NumPy/SciPy restatement of the algorithms in nmayorov/bounded-lsq
(``bounded_lsq/{bounds,trust_region,trf,dogbox,least_squares}.py``), organised
by the same stage boundaries the CUDA kernels use (linearise -> propose ->
judge) so that every kernel has a stage-level checker.  The floating-point
operations are issued through the same NumPy/SciPy entry points and in the same
order as the reference, so on one machine the oracle is bit-identical to the
reference; ``tests/golden/make_golden.py`` pins that claim by running the
unmodified reference from ``/root/reference`` and storing its outputs, and
``tests/test_oracle_golden.py`` checks the oracle against those files.

Parity status: PINNED (bit-for-bit against the reference's own ``trf`` /
``dogbox`` / helper functions on the committed golden vectors; see
DESIGN.md section "Oracle").

Third-party arithmetic that is not under /root/reference and is therefore
restated from its published behaviour (SciPy 1.18.1 / NumPy 2.3.5 here):
``scipy.optimize._numdiff.approx_derivative`` (2-point and 3-point schemes,
call site least_squares.py:359-361) -> :func:`fd_steps` / :func:`fd_jacobian`.
``scipy.linalg.svd`` (trf.py:272) and ``numpy.linalg.lstsq`` (dogbox.py:197)
are called directly.
"""
from __future__ import annotations

from math import copysign
from warnings import warn

import numpy as np
from numpy.linalg import lstsq, norm
from scipy.linalg import svd

EPS = np.finfo(float).eps
SQRT_EPS = EPS ** 0.5

MESSAGES = {
    0: "The maximum number of function evaluations is exceeded.",
    1: "`gtol` termination condition is satisfied.",
    2: "`ftol` termination condition is satisfied.",
    3: "`xtol` termination condition is satisfied.",
    4: "Both `ftol` and `xtol` termination conditions are satisfied.",
}


class Result(dict):
    """Attribute-style result record (same fields as trf.py:258-261)."""

    __getattr__ = dict.get
    __setattr__ = dict.__setitem__


# --------------------------------------------------------------------------
# Bound geometry  (bounds.py)
# --------------------------------------------------------------------------

def expand_bounds(bounds, x0):
    """bounds.py:7-16 -- scalar bounds become full vectors."""
    lo, hi = (np.asarray(b, dtype=float) for b in bounds)
    lo = np.resize(lo, x0.shape) if lo.ndim == 0 else lo
    hi = np.resize(hi, x0.shape) if hi.ndim == 0 else hi
    return lo, hi


def in_bounds(x, lb, ub):
    """bounds.py:19-21."""
    return bool(np.all((x >= lb) & (x <= ub)))


def step_size_to_bound(x, d, lb, ub):
    """bounds.py:24-48 -- smallest t>0 with x+t*d on the box; all ties hit."""
    nz = np.nonzero(d)
    dn = d[nz]
    t = np.full_like(x, np.inf)
    with np.errstate(over='ignore', invalid='ignore'):
        t[nz] = np.maximum((lb - x)[nz] / dn, (ub - x)[nz] / dn)
    tmin = np.min(t)
    return tmin, np.equal(t, tmin) * np.sign(d).astype(int)


def find_active_constraints(x, lb, ub, rtol=1e-12):
    """bounds.py:51-76."""
    out = np.zeros_like(x, dtype=int)
    below = x - lb
    above = ub - x
    nearer_lower = below < above
    hit = below[nearer_lower] < rtol * np.maximum(1, np.abs(lb[nearer_lower]))
    out[nearer_lower] = -hit.astype(int)
    hit = above[~nearer_lower] < rtol * np.maximum(1, np.abs(ub[~nearer_lower]))
    out[~nearer_lower] = hit.astype(int)
    return out


def make_strictly_feasible(x, lb, ub, rstep=0):
    """bounds.py:79-103 -- lower rule first, upper rule second."""
    y = x.copy()
    at_lo = x <= lb
    if rstep == 0:
        y[at_lo] = np.nextafter(lb[at_lo], ub[at_lo])
    else:
        y[at_lo] = lb[at_lo] + rstep * (1 + np.abs(lb[at_lo]))
    at_hi = x >= ub
    if rstep == 0:
        y[at_hi] = np.nextafter(ub[at_hi], lb[at_hi])
    else:
        y[at_hi] = ub[at_hi] - rstep * (1 + np.abs(ub[at_hi]))
    return y


def scaling_vector(x, g, lb, ub):
    """bounds.py:106-149 -- Coleman-Li v and dv/dx."""
    v = np.ones_like(x)
    jv = np.zeros_like(x)
    sel = (g < 0) & np.isfinite(ub)
    v[sel] = ub[sel] - x[sel]
    jv[sel] = -1
    sel = (g > 0) & np.isfinite(lb)
    v[sel] = x[sel] - lb[sel]
    jv[sel] = 1
    return v, jv


def cl_optimality(x, g, lb, ub):
    """bounds.py:152-156."""
    lb = np.resize(lb, x.shape)
    ub = np.resize(ub, x.shape)
    v, _ = scaling_vector(x, g, lb, ub)
    return norm(v * g, ord=np.inf)


def find_intersection(x, tr, lb, ub):
    """dogbox.py:9-35 -- box = bounds intersected with the rectangular region."""
    lo_c = lb - x
    hi_c = ub - x
    lo = np.maximum(lo_c, -tr)
    hi = np.minimum(hi_c, tr)
    return (lo, hi, np.equal(lo, lo_c), np.equal(hi, hi_c),
            np.equal(lo, -tr), np.equal(hi, tr))


# --------------------------------------------------------------------------
# Trust-region subproblem  (trust_region.py)
# --------------------------------------------------------------------------

def intersect_trust_region(x, s, Delta):
    """trust_region.py:11-44 -- roots of |x + t s| = Delta."""
    a = np.dot(s, s)
    if a == 0:
        raise ValueError("`s` is zero.")
    b = np.dot(x, s)
    c = np.dot(x, x) - Delta ** 2
    if c > 0:
        raise ValueError("`x` is not within the trust region.")
    disc = np.sqrt(b * b - a * c)
    q = -(b + copysign(disc, b))
    r1 = q / a
    r2 = c / q
    return (r1, r2) if r1 < r2 else (r2, r1)


def phi_and_derivative(alpha, suf, s, Delta):
    """trust_region.py:47-53."""
    den = s ** 2 + alpha
    pn = norm(suf / den)
    return pn - Delta, -np.sum(suf ** 2 / den ** 3) / pn


def solve_lsq_trust_region(n, m, uf, s, V, Delta, initial_alpha=None,
                           rtol=0.01, max_iter=10):
    """trust_region.py:56-152 -- More's iteration on phi(alpha) from an SVD."""
    suf = s * uf
    if m >= n:
        full_rank = s[-1] > EPS * m * s[0]
    else:
        full_rank = False

    if full_rank:
        p = -V.dot(uf / s)
        if norm(p) <= Delta:
            return p, 0.0, 0

    hi = norm(suf) / Delta
    if full_rank:
        phi, dphi = phi_and_derivative(0.0, suf, s, Delta)
        lo = -phi / dphi
    else:
        lo = 0.0

    if initial_alpha is None or not full_rank and initial_alpha == 0:
        alpha = max(0.001 * hi, (lo * hi) ** 0.5)
    else:
        alpha = initial_alpha

    for it in range(max_iter):
        if alpha < lo or alpha > hi:
            alpha = max(0.001 * hi, (lo * hi) ** 0.5)
        phi, dphi = phi_and_derivative(alpha, suf, s, Delta)
        if np.abs(phi) < rtol * Delta:
            break
        if phi < 0:
            hi = alpha
        q = phi / dphi
        lo = max(lo, alpha - q)
        alpha -= (phi + Delta) * q / Delta

    p = -V.dot(suf / (s ** 2 + alpha))
    if phi > 0:
        p *= Delta / norm(p)
    return p, alpha, it + 1


# --------------------------------------------------------------------------
# 1-D quadratic helpers  (trf.py:15-102)
# --------------------------------------------------------------------------

def minimize_quadratic(a, b, lo, hi):
    """trf.py:15-34 -- argmin of a t^2 + b t on [lo, hi], first minimum wins."""
    t = np.array([lo, hi])
    if a != 0:
        ext = -0.5 * b / a
        if lo <= ext <= hi:
            t = np.hstack((t, ext))
    y = a * t ** 2 + b * t
    k = np.argmin(y)
    return t[k], y[k]


def build_quadratic_1d(J, diag, g, s, s0=None):
    """trf.py:37-76."""
    v = J.dot(s)
    a = 0.5 * (np.dot(v, v) + np.dot(s * diag, s))
    b = np.dot(g, s)
    if s0 is not None:
        u = J.dot(s0)
        b += np.dot(u, v) + np.dot(s0 * diag, s)
    return a, b


def evaluate_quadratic(J, diag, g, steps):
    """trf.py:79-102 -- steps is (k, n)."""
    Js = J.dot(steps.T)
    return 0.5 * (np.sum(Js ** 2, axis=0) +
                  np.sum(diag * steps ** 2, axis=1)) + np.dot(steps, g)


def reflected_step(x, J_h, diag_h, g_h, p, p_h, d, Delta, lb, ub, theta):
    """trf.py:105-156.  Mutates p and p_h in place like the reference."""
    stride_p, hits = step_size_to_bound(x, p, lb, ub)
    r_h = np.copy(p_h)
    r_h[hits.astype(bool)] *= -1
    r = d * r_h
    p *= stride_p
    p_h *= stride_p
    x_edge = x + p
    _, to_tr = intersect_trust_region(p_h, r_h, Delta)
    to_bound, _ = step_size_to_bound(x_edge, r, lb, ub)
    to_bound *= theta
    hi = min(to_bound, to_tr)
    if hi > 0:
        lo = (1 - theta) * stride_p / hi
    else:
        lo = -1
    if lo <= hi:
        a, b = build_quadratic_1d(J_h, diag_h, g_h, r_h, s0=p_h)
        t, _ = minimize_quadratic(a, b, lo, hi)
        r_h = p_h + r_h * t
    else:
        r_h = None
    p_h *= theta
    return (p_h, p_h) if r_h is None else (p_h, r_h)


def gradient_step(x, J_h, diag_h, g_h, d, Delta, lb, ub, theta):
    """trf.py:159-170."""
    to_bound, _ = step_size_to_bound(x, -g_h * d, lb, ub)
    to_bound *= theta
    to_tr = Delta / norm(g_h)
    hi = min(to_bound, to_tr)
    a, b = build_quadratic_1d(J_h, diag_h, g_h, -g_h)
    t, _ = minimize_quadratic(a, b, 0.0, hi)
    return -t * g_h


def _is_jac_scaling(scaling):
    return isinstance(scaling, str) and scaling == 'jac'


# --------------------------------------------------------------------------
# Trust Region Reflective  (trf.py:173-358), staged
# --------------------------------------------------------------------------

class TRFStepper:
    """trf.py:173-358 cut at the kernel boundaries.

    start()      trf.py:201-235   strictly feasible x, f, J, scale, Delta0
    linearise()  trf.py:239-277   g, CL scaling, hat-space SVD, theta
    propose()    trf.py:284-308   TR solve, reflective/gradient candidates
    judge()      trf.py:313-344   ratio test, Delta/alpha update, ftol/xtol
    accept()     trf.py:346-352
    """

    def __init__(self, fun, jac, x0, lb, ub, ftol, xtol, gtol, max_nfev,
                 scaling, trace=None):
        self.fun, self.jac = fun, jac
        self.x0, self.lb, self.ub = x0, lb, ub
        self.ftol, self.xtol, self.gtol = ftol, xtol, gtol
        self.max_nfev = max_nfev
        self.scaling = scaling
        self.trace = trace
        self.status = None

    def start(self):
        self.x = make_strictly_feasible(self.x0, self.lb, self.ub, rstep=1e-10)
        self.f = self.fun(self.x)
        self.nfev = 1
        self.J = self.jac(self.x, self.f)
        self.njev = 1
        if self.f.shape[0] != self.J.shape[0]:
            raise RuntimeError("Inconsistent dimensions between the returns "
                               "of `fun` and `jac` on the first iteration.")
        g = self.J.T.dot(self.f)
        self.m, self.n = self.J.shape
        if _is_jac_scaling(self.scaling):
            cn = np.linalg.norm(self.J, axis=0)
            cn[cn == 0] = 1
            self.scale = 1 / cn
        else:
            self.scale = 1 / self.scaling
        v, _ = scaling_vector(self.x, g, self.lb, self.ub)
        self.Delta = norm(self.x0 / (self.scale * v ** 0.5))
        if self.Delta == 0:
            self.Delta = 1.0
        self.J_aug = np.empty((self.m + self.n, self.n))
        self.f_aug = np.zeros((self.m + self.n))
        self.obj = np.dot(self.f, self.f)
        self.alpha = 0.0
        if self.max_nfev is None:
            self.max_nfev = self.x0.size * 100
        self.g_norm = None

    def linearise(self):
        """Returns False when the solve ends here (gtol or a pending status)."""
        if _is_jac_scaling(self.scaling):
            cn = np.linalg.norm(self.J, axis=0)
            with np.errstate(divide='ignore'):
                self.scale = np.minimum(self.scale, 1 / cn)
        g = self.J.T.dot(self.f)
        v, jv = scaling_vector(self.x, g, self.lb, self.ub)
        self.d = v ** 0.5 * self.scale
        self.g_h = self.d * g
        self.diag_h = g * jv * self.scale ** 2
        self.g = g
        self.g_norm = norm(g * v, ord=np.inf)
        if self.g_norm < self.gtol:
            self.status = 1
        if self.status is not None:
            return False
        m = self.m
        self.J_h = self.J * self.d
        self.J_aug[:m] = self.J_h
        self.J_aug[m:] = np.diag(self.diag_h ** 0.5)
        self.f_aug[:m] = self.f
        U, s, Vt = svd(self.J_aug, full_matrices=False)
        self.s, self.V = s, Vt.T
        self.uf = U.T.dot(self.f_aug)
        self.theta = max(0.995, 1 - self.g_norm)
        return True

    def propose(self):
        p_h, self.alpha, n_it = solve_lsq_trust_region(
            self.n, self.m, self.uf, self.s, self.V, self.Delta,
            initial_alpha=self.alpha)
        p = self.d * p_h
        to_bound, _ = step_size_to_bound(self.x, p, self.lb, self.ub)
        if to_bound >= 1:
            p_h *= min(self.theta * to_bound, 1)
            cands = np.atleast_2d(p_h)
        else:
            p_h, r_h = reflected_step(self.x, self.J_h, self.diag_h, self.g_h,
                                      p, p_h, self.d, self.Delta, self.lb,
                                      self.ub, self.theta)
            c_h = gradient_step(self.x, self.J_h, self.diag_h, self.g_h,
                                self.d, self.Delta, self.lb, self.ub,
                                self.theta)
            cands = np.array([p_h, r_h, c_h])
        q = evaluate_quadratic(self.J_h, self.diag_h, self.g_h, cands)
        k = np.argmin(q)
        self.step_h = cands[k]
        self.predicted = -2 * q[k]
        self.step = self.d * self.step_h
        self.choice = int(k) if len(q) > 1 else -1
        self.x_new = make_strictly_feasible(self.x + self.step, self.lb,
                                            self.ub)
        return self.x_new

    def judge(self, f_new):
        """Returns True when a termination status was set (inner break)."""
        self.nfev += 1
        self.f_new = f_new
        self.obj_new = np.dot(f_new, f_new)
        self.actual = self.obj - self.obj_new
        corr = np.dot(self.step_h * self.diag_h, self.step_h)
        if self.predicted > 0:
            ratio = (self.actual - corr) / self.predicted
        else:
            ratio = 0
        Delta_used = self.Delta
        if ratio < 0.25:
            Dn = 0.25 * norm(self.step_h)
            self.alpha *= self.Delta / Dn
            self.Delta = Dn
        elif ratio > 0.75 and norm(self.step_h) > 0.95 * self.Delta:
            self.Delta *= 2.0
            self.alpha *= 0.5
        f_ok = abs(self.actual) < self.ftol * self.obj and ratio > 0.25
        x_ok = norm(self.step) < self.xtol * max(SQRT_EPS, norm(self.x))
        if f_ok and x_ok:
            self.status = 4
        elif f_ok:
            self.status = 2
        elif x_ok:
            self.status = 3
        if self.trace is not None:
            self.trace.append(dict(
                x=self.x.copy(), x_new=self.x_new.copy(),
                step=self.step.copy(), step_h=self.step_h.copy(),
                Delta=Delta_used, Delta_next=self.Delta, alpha=self.alpha,
                predicted=self.predicted, actual=self.actual, ratio=ratio,
                choice=self.choice, accepted=bool(self.actual > 0),
                status=self.status))
        return self.status is not None

    def accept(self):
        self.x = self.x_new
        self.f = self.f_new
        self.obj = self.obj_new
        self.J = self.jac(self.x, self.f)
        self.njev += 1

    def result(self, status):
        mask = find_active_constraints(self.x, self.lb, self.ub,
                                       rtol=self.xtol)
        return Result(x=self.x, fun=self.f, jac=self.J, obj_value=self.obj,
                      optimality=self.g_norm, active_mask=mask,
                      nfev=self.nfev, njev=self.njev, status=status,
                      x_covariance=None)

    def run(self):
        self.start()
        while self.nfev < self.max_nfev:
            if not self.linearise():
                return self.result(self.status)
            self.actual = -1
            while self.actual <= 0 and self.nfev < self.max_nfev:
                x_new = self.propose()
                if self.judge(self.fun(x_new)):
                    break
            if self.actual > 0:
                self.accept()
        return self.result(0)


def trf(fun, jac, x0, lb, ub, ftol, xtol, gtol, max_nfev, scaling,
        trace=None):
    """Drop-in for ``bounded_lsq.trf.trf`` (trf.py:173); ``jac(x, f)``."""
    return TRFStepper(fun, jac, x0, lb, ub, ftol, xtol, gtol, max_nfev,
                      scaling, trace).run()


# --------------------------------------------------------------------------
# dogbox  (dogbox.py)
# --------------------------------------------------------------------------

def _hit_bookkeeping(x, hits, o_l, o_u, t_l, t_u):
    marks = np.zeros_like(x, dtype=int)
    marks[(hits < 0) & o_l] = -1
    marks[(hits > 0) & o_u] = 1
    on_tr = np.any((hits < 0) & t_l | (hits > 0) & t_u)
    return marks, on_tr


def dogleg_step(x, cauchy, newton, tr, lb, ub):
    """dogbox.py:38-75."""
    lo, hi, o_l, o_u, t_l, t_u = find_intersection(x, tr, lb, ub)
    if in_bounds(newton, lo, hi):
        return newton, np.zeros_like(x, dtype=int), False
    if not in_bounds(cauchy, lo, hi):
        beta, _ = step_size_to_bound(np.zeros_like(cauchy), cauchy, lo, hi)
        cauchy = beta * cauchy
    diff = newton - cauchy
    t, hits = step_size_to_bound(cauchy, diff, lo, hi)
    marks, on_tr = _hit_bookkeeping(x, hits, o_l, o_u, t_l, t_u)
    return cauchy + t * diff, marks, on_tr


def constrained_cauchy_step(x, cauchy, tr, lb, ub):
    """dogbox.py:78-97."""
    lo, hi, o_l, o_u, t_l, t_u = find_intersection(x, tr, lb, ub)
    if in_bounds(cauchy, lo, hi):
        return cauchy, np.zeros_like(x, dtype=int), False
    beta, hits = step_size_to_bound(np.zeros_like(cauchy), cauchy, lo, hi)
    marks, on_tr = _hit_bookkeeping(x, hits, o_l, o_u, t_l, t_u)
    return beta * cauchy, marks, on_tr


class DogboxStepper:
    """dogbox.py:100-272 cut at the kernel boundaries."""

    def __init__(self, fun, jac, x0, lb, ub, ftol, xtol, gtol, max_nfev,
                 scaling, trace=None):
        self.fun, self.jac = fun, jac
        self.x0, self.lb, self.ub = x0, lb, ub
        self.ftol, self.xtol, self.gtol = ftol, xtol, gtol
        self.max_nfev = max_nfev
        self.scaling = scaling
        self.trace = trace
        self.status = None

    def start(self):
        x0 = self.x0
        self.f = self.fun(x0)
        self.nfev = 1
        self.J = self.jac(x0, self.f)
        self.njev = 1
        if self.f.shape[0] != self.J.shape[0]:
            raise RuntimeError("Inconsistent dimensions between the returns "
                               "of `fun` and `jac` on the first iteration.")
        if _is_jac_scaling(self.scaling):
            cn = np.linalg.norm(self.J, axis=0)
            cn[cn == 0] = 1
            self.scale = 1 / cn
        else:
            self.scale = 1 / self.scaling
        self.Delta = np.linalg.norm(x0 / self.scale, ord=np.inf)
        if self.Delta == 0:
            self.Delta = 1.0
        self.on_bound = np.zeros_like(x0, dtype=int)
        self.on_bound[np.equal(x0, self.lb)] = -1
        self.on_bound[np.equal(x0, self.ub)] = 1
        self.x = x0.copy()
        self.step = np.empty_like(x0)
        self.obj = np.dot(self.f, self.f)
        if self.max_nfev is None:
            self.max_nfev = x0.size * 100
        self.g_norm = None

    def linearise(self):
        if _is_jac_scaling(self.scaling):
            cn = np.linalg.norm(self.J, axis=0)
            with np.errstate(divide='ignore'):
                self.scale = np.minimum(self.scale, 1 / cn)
        g = self.J.T.dot(self.f)
        self.g = g
        active = self.on_bound * g < 0
        free = ~active
        self.free = free
        self.J_free = self.J[:, free]
        self.g_free = g[free]
        self.x_free = self.x[free]
        self.l_free = self.lb[free]
        self.u_free = self.ub[free]
        self.scale_free = self.scale[free]
        if np.all(active):
            self.g_norm = 0.0
            self.status = 1
        else:
            self.g_norm = norm(self.g_free, ord=np.inf)
            if self.g_norm < self.gtol:
                self.status = 1
        if self.status is not None:
            return False
        self.newton = lstsq(self.J_free, -self.f, rcond=None)[0]
        Jg = self.J_free.dot(self.g_free)
        with np.errstate(divide='ignore', invalid='ignore'):
            self.cauchy = (-np.dot(self.g_free, self.g_free) /
                           np.dot(Jg, Jg) * self.g_free)
        return True

    def propose(self):
        tr = self.Delta * self.scale_free
        sf, self.marks_free, self.tr_hit = dogleg_step(
            self.x_free, self.cauchy, self.newton, tr, self.l_free,
            self.u_free)
        Js = self.J_free.dot(sf)
        self.predicted = -np.dot(Js, Js) - 2 * np.dot(Js, self.f)
        self.fallback = False
        if self.predicted <= 0:
            sf, self.marks_free, self.tr_hit = constrained_cauchy_step(
                self.x_free, self.cauchy, tr, self.l_free, self.u_free)
            self.predicted = -np.dot(Js, Js) - 2 * np.dot(Js, self.f)
            self.fallback = True
        self.step.fill(0.0)
        self.step[self.free] = sf
        self.x_new = self.x + self.step
        return self.x_new

    def judge(self, f_new):
        self.nfev += 1
        self.f_new = f_new
        self.obj_new = np.dot(f_new, f_new)
        self.actual = self.obj - self.obj_new
        if self.predicted > 0:
            ratio = self.actual / self.predicted
        else:
            ratio = 0
        Delta_used = self.Delta
        if ratio < 0.25:
            self.Delta = 0.25 * norm(self.step / self.scale, ord=np.inf)
        elif ratio > 0.75 and self.tr_hit:
            self.Delta *= 2.0
        f_ok = abs(self.actual) < self.ftol * self.obj and ratio > 0.25
        x_ok = self.Delta < self.xtol * max(
            SQRT_EPS, norm(self.x / self.scale, ord=np.inf))
        if f_ok and x_ok:
            self.status = 4
        elif f_ok:
            self.status = 2
        elif x_ok:
            self.status = 3
        if self.trace is not None:
            self.trace.append(dict(
                x=self.x.copy(), x_new=self.x_new.copy(),
                step=self.step.copy(), Delta=Delta_used,
                Delta_next=self.Delta, predicted=self.predicted,
                actual=self.actual, ratio=ratio, fallback=self.fallback,
                tr_hit=bool(self.tr_hit), accepted=bool(self.actual > 0),
                status=self.status))
        return self.status is not None

    def accept(self):
        self.on_bound[self.free] = self.marks_free
        self.x = self.x_new
        sel = self.on_bound == -1
        self.x[sel] = self.lb[sel]
        sel = self.on_bound == 1
        self.x[sel] = self.ub[sel]
        self.f = self.f_new
        self.obj = self.obj_new
        self.J = self.jac(self.x, self.f)
        self.njev += 1

    def result(self, status):
        return Result(x=self.x, fun=self.f, jac=self.J, obj_value=self.obj,
                      optimality=self.g_norm, active_mask=self.on_bound,
                      nfev=self.nfev, njev=self.njev, status=status,
                      x_covariance=None)

    def run(self):
        self.start()
        while self.nfev < self.max_nfev:
            if not self.linearise():
                return self.result(self.status)
            self.actual = -1.0
            while self.actual <= 0 and self.nfev < self.max_nfev:
                x_new = self.propose()
                if self.judge(self.fun(x_new)):
                    break
            if self.actual > 0:
                self.accept()
        return self.result(0)


def dogbox(fun, jac, x0, lb, ub, ftol, xtol, gtol, max_nfev, scaling,
           trace=None):
    """Drop-in for ``bounded_lsq.dogbox.dogbox`` (dogbox.py:100)."""
    return DogboxStepper(fun, jac, x0, lb, ub, ftol, xtol, gtol, max_nfev,
                         scaling, trace).run()


# --------------------------------------------------------------------------
# Finite-difference Jacobian (call site least_squares.py:357-365; arithmetic
# restated from scipy.optimize._numdiff, SciPy 1.18.1)
# --------------------------------------------------------------------------

def fd_steps(x, lb, ub, diff_step=None, method='2-point'):
    """Absolute steps h and the one-sided flags, after the bound adjustment.

    _compute_absolute_step + _adjust_scheme_to_bounds of SciPy 1.18.1.
    """
    sgn = (x >= 0).astype(float) * 2 - 1
    rel = SQRT_EPS if method == '2-point' else EPS ** (1 / 3)
    h_default = rel * sgn * np.maximum(1.0, np.abs(x))
    if diff_step is None:
        h = h_default
    else:
        h = diff_step * sgn * np.abs(x)
        dx = (x + h) - x
        h = np.where(dx == 0, h_default, h)

    if method == '2-point':
        one_sided = np.ones_like(h, dtype=bool)
        if np.all((lb == -np.inf) & (ub == np.inf)):
            return h, one_sided
        h_adj = h.copy()
        below = x - lb
        above = ub - x
        xp = x + h
        violated = (xp < lb) | (xp > ub)
        fitting = np.abs(h) <= np.maximum(below, above)
        h_adj[violated & fitting] *= -1
        fwd = (above >= below) & ~fitting
        h_adj[fwd] = above[fwd] / 1
        bwd = (above < below) & ~fitting
        h_adj[bwd] = -below[bwd] / 1
        return h_adj, one_sided

    # '3-point': central where it fits, else one-sided with 2 steps
    h = np.abs(h)
    one_sided = np.zeros_like(h, dtype=bool)
    if np.all((lb == -np.inf) & (ub == np.inf)):
        return h, one_sided
    h_adj = h.copy()
    below = x - lb
    above = ub - x
    central = (below >= h) & (above >= h)
    fwd = (above >= below) & ~central
    h_adj[fwd] = np.minimum(h[fwd], 0.5 * above[fwd] / 1)
    one_sided[fwd] = True
    bwd = (above < below) & ~central
    h_adj[bwd] = -np.minimum(h[bwd], 0.5 * below[bwd] / 1)
    one_sided[bwd] = True
    min_dist = np.minimum(above, below) / 1
    adj_central = ~central & (np.abs(h_adj) <= min_dist)
    h_adj[adj_central] = min_dist[adj_central]
    one_sided[adj_central] = False
    return h_adj, one_sided


def fd_jacobian(fun, x, f0, lb, ub, diff_step=None, method='2-point'):
    """Dense forward/central difference Jacobian (_dense_difference)."""
    if np.any((x < lb) | (x > ub)):
        raise ValueError("`x0` violates bound constraints.")
    f0 = np.atleast_1d(f0)
    h, one_sided = fd_steps(x, lb, ub, diff_step, method)
    n, m = x.size, f0.size
    Jt = np.empty((n, m))
    for i in range(n):
        if method == '2-point':
            x1 = np.copy(x)
            x1[i] = x[i] + h[i]
            dx = (x[i] + h[i]) - x[i]
            df = np.atleast_1d(fun(x1)) - f0
        else:
            x1 = np.copy(x)
            x2 = np.copy(x)
            if one_sided[i]:
                x1[i] = x[i] + h[i]
                x2[i] = x[i] + 2 * h[i]
                dx = x2[i] - x[i]
                f1 = np.atleast_1d(fun(x1))
                f2 = np.atleast_1d(fun(x2))
                df = -3.0 * f0 + 4 * f1 - f2
            else:
                x1[i] = x[i] - h[i]
                x2[i] = x[i] + h[i]
                dx = x2[i] - x1[i]
                f1 = np.atleast_1d(fun(x1))
                f2 = np.atleast_1d(fun(x2))
                df = f2 - f1
        Jt[i] = df / dx
    if m == 1:
        Jt = np.ravel(Jt)
    return Jt.T


# --------------------------------------------------------------------------
# Front end  (least_squares.py:120-383, trf/dogbox branches only)
# --------------------------------------------------------------------------

def check_tolerances(ftol, xtol, gtol):
    """least_squares.py:15-27."""
    msg = "{} is too low, setting to machine epsilon {}."
    out = []
    for name, tol in (("`ftol`", ftol), ("`xtol`", xtol), ("`gtol`", gtol)):
        if tol < EPS:
            warn(msg.format(name, EPS))
            tol = EPS
        out.append(tol)
    return tuple(out)


def check_scaling(scaling, x0):
    """least_squares.py:100-117."""
    if _is_jac_scaling(scaling):
        return scaling
    try:
        scaling = np.asarray(scaling, dtype=float)
    except ValueError:
        raise ValueError("`scaling` must be 'jac' or array-like with numbers.")
    if np.any(scaling <= 0):
        raise ValueError("`scaling` must contain only positive values.")
    if scaling.ndim == 0:
        scaling = np.resize(scaling, x0.shape)
    if scaling.shape != x0.shape:
        raise ValueError("Inconsistent shapes between `scaling` and `x0`.")
    return scaling


def least_squares(fun, x0, jac='2-point', bounds=(-np.inf, np.inf),
                  method='trf', ftol=SQRT_EPS, xtol=SQRT_EPS, gtol=SQRT_EPS,
                  max_nfev=None, scaling=1.0, diff_step=None, args=(),
                  kwargs={}, options={}, trace=None):
    """least_squares.py:120 for method in {'trf','dogbox'}."""
    if method not in ('trf', 'dogbox'):
        raise ValueError("`method` must be 'trf' or 'dogbox' "
                         "('lm' is outside the oracle's scope).")
    if len(bounds) != 2:
        raise ValueError("`bounds` must contain 2 elements.")
    x0 = np.atleast_1d(x0).astype(float)
    if x0.ndim > 1:
        raise ValueError("`x0` must have at most 1 dimension.")
    lb, ub = expand_bounds(bounds, x0)
    if lb.shape != x0.shape or ub.shape != x0.shape:
        raise ValueError("Inconsistent shapes between bounds and `x0`.")
    if np.any(lb >= ub):
        raise ValueError("Each lower bound mush be strictly less than each "
                         "upper bound.")
    if jac not in ('2-point', '3-point') and not callable(jac):
        raise ValueError("`jac` must be '2-point', '3-point' or callable.")
    scaling = check_scaling(scaling, x0)
    ftol, xtol, gtol = check_tolerances(ftol, xtol, gtol)
    if not in_bounds(x0, lb, ub):
        raise ValueError("`x0` is infeasible.")

    def fun_w(x):
        f = np.atleast_1d(fun(x, *args, **kwargs))
        if f.ndim > 1:
            raise RuntimeError("`fun` must return at most 1-d array_like.")
        return f

    if callable(jac):
        def jac_w(x, f):
            J = np.atleast_2d(jac(x, *args, **kwargs))
            if J.ndim > 2:
                raise RuntimeError("`jac` must return at most 2-d "
                                   "array_like.")
            return J
    else:
        def jac_w(x, f):
            J = fd_jacobian(lambda z: fun(z, *args, **kwargs), x, f, lb, ub,
                            diff_step, jac)
            return np.atleast_2d(J)

    solver = trf if method == 'trf' else dogbox
    res = solver(fun_w, jac_w, x0, lb, ub, ftol, xtol, gtol, max_nfev,
                 scaling, trace=trace, **options)
    res.message = MESSAGES[res.status]
    res.success = res.status > 0
    return res
