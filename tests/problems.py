"""Small classical least-squares test problems (More, Garbow, Hillstrom 1981).

Written from the published formulae; used as the config-#1 style parity corpus
(BASELINE.json configs[0]).  Starting points and bound sets follow the
instances the reference benchmark suite runs (benchmarks/lsq_problems.py,
listed by `extract_lsq_problems`, lsq_problems.py:1003-1018), so results are
comparable with its published tables.  Every problem is a `Problem` with
`fun(x) -> (m,)`, `jac(x) -> (m, n)`, `x0`, `lb`, `ub`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

INF = np.inf


@dataclass
class Problem:
    name: str
    fun: Callable
    jac: Callable
    x0: np.ndarray
    lb: np.ndarray
    ub: np.ndarray

    @property
    def n(self):
        return self.x0.size


def _mk(name, fun, jac, x0, lb=None, ub=None):
    x0 = np.asarray(x0, dtype=float)
    n = x0.size
    lb = np.full(n, -INF) if lb is None else np.asarray(lb, dtype=float)
    ub = np.full(n, INF) if ub is None else np.asarray(ub, dtype=float)
    return Problem(name, fun, jac, x0, lb, ub)


# --- Rosenbrock -----------------------------------------------------------
def rosen_f(x):
    return np.array([10.0 * (x[1] - x[0] ** 2), 1.0 - x[0]])


def rosen_j(x):
    return np.array([[-20.0 * x[0], 10.0], [-1.0, 0.0]])


# --- Freudenstein and Roth --------------------------------------------------
def freud_f(x):
    return np.array([-13.0 + x[0] + ((5.0 - x[1]) * x[1] - 2.0) * x[1],
                     -29.0 + x[0] + ((x[1] + 1.0) * x[1] - 14.0) * x[1]])


def freud_j(x):
    return np.array([[1.0, 10.0 * x[1] - 3.0 * x[1] ** 2 - 2.0],
                     [1.0, 3.0 * x[1] ** 2 + 2.0 * x[1] - 14.0]])


# --- Powell badly scaled ----------------------------------------------------
def powbad_f(x):
    return np.array([1e4 * x[0] * x[1] - 1.0,
                     np.exp(-x[0]) + np.exp(-x[1]) - 1.0001])


def powbad_j(x):
    return np.array([[1e4 * x[1], 1e4 * x[0]],
                     [-np.exp(-x[0]), -np.exp(-x[1])]])


# --- Brown badly scaled -----------------------------------------------------
def brownbad_f(x):
    return np.array([x[0] - 1e6, x[1] - 2e-6, x[0] * x[1] - 2.0])


def brownbad_j(x):
    return np.array([[1.0, 0.0], [0.0, 1.0], [x[1], x[0]]])


# --- Beale ------------------------------------------------------------------
_BEALE_Y = np.array([1.5, 2.25, 2.625])
_BEALE_K = np.array([1.0, 2.0, 3.0])


def beale_f(x):
    return _BEALE_Y - x[0] * (1.0 - x[1] ** _BEALE_K)


def beale_j(x):
    J = np.empty((3, 2))
    J[:, 0] = -(1.0 - x[1] ** _BEALE_K)
    J[:, 1] = x[0] * _BEALE_K * x[1] ** (_BEALE_K - 1.0)
    return J


# --- Jennrich and Sampson (m = 10) -----------------------------------------
_JS_I = np.arange(1.0, 11.0)


def jensam_f(x):
    # model - data, the sign the reference's JenrichAndSampson uses
    return np.exp(_JS_I * x[0]) + np.exp(_JS_I * x[1]) - (2.0 + 2.0 * _JS_I)


def jensam_j(x):
    J = np.empty((10, 2))
    J[:, 0] = _JS_I * np.exp(_JS_I * x[0])
    J[:, 1] = _JS_I * np.exp(_JS_I * x[1])
    return J


# --- Helical valley ---------------------------------------------------------
def helix_f(x):
    # theta = arctan(x2 / x1) / (2 pi), + 1/2 on the half plane x1 <= 0 (the
    # branch the reference's HelicalValley takes, lsq_problems.py:546-553)
    theta = np.arctan(x[1] / x[0]) / (2.0 * np.pi) + (0.5 if x[0] <= 0 else 0.0)
    return np.array([10.0 * (x[2] - 10.0 * theta),
                     10.0 * (np.sqrt(x[0] ** 2 + x[1] ** 2) - 1.0),
                     x[2]])


def helix_j(x):
    r2 = x[0] ** 2 + x[1] ** 2
    r = np.sqrt(r2)
    c = 100.0 / (2.0 * np.pi)
    return np.array([[c * x[1] / r2, -c * x[0] / r2, 10.0],
                     [10.0 * x[0] / r, 10.0 * x[1] / r, 0.0],
                     [0.0, 0.0, 1.0]])


# --- Box three-dimensional (m = 10) -----------------------------------------
_BOX_T = 0.1 * np.arange(1.0, 11.0)


def box3_f(x):
    t = _BOX_T
    return (np.exp(-t * x[0]) - np.exp(-t * x[1]) -
            x[2] * (np.exp(-t) - np.exp(-10.0 * t)))


def box3_j(x):
    t = _BOX_T
    J = np.empty((10, 3))
    J[:, 0] = -t * np.exp(-t * x[0])
    J[:, 1] = t * np.exp(-t * x[1])
    J[:, 2] = -(np.exp(-t) - np.exp(-10.0 * t))
    return J


# --- Powell singular --------------------------------------------------------
_R5, _R10 = np.sqrt(5.0), np.sqrt(10.0)


def powsing_f(x):
    return np.array([x[0] + 10.0 * x[1], _R5 * (x[2] - x[3]),
                     (x[1] - 2.0 * x[2]) ** 2, _R10 * (x[0] - x[3]) ** 2])


def powsing_j(x):
    a = 2.0 * (x[1] - 2.0 * x[2])
    b = 2.0 * _R10 * (x[0] - x[3])
    return np.array([[1.0, 10.0, 0.0, 0.0],
                     [0.0, 0.0, _R5, -_R5],
                     [0.0, a, -2.0 * a, 0.0],
                     [b, 0.0, 0.0, -b]])


# --- Wood -------------------------------------------------------------------
_R90 = np.sqrt(90.0)


def wood_f(x):
    return np.array([10.0 * (x[1] - x[0] ** 2), 1.0 - x[0],
                     _R90 * (x[3] - x[2] ** 2), 1.0 - x[2],
                     _R10 * (x[1] + x[3] - 2.0), (x[1] - x[3]) / _R10])


def wood_j(x):
    return np.array([[-20.0 * x[0], 10.0, 0.0, 0.0],
                     [-1.0, 0.0, 0.0, 0.0],
                     [0.0, 0.0, -2.0 * _R90 * x[2], _R90],
                     [0.0, 0.0, -1.0, 0.0],
                     [0.0, _R10, 0.0, _R10],
                     [0.0, 1.0 / _R10, 0.0, -1.0 / _R10]])


# --- Kowalik and Osborne (m = 11) -------------------------------------------
_KO_Y = np.array([0.1957, 0.1947, 0.1735, 0.1600, 0.0844, 0.0627, 0.0456,
                  0.0342, 0.0323, 0.0235, 0.0246])
_KO_U = np.array([4.0, 2.0, 1.0, 0.5, 0.25, 0.167, 0.125, 0.1, 0.0833,
                  0.0714, 0.0625])


def kowosb_f(x):
    # model - data, the sign the reference's EnzymeReaction uses
    u = _KO_U
    return x[0] * (u * u + u * x[1]) / (u * u + u * x[2] + x[3]) - _KO_Y


def kowosb_j(x):
    u = _KO_U
    num = u * u + u * x[1]
    den = u * u + u * x[2] + x[3]
    J = np.empty((11, 4))
    J[:, 0] = num / den
    J[:, 1] = x[0] * u / den
    J[:, 2] = -x[0] * num * u / den ** 2
    J[:, 3] = -x[0] * num / den ** 2
    return J


# --- Brown and Dennis (m = 20) ----------------------------------------------
_BD_T = np.arange(1.0, 21.0) / 5.0


def brownden_f(x):
    t = _BD_T
    return ((x[0] + t * x[1] - np.exp(t)) ** 2 +
            (x[2] + x[3] * np.sin(t) - np.cos(t)) ** 2)


def brownden_j(x):
    t = _BD_T
    a = 2.0 * (x[0] + t * x[1] - np.exp(t))
    b = 2.0 * (x[2] + x[3] * np.sin(t) - np.cos(t))
    J = np.empty((20, 4))
    J[:, 0] = a
    J[:, 1] = a * t
    J[:, 2] = b
    J[:, 3] = b * np.sin(t)
    return J


# --- Biggs EXP6 (m = 13) ------------------------------------------------------
_BG_T = 0.1 * np.arange(1.0, 14.0)
_BG_Y = np.exp(-_BG_T) - 5.0 * np.exp(-10.0 * _BG_T) + 3.0 * np.exp(-4.0 * _BG_T)


def biggs_f(x):
    t = _BG_T
    return (x[2] * np.exp(-t * x[0]) - x[3] * np.exp(-t * x[1]) +
            x[5] * np.exp(-t * x[4]) - _BG_Y)


def biggs_j(x):
    t = _BG_T
    e0, e1, e4 = np.exp(-t * x[0]), np.exp(-t * x[1]), np.exp(-t * x[4])
    J = np.empty((13, 6))
    J[:, 0] = -t * x[2] * e0
    J[:, 1] = t * x[3] * e1
    J[:, 2] = e0
    J[:, 3] = -e1
    J[:, 4] = -t * x[5] * e4
    J[:, 5] = e4
    return J


# --- Watson (m = 31) ----------------------------------------------------------
_WT = np.arange(1.0, 30.0) / 29.0


def _watson(n):
    jj = np.arange(n)
    P = _WT[:, None] ** jj                      # t^(j-1), j = 1..n
    Q = np.zeros((29, n))
    Q[:, 1:] = jj[1:] * _WT[:, None] ** (jj[1:] - 1)   # (j-1) t^(j-2)

    def f(x):
        s = P.dot(x)
        out = np.empty(31)
        out[:29] = Q.dot(x) - s * s - 1.0
        out[29] = x[0]
        out[30] = x[1] - x[0] ** 2 - 1.0
        return out

    def j(x):
        s = P.dot(x)
        J = np.zeros((31, n))
        J[:29] = Q - 2.0 * s[:, None] * P
        J[29, 0] = 1.0
        J[30, 0] = -2.0 * x[0]
        J[30, 1] = 1.0
        return J

    return f, j


# --- Penalty I (n = 10) -------------------------------------------------------
_PA = np.sqrt(1e-5)


def pen1_f(x):
    return np.concatenate([_PA * (x - 1.0), [np.dot(x, x) - 0.25]])


def pen1_j(x):
    n = x.size
    J = np.zeros((n + 1, n))
    J[:n] = _PA * np.eye(n)
    J[n] = 2.0 * x
    return J


# --- Trigonometric (n = 10) ---------------------------------------------------
def trig_f(x):
    n = x.size
    i = np.arange(1.0, n + 1.0)
    return n - np.sum(np.cos(x)) + i * (1.0 - np.cos(x)) - np.sin(x)


def trig_j(x):
    n = x.size
    i = np.arange(1.0, n + 1.0)
    J = np.tile(np.sin(x), (n, 1))
    J[np.diag_indices(n)] += i * np.sin(x) - np.cos(x)
    return J


# --- Meyer (thermistor, m = 16) ----------------------------------------------
# The reference's ThermistorResistance instance samples t = 5 + 45 i
# (lsq_problems.py:311); MGH / MINPACK-2 print t = 45 + 5 i.  Config #1 is the
# reference's suite, so its abscissae are used here.
_MY_T = 5.0 + 45.0 * np.arange(1.0, 17.0)
_MY_Y = np.array([34780., 28610., 23650., 19630., 16370., 13720., 11540.,
                  9744., 8261., 7030., 6005., 5147., 4427., 3820., 3307.,
                  2872.])


def meyer_f(x):
    return x[0] * np.exp(x[1] / (_MY_T + x[2])) - _MY_Y


def meyer_j(x):
    e = np.exp(x[1] / (_MY_T + x[2]))
    J = np.empty((16, 3))
    J[:, 0] = e
    J[:, 1] = x[0] * e / (_MY_T + x[2])
    J[:, 2] = -x[0] * x[1] * e / (_MY_T + x[2]) ** 2
    return J


# --- Gaussian (m = 15) --------------------------------------------------------
_GS_T = (8.0 - np.arange(1.0, 16.0)) / 2.0
_GS_Y = np.array([0.0009, 0.0044, 0.0175, 0.0540, 0.1295, 0.2420, 0.3521,
                  0.3989, 0.3521, 0.2420, 0.1295, 0.0540, 0.0175, 0.0044,
                  0.0009])


def gauss_f(x):
    return x[0] * np.exp(-x[1] * (_GS_T - x[2]) ** 2 / 2.0) - _GS_Y


def gauss_j(x):
    q = (_GS_T - x[2])
    e = np.exp(-x[1] * q ** 2 / 2.0)
    J = np.empty((15, 3))
    J[:, 0] = e
    J[:, 1] = -x[0] * e * q ** 2 / 2.0
    J[:, 2] = x[0] * x[1] * e * q
    return J


# --- Chebyshev quadrature (MGH 35; m = n) -------------------------------------
def _chebyquad(n):
    """f_i = (1/n) sum_j T_i(x_j) - int_0^1 T_i, T_i the Chebyshev polynomial
    shifted to [0, 1]; the integral is 0 for odd i and -1/(i^2 - 1) for even i.
    Three-term recurrences in t = 2 x - 1 for T_i and T_i'."""
    ii = np.arange(1, n + 1)
    integ = np.array([-1.0 / (i * i - 1.0) if i % 2 == 0 else 0.0 for i in ii])

    def tables(x):
        t = 2.0 * x - 1.0
        T = np.empty((n + 1, n))
        D = np.empty((n + 1, n))
        T[0], D[0] = 1.0, 0.0
        T[1], D[1] = t, 1.0
        for k in range(1, n):
            T[k + 1] = 2.0 * t * T[k] - T[k - 1]
            D[k + 1] = 2.0 * T[k] + 2.0 * t * D[k] - D[k - 1]
        return T[1:], D[1:]

    def f(x):
        T, _ = tables(x)
        return T.mean(axis=1) - integ

    def j(x):
        _, D = tables(x)
        return 2.0 * D / n

    return f, j


def _cheby_x0(n):
    return np.arange(1.0, n + 1.0) / (n + 1.0)


# --- Gulf research and development (m = 100) ----------------------------------
_GU_T = np.arange(1.0, 101.0) / 100.0
_GU_Y = 25.0 + (-50.0 * np.log(_GU_T)) ** (2.0 / 3.0)


def gulf_f(x):
    return np.exp(-np.abs(x[1] - _GU_Y) ** x[2] / x[0]) - _GU_T


def gulf_j(x):
    d = x[1] - _GU_Y
    ad = np.abs(d)
    pw = ad ** x[2]
    e = np.exp(-pw / x[0])
    J = np.empty((100, 3))
    J[:, 0] = pw / x[0] ** 2 * e
    J[:, 1] = -np.sign(d) * x[2] * ad ** (x[2] - 1.0) / x[0] * e
    with np.errstate(divide="ignore", invalid="ignore"):
        J[:, 2] = np.nan_to_num(-np.log(ad) * pw / x[0] * e)
    return J


# --- Penalty II (m = 2 n) ------------------------------------------------------
def _penalty2(n):
    y = np.exp(0.1 * np.arange(1.0, n) + 0.1) + np.exp(0.1 * np.arange(1.0, n))
    w = np.arange(n, 0, -1.0)                       # n - j + 1

    def f(x):
        out = np.empty(2 * n)
        out[0] = x[0]
        out[1:n] = _PA * (np.exp(0.1 * x[1:]) + np.exp(0.1 * x[:-1]) - y)
        out[n:2 * n - 1] = _PA * (np.exp(0.1 * x[1:]) - np.exp(-0.1))
        out[2 * n - 1] = np.sum(w * x * x) - 1.0
        return out

    def j(x):
        J = np.zeros((2 * n, n))
        J[0, 0] = 1.0
        e = _PA * 0.1 * np.exp(0.1 * x)
        r = np.arange(1, n)
        J[r, r] = e[1:]
        J[r, r - 1] = e[:-1]
        J[n - 1 + r, r] = e[1:]
        J[2 * n - 1] = 2.0 * w * x
        return J

    return f, j


# --- data sets of the MINPACK-2 collection (Averick et al. 1992): Osborne's
#     exponential and Gaussian fitting data, the coating thickness measurements.
#     The tables live in tests/golden/problem_data.npz (written by
#     tests/golden/make_golden.py from the reference's benchmark module).
_DATA = None


def _data():
    global _DATA
    if _DATA is None:
        import os
        _DATA = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                     "golden", "problem_data.npz"))
    return _DATA


# --- Osborne 1: exponential fitting (n = 5, m = 33) ---------------------------
def _osborne1():
    y = _data()["osborne1_y"]
    t = 10.0 * np.arange(33.0)

    def f(x):
        return x[0] + x[1] * np.exp(-x[3] * t) + x[2] * np.exp(-x[4] * t) - y

    def j(x):
        e3, e4 = np.exp(-x[3] * t), np.exp(-x[4] * t)
        return np.stack([np.ones(33), e3, e4, -x[1] * t * e3, -x[2] * t * e4], axis=1)

    return f, j


# --- Osborne 2: Gaussian fitting (n = 11, m = 65) -----------------------------
def _osborne2():
    y = _data()["osborne2_y"]
    t = 0.1 * np.arange(65.0)

    def parts(x):
        q = [t - x[8], t - x[9], t - x[10]]
        e = [np.exp(-x[4] * t)] + [np.exp(-x[5 + k] * q[k] ** 2) for k in range(3)]
        return q, e

    def f(x):
        _, e = parts(x)
        return x[0] * e[0] + x[1] * e[1] + x[2] * e[2] + x[3] * e[3] - y

    def j(x):
        q, e = parts(x)
        cols = [e[0], e[1], e[2], e[3], -x[0] * t * e[0]]
        cols += [-x[1 + k] * q[k] ** 2 * e[1 + k] for k in range(3)]
        cols += [2.0 * x[1 + k] * x[5 + k] * q[k] * e[1 + k] for k in range(3)]
        return np.stack(cols, axis=1)

    return f, j


# --- coating thickness standardisation (n = 134, m = 252) ---------------------
def _coating():
    xi = _data()["coating_xi"]          # (2, 63)
    y = _data()["coating_y"]            # (126,)
    q = 63
    s1, s2 = 4.08, 0.417

    def shifted(x):
        return xi[0] + x[8:8 + q], xi[1] + x[8 + q:]

    def f(x):
        a, b = shifted(x)
        z1 = x[0] + x[1] * a + x[2] * b + x[3] * a * b
        z2 = x[4] + x[5] * a + x[6] * b + x[7] * a * b
        return np.concatenate([z1 - y[:q], z2 - y[q:], s1 * x[8:8 + q], s2 * x[8 + q:]])

    def j(x):
        a, b = shifted(x)
        J = np.zeros((4 * q, 8 + 2 * q))
        r = np.arange(q)
        for blk, o in ((0, 0), (1, 4)):
            rows = r + blk * q
            J[rows, o] = 1.0
            J[rows, o + 1] = a
            J[rows, o + 2] = b
            J[rows, o + 3] = a * b
            J[rows, 8 + r] = x[o + 1] + x[o + 3] * b
            J[rows, 8 + q + r] = x[o + 2] + x[o + 3] * a
        J[2 * q + r, 8 + r] = s1
        J[3 * q + r, 8 + q + r] = s2
        return J

    return f, j


def corpus(max_n=None):
    """The parity corpus = config #1: the 32 unbounded + 26 bounded instances
    the reference's benchmark runs (lsq_problems.py:1003-1018), under the
    reference's names.  max_n drops the instances with more parameters."""
    P = []
    a = P.append
    wat = {n: _watson(n) for n in (6, 9, 12, 20)}
    cheb = {n: _chebyquad(n) for n in (7, 8, 9, 10, 11)}
    pen2 = {n: _penalty2(n) for n in (4, 10)}
    o1f, o1j = _osborne1()
    o2f, o2j = _osborne2()
    ctf, ctj = _coating()
    # ---- unbounded (32) ----
    a(_mk("Beale", beale_f, beale_j, [1.0, 1.0]))
    a(_mk("Biggs", biggs_f, biggs_j, [1.0, 2.0, 1.0, 1.0, 1.0, 1.0]))
    a(_mk("Box3D", box3_f, box3_j, [0.0, 10.0, 20.0]))
    a(_mk("BrownAndDennis", brownden_f, brownden_j, [25.0, 5.0, -5.0, -1.0]))
    a(_mk("BrownBadlyScaled", brownbad_f, brownbad_j, [1.0, 1.0]))
    for n in (7, 8, 9, 10, 11):
        a(_mk(f"ChebyshevQuadrature{n}", *cheb[n], _cheby_x0(n)))
    a(_mk("CoatingThickness", ctf, ctj,
          np.concatenate([[-8.0, 13.0, 1.2, 0.2, 0.1, 6.0, 5.5, -5.2], np.zeros(126)])))
    a(_mk("EnzymeReaction", kowosb_f, kowosb_j,
          np.array([2.5, 3.9, 4.15, 3.9]) * 1e-1))
    a(_mk("ExponentialFitting", o1f, o1j, [0.5, 1.5, -1.0, 0.01, 0.02]))
    a(_mk("ExtendedPowellSingular", powsing_f, powsing_j, [3.0, -1.0, 0.0, 1.0]))
    a(_mk("FreudensteinAndRoth", freud_f, freud_j, [-0.5, 2.0]))
    a(_mk("GaussianFittingI", o2f, o2j,
          [1.3, 0.65, 0.65, 0.7, 0.6, 3.0, 5.0, 7.0, 2.0, 4.5, 5.5]))
    a(_mk("GaussianFittingII", gauss_f, gauss_j, [0.4, 1.0, 0.0]))
    a(_mk("GulfRnD", gulf_f, gulf_j, [5.0, 2.5, 0.15]))
    a(_mk("HelicalValley", helix_f, helix_j, [-1.0, 0.0, 0.0]))
    a(_mk("JenrichAndSampson10", jensam_f, jensam_j, [0.3, 0.4]))
    a(_mk("PenaltyI", pen1_f, pen1_j, np.arange(1.0, 11.0)))
    a(_mk("PenaltyII10", *pen2[10], np.full(10, 0.5)))
    a(_mk("PenaltyII4", *pen2[4], np.full(4, 0.5)))
    a(_mk("PowellBadlyScaled", powbad_f, powbad_j, [0.0, 1.0]))
    a(_mk("Rosenbrock", rosen_f, rosen_j, [-2.0, 1.0]))
    a(_mk("ThermistorResistance", meyer_f, meyer_j, [0.02, 4000.0, 250.0]))
    a(_mk("Trigonometric", trig_f, trig_j, np.full(10, 0.1)))
    for n in (12, 20, 6, 9):
        a(_mk(f"Watson{n}", *wat[n], np.zeros(n)))
    a(_mk("Wood", wood_f, wood_j, [-3.0, -1.0, -3.0, -1.0]))
    # ---- bounded (26): the bound sets of the reference suite's *_B instances ----
    a(_mk("Beale_B", beale_f, beale_j, [1.0, 1.0], [0.6, 0.5], [10.0, 100.0]))
    a(_mk("Biggs_B", biggs_f, biggs_j, [1.0, 2.0, 1.0, 1.0, 1.0, 1.0],
          [0.0, 0.0, 0.0, 1.0, 0.0, 0.0], [2.0, 8.0, 1.0, 7.0, 5.0, 5.0]))
    a(_mk("Box3D_B", box3_f, box3_j, [0.0, 7.5, 20.0], [0.0, 5.0, 0.0],
          [2.0, 9.5, 20.0]))
    a(_mk("BrownAndDennis_B", brownden_f, brownden_j, [25.0, 5.0, -5.0, -1.0],
          [-10.0, 0.0, -100.0, -20.0], [100.0, 15.0, 0.0, 0.2]))
    a(_mk("BrownBadlyScaled_B", brownbad_f, brownbad_j, [1.0, 1.0],
          [0.0, 3e-5], [1e6, 100.0]))
    a(_mk("ChebyshevQuadrature10_B", *cheb[10], _cheby_x0(10),
          [0.0, 0.1, 0.2, 0.0, 0.0, 0.5, 0.5, 0.5, 0.5, 0.5],
          [1.0, 0.2, 0.3, 0.4, 0.5, 1.0, 1.0, 1.0, 1.0, 1.0]))
    x7 = _cheby_x0(7)
    x7[:3] = [0.025, 0.1, 0.15]
    a(_mk("ChebyshevQuadrature7_B", *cheb[7], x7, np.zeros(7),
          [0.05, 0.23, 0.333, 1.0, 1.0, 1.0, 1.0]))
    x8 = _cheby_x0(8)
    x8[:3] = [0.02, 0.1, 0.2]
    a(_mk("ChebyshevQuadrature8_B", *cheb[8], x8,
          [0.0, 0.0, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0],
          [0.04, 0.2, 0.3, 1.0, 1.0, 1.0, 1.0, 1.0]))
    a(_mk("ExtendedPowellSingular_B", powsing_f, powsing_j, [3.0, -1.0, 0.0, 1.0],
          [0.1, -20.0, -1.0, -1.0], [100.0, 20.0, 1.0, 50.0]))
    a(_mk("GaussianFittingII_B", gauss_f, gauss_j, [0.4, 1.0, 0.0],
          [0.398, 1.0, -0.5], [4.2, 2.0, 0.1]))
    a(_mk("GulfRnD_B", gulf_f, gulf_j, [5.0, 2.5, 0.15], np.zeros(3), np.full(3, 10.0)))
    a(_mk("HelicalValley_B", helix_f, helix_j, [-1.0, 0.0, 0.0],
          [-100.0, -1.0, -1.0], [0.8, 1.0, 1.0]))
    a(_mk("PenaltyI_B", pen1_f, pen1_j, np.arange(1.0, 11.0),
          [0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0], [100.0] * 10))
    a(_mk("PenaltyII10_B", *pen2[10], np.full(10, 0.5),
          [-10.0, 0.1, 0.0, 0.05, 0.0, -10.0, 0.0, 0.2, 0.0, 0.0],
          [50.0] * 9 + [0.5]))
    a(_mk("PenaltyII4_B", *pen2[4], np.full(4, 0.5),
          [-10.0, 0.3, 0.0, -1.0], [50.0, 50.0, 50.0, 0.5]))
    a(_mk("PowellBadlyScaled_B", powbad_f, powbad_j, [0.0, 1.0], [0.0, 1.0],
          [1.0, 9.0]))
    a(_mk("Rosenbrock_B_0", rosen_f, rosen_j, [-2.0, 1.0], [-INF, -1.5], None))
    a(_mk("Rosenbrock_B_1", rosen_f, rosen_j, [2.0, 2.0], [-INF, 1.5], None))
    a(_mk("Rosenbrock_B_2", rosen_f, rosen_j, [-2.0, 2.0], [-INF, 1.5], None))
    a(_mk("Rosenbrock_B_3", rosen_f, rosen_j, [0.0, 2.0], [-INF, 1.5],
          [1.0, INF]))
    a(_mk("Rosenbrock_B_4", rosen_f, rosen_j, [2.0, 2.0], [1.0, 1.5],
          [3.0, 3.0]))
    a(_mk("Rosenbrock_B_5", rosen_f, rosen_j, [-1.2, 1.0], [-50.0, 0.0],
          [0.5, 100.0]))
    a(_mk("Trigonometric_B", trig_f, trig_j, 5.0 + 10.0 * np.arange(10),
          10.0 * np.arange(10), 10.0 * np.arange(10) + 10.0))
    a(_mk("Watson12_B", *wat[12], np.zeros(12),
          [-1.0, 0.0, -1.0, -1.0, -1.0, 0.0, -3.0, 0.0, -10.0, 0.0, -5.0, 0.0],
          [0.0, 0.9, 0.0, 0.3, 0.0, 1.0, 0.0, 10.0, 0.0, 10.0, 0.0, 1.0]))
    a(_mk("Watson9_B", *wat[9], np.zeros(9),
          [-1e-5, 0.0, 0.0, 0.0, 0.0, -3.0, 0.0, -3.0, 0.0],
          [1e-5, 0.9, 0.1, 1.0, 1.0, 0.0, 4.0, 0.0, 2.0]))
    a(_mk("Wood_B", wood_f, wood_j, [-3.0, -1.0, -3.0, -1.0],
          [-100.0] * 4, [0.0, 10.0, 100.0, 100.0]))
    if max_n is not None:
        P = [p for p in P if p.n <= max_n]
    return P
