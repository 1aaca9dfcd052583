"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's Bezier
constraint path (oracle for the CUDA kernels).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``optimalbeziertrajectorygeneration_b200``) never does.

Every function cites the reference lines it restates (paths are relative to
the upstream repository root).  The restatement keeps the reference's quirks
(SURVEY.md section 0): Q1 (normSquare = dim/2 * |c|^2), Q3 (diff keeps the
degree), Q8 (point obstacles as constant curves, all pairs), Q14 (tables from
scipy.special.binom).

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this module
against ``tests/golden/*.npz`` which ``oracle/make_golden.py`` produced by
running the unmodified reference in the authoring container.

All functions accept a ``dtype`` so the same code can run in ``np.longdouble``
to produce an (almost) exactly rounded finite-difference quotient, which is
what Jacobian parity is judged against (hard part 1).
"""
from fractions import Fraction

import numpy as np
from scipy.special import binom


def _cast(a, dtype):
    """array -> dtype; dtype=object means exact rationals (fractions.Fraction of
    the fp64 values), used for the exactly rounded finite-difference quotient."""
    a = np.asarray(a)
    if dtype is object:
        if a.dtype == object:
            return a
        out = np.empty(a.shape, dtype=object)
        flat = out.reshape(-1)
        for i, v in enumerate(a.reshape(-1)):
            flat[i] = Fraction(float(v))
        return out
    return a.astype(dtype)


def _num(v, dtype):
    return Fraction(v) if dtype is object else dtype(v)

# --------------------------------------------------------------------------
# constant tables (bezier.py:1127-1147, 1151-1176, 1183-1208)
# --------------------------------------------------------------------------


def elev_matrix(N, R):
    """bezier.py:1127-1147  T[j,i] = C(N,j) C(R,i-j) / C(N+R,i)."""
    T = np.zeros((N + 1, N + R + 1))
    for i in range(N + R + 1):
        den = binom(N + R, i)
        for j in range(N + 1):
            T[j, i] = binom(N, j) * binom(R, i - j) / den
    return T


def prod_weights(m, n=None):
    """Weights of the Bernstein product, W[i,j] = C(m,i) C(n,j) / C(m+n,i+j).

    bezier.py:1151-1176 (prodMatrix, m == n) and bezier.py:1183-1208
    (bezProductCoefficients).  Both tables hold exactly one such weight per
    (i,j); they are stored here as the dense (m+1, n+1) array of those weights
    computed with the very same ``binom`` expression so values are bit-equal.
    """
    if n is None:
        n = m
    W = np.zeros((m + 1, n + 1))
    for k in range(m + n + 1):
        den = binom(m + n, k)
        for i in range(max(0, k - n), min(m, k) + 1):
            W[i, k - i] = binom(m, i) * binom(n, k - i) / den
    return W


# --------------------------------------------------------------------------
# Bezier algebra on raw control-point arrays [dim, n+1]
# --------------------------------------------------------------------------


def elev(cpts, R, dtype=np.float64):
    """Bezier.elev, bezier.py:469-495 (np.dot(cpts_row, elevMat))."""
    cpts = np.atleast_2d(_cast(cpts, dtype))
    N = cpts.shape[1] - 1
    T = _cast(elev_matrix(N, R), dtype)
    out = np.zeros((cpts.shape[0], N + R + 1), dtype=dtype)
    for j in range(N + 1):          # ascending-j accumulation
        out += cpts[:, j:j + 1] * T[j:j + 1, :]
    return out


def mul(a, b, dtype=np.float64):
    """Bezier.mul / multiplyBezCurves for 1-D rows (bezier.py:376-432,
    1211-1246).  General (correct) product; the reference is only right for
    equal degrees (Q2), which is all the hot path uses."""
    a = _cast(a, dtype).ravel()
    b = _cast(b, dtype).ravel()
    m, n = a.size - 1, b.size - 1
    W = _cast(prod_weights(m, n), dtype)
    out = np.zeros(m + n + 1, dtype=dtype)
    for i in range(m + 1):
        for j in range(n + 1):
            out[i + j] += (a[i] * b[j]) * W[i, j]
    return out


def norm_square(cpts, dtype=np.float64):
    """Bezier.normSquare -> _normSquare (bezier.py:869-889, 1724-1756).

    Returns the 2n+1 Bernstein coefficients of (dim/2) * sum_d c_d(t)^2 (Q1).
    """
    cpts = np.atleast_2d(_cast(cpts, dtype))
    dim, n1 = cpts.shape
    n = n1 - 1
    W = _cast(prod_weights(n), dtype)
    G = np.zeros((n1, n1), dtype=dtype)           # x.T @ x  (bezier.py:1746)
    for d in range(dim):
        G += np.outer(cpts[d], cpts[d])
    out = np.zeros(2 * n + 1, dtype=dtype)
    for i in range(n1):                           # prodM gather (bezier.py:1748)
        for j in range(n1):
            out[i + j] += W[i, j] * G[i, j]
    # S sums `dim` identical rows, then /2  (bezier.py:1750-1756, 884)
    return (out * _num(dim, dtype)) / _num(2, dtype)


def diff(cpts, T, dtype=np.float64):
    """Bezier.diff (bezier.py:497-519, 1100-1123): derivative control points
    n/T * (P[i+1]-P[i]) followed by one degree elevation (same degree n)."""
    cpts = np.atleast_2d(_cast(cpts, dtype))
    n = cpts.shape[1] - 1
    val = _num(n, dtype) / _num(T, dtype)
    # np.dot(cpts, Dm): column i = -val*P[i] + val*P[i+1]
    d = cpts[:, :-1] * (-val) + cpts[:, 1:] * val
    return elev(d, 1, dtype=dtype)


def de_casteljau_split(c, t, dtype=np.float64):
    """deCasteljauSplit (bezier.py:985-1027) for one 1-D row at local
    parameter t in [0,1]; returns (left, right) with right already flipped to
    ascending order as Bezier.split does (bezier.py:563)."""
    c = np.asarray(c, dtype=dtype).copy()
    n1 = c.size
    left = np.zeros(n1, dtype=dtype)
    right = np.zeros(n1, dtype=dtype)
    t = dtype(t)
    for lvl in range(n1 - 1):
        left[lvl] = c[0]
        right[lvl] = c[n1 - 1 - lvl]
        for i in range(n1 - 1 - lvl):
            c[i] = (1 - t) * c[i] + t * c[i + 1]
    left[n1 - 1] = right[n1 - 1] = c[0]
    return left, right[::-1].copy()


def split(cpts, tDiv, t0=0.0, tf=1.0, dtype=np.float64):
    """Bezier.split (bezier.py:533-572)."""
    cpts = np.atleast_2d(_cast(cpts, dtype))
    if np.isnan(tDiv):
        tDiv = 0
    t = (dtype(tDiv) - dtype(t0)) / (dtype(tf) - dtype(t0))
    L = np.empty_like(cpts)
    Rr = np.empty_like(cpts)
    for d in range(cpts.shape[0]):
        L[d], Rr[d] = de_casteljau_split(cpts[d], t, dtype=dtype)
    return L, Rr


def de_casteljau_eval(cpts, tau, t0=0.0, tf=1.0):
    """Bezier.__call__ / deCasteljauCurve (bezier.py:187-203, 944-982)."""
    cpts = np.atleast_2d(np.asarray(cpts, dtype=np.float64))
    tau = np.atleast_1d(np.asarray(tau, dtype=np.float64))
    T = (tau - t0) / (tf - t0)
    out = np.empty((cpts.shape[0], T.size))
    for d in range(cpts.shape[0]):
        for k, t in enumerate(T):
            c = cpts[d].copy()
            for lvl in range(c.size - 1):
                c[:c.size - 1 - lvl] = (1 - t) * c[:c.size - 1 - lvl] + t * c[1:c.size - lvl]
            out[d, k] = c[0]
    return out


def bez_extreme(cpts_row, tol=1e-6, maximum=False, max_depth=64):
    """Intended algorithm of Bezier.min / Bezier.max (bezier.py:631-667,
    727-763): split at the extreme control point (local parameter idx/deg)
    until an end point is extreme or the bound moves by < tol.

    The reference passes idx/deg as an *absolute* time to split(), which is
    only the same thing on the root curve with t0=0, tf=1 (Q4); beyond depth 1
    the reference extrapolates and is not a usable oracle.  This restatement
    is what the kernels implement; tests pin it against the reference only on
    inputs whose recursion depth is <= 1."""
    sgn = -1.0 if maximum else 1.0

    def rec(c, glob, depth):
        idx = int(np.argmin(sgn * c))
        new = c[idx]
        if abs(glob - new) < tol or depth >= max_depth:
            return new
        if idx != 0 and idx != c.size - 1:
            l, r = de_casteljau_split(c, idx / (c.size - 1))
            a = rec(l, new, depth + 1)
            b = rec(r, new, depth + 1)
            new = max(a, b) if maximum else min(a, b)
        return new

    c = np.asarray(cpts_row, dtype=np.float64)
    return rec(c, np.inf if maximum else -np.inf, 0)


# --------------------------------------------------------------------------
# problem assembly (optimization.py)
# --------------------------------------------------------------------------


class Model:
    """The fields of BezOptimization.model that the path reads
    (optimization.py:21-63)."""

    def __init__(self, numVeh=1, dimension=1, degree=5, minimizeGoal='Euclidean',
                 maxSep=0.9, minSpeed=0, maxSpeed=1e6, maxAngRate=1e6,
                 initPoints=None, finalPoints=None, initSpeeds=None,
                 finalSpeeds=None, initAngs=None, finalAngs=None, tf=1.0,
                 pointObstacles=None, shapeObstacles=None):
        self.numVeh, self.dim, self.deg = numVeh, dimension, degree
        self.minGoal = minimizeGoal
        self.maxSep, self.minSpeed, self.maxSpeed = maxSep, minSpeed, maxSpeed
        self.maxAngRate = maxAngRate
        self.has_pts = initPoints is not None
        self.has_spd = initSpeeds is not None
        self.initPoints = np.atleast_2d(initPoints)
        self.finalPoints = np.atleast_2d(finalPoints)
        self.initSpeeds = np.atleast_1d(initSpeeds)
        self.finalSpeeds = np.atleast_1d(finalSpeeds)
        self.initAngs = np.atleast_1d(initAngs)
        self.finalAngs = np.atleast_1d(finalAngs)
        self.tf = tf
        self.pointObstacles = pointObstacles
        self.shapeObstacles = shapeObstacles
        self.numCols = degree + 1 - (2 if self.has_pts else 0) - (2 if self.has_spd else 0)

    @property
    def timeopt(self):
        return self.minGoal.lower() == 'timeopt'

    @property
    def nvar(self):
        return self.numVeh * self.dim * self.numCols + (1 if self.timeopt else 0)


def reshape_vector(model, x, dtype=np.float64):
    """BezOptimization.reshapeVector (optimization.py:242-285)."""
    x = _cast(x, dtype)
    dim, deg, numVeh = model.dim, model.deg, model.numVeh
    tf = _num(model.tf, dtype)
    if model.timeopt:
        tf = x[-1]
        x = x[:-1]
    y = np.empty((dim * numVeh, deg + 1), dtype=dtype)
    offset = 0
    if model.has_pts:
        offset += 1
        for i in range(model.initPoints.shape[0]):
            y[i * dim:(i + 1) * dim, 0] = _cast(model.initPoints[i], dtype)
            y[i * dim:(i + 1) * dim, -1] = _cast(model.finalPoints[i], dtype)
    if model.has_spd:
        offset += 1
        initMag = _cast(model.initSpeeds.astype(np.float64), dtype) * tf / deg
        finalMag = _cast(model.finalSpeeds.astype(np.float64), dtype) * tf / deg
        ia = model.initAngs.astype(np.float64)
        fa = model.finalAngs.astype(np.float64)
        ip, fp = _cast(model.initPoints, dtype), _cast(model.finalPoints, dtype)
        y[::2, 1] = ip[:, 0] + initMag * _cast(np.cos(ia), dtype)
        y[1::2, 1] = ip[:, 1] + initMag * _cast(np.sin(ia), dtype)
        y[::2, -2] = fp[:, 0] - finalMag * _cast(np.cos(fa), dtype)
        y[1::2, -2] = fp[:, 1] - finalMag * _cast(np.sin(fa), dtype)
    y[:, offset:deg + 1 - offset] = x.reshape((dim * numVeh, model.numCols))
    return y


def model_tf(model, x):
    """optimization.py:140-143 / 158-161 / 176-179."""
    return x[-1] if model.timeopt else model.tf


def stack_obstacles(model, y, dtype=np.float64):
    """optimization.py:86-94: point obstacles appended as constant curves."""
    if model.pointObstacles is None:
        return y, model.numVeh
    rows = []
    for obstacle in model.pointObstacles:
        for d in range(model.dim):
            rows.append([obstacle[d]] * (model.deg + 1))
    return np.vstack((y, _cast(np.asarray(rows, dtype=np.float64), dtype))), model.numVeh + len(model.pointObstacles)


def temporal_separation(y, nVeh, dim, maxSep, elev_R, dtype=np.float64):
    """_temporalSeparationConstraints (optimization.py:311-346): for every
    pair i<j (lexicographic) elev(normSquare(c_i - c_j), E) - maxSep^2."""
    if nVeh <= 1:
        return None
    y = _cast(y, dtype)
    out = []
    for i in range(nVeh - 1):
        for j in range(i + 1, nVeh):
            dv = y[i * dim:(i + 1) * dim] - y[j * dim:(j + 1) * dim]
            out.append(elev(norm_square(dv, dtype=dtype), elev_R, dtype=dtype)[0])
    return np.concatenate(out) - _num(maxSep, dtype) ** 2


def speed_sq(y, nVeh, dim, tf, elev_R, dtype=np.float64):
    """Shared body of _min/_maxSpeedConstraints (optimization.py:373-382,
    411-420): elev(normSquare(diff(c_i)), E) for every vehicle."""
    y = _cast(y, dtype)
    out = []
    for i in range(nVeh):
        v = diff(y[i * dim:(i + 1) * dim], tf, dtype=dtype)
        out.append(elev(norm_square(v, dtype=dtype), elev_R, dtype=dtype)[0])
    return np.concatenate(out)


def max_speed(y, nVeh, dim, tf, maxSpeed, elev_R, dtype=np.float64):
    """_maxSpeedConstraints (optimization.py:387-422)."""
    return _num(maxSpeed, dtype) ** 2 - speed_sq(y, nVeh, dim, tf, elev_R, dtype)


def min_speed(y, nVeh, dim, tf, minSpeed, elev_R, dtype=np.float64):
    """_minSpeedConstraints (optimization.py:349-384)."""
    return speed_sq(y, nVeh, dim, tf, elev_R, dtype) - _num(minSpeed, dtype) ** 2


def angular_rate_sq(cpts2d, tf, dtype=np.float64):
    """_angularRateSqr (optimization.py:578-611) on a 2-D curve of degree m:
    control-point-wise ratio of two degree-4m Bernstein polynomials (Q11)."""
    cpts2d = _cast(cpts2d, dtype)
    if cpts2d.shape[0] != 2:
        raise ValueError('The input curve must be two dimensional,\n'
                         'instead it is {} dimensional'.format(cpts2d.shape[0]))
    xD = diff(cpts2d[0], tf, dtype)[0]
    xDD = diff(xD, tf, dtype)[0]
    yD = diff(cpts2d[1], tf, dtype)[0]
    yDD = diff(yD, tf, dtype)[0]
    num = mul(yDD, xD, dtype) - mul(xDD, yD, dtype)
    num = mul(num, num, dtype)
    den = mul(xD, xD, dtype) + mul(yD, yD, dtype)
    den = mul(den, den, dtype)
    return num / den


def max_angular_rate(y, nVeh, dim, tf, maxAngRate, elev_R, dtype=np.float64):
    """_maxAngularRateConstraints (optimization.py:425-459)."""
    y = _cast(y, dtype)
    out = []
    for i in range(nVeh):
        pos = elev(y[i * dim:(i + 1) * dim], elev_R, dtype=dtype)
        out.append(angular_rate_sq(pos, tf, dtype=dtype))
    return _num(maxAngRate, dtype) ** 2 - np.concatenate(out)


# closures with the reference's signatures -----------------------------------


def make_callables(model, elev_R, dtype=np.float64):
    """The closures BezOptimization hands to SciPy (optimization.py:83-187),
    with DEG_ELEV bound to ``elev_R``."""

    def sep(x):
        y = reshape_vector(model, x, dtype)
        y, nObs = stack_obstacles(model, y, dtype)
        return temporal_separation(y, nObs, model.dim, model.maxSep, elev_R, dtype)

    def maxspd(x):
        return max_speed(reshape_vector(model, x, dtype), model.numVeh, model.dim,
                         model_tf(model, _cast(x, dtype)), model.maxSpeed, elev_R, dtype)

    def minspd(x):
        return min_speed(reshape_vector(model, x, dtype), model.numVeh, model.dim,
                         model_tf(model, _cast(x, dtype)), model.minSpeed, elev_R, dtype)

    def angrate(x):
        return max_angular_rate(reshape_vector(model, x, dtype), model.numVeh, model.dim,
                                model_tf(model, _cast(x, dtype)), model.maxAngRate,
                                elev_R, dtype)

    return {'sep': sep, 'maxspeed': maxspd, 'minspeed': minspd, 'angrate': angrate}


# --------------------------------------------------------------------------
# objectives (optimization.py:462-519)
# --------------------------------------------------------------------------


def euclidean_objective(y, nVeh, dim):
    """_euclideanObjective (optimization.py:462-489) for dim == 3 (for dim 2
    the reference reads an uninitialised third component, Q10; the
    restatement treats it as 0)."""
    y = np.asarray(y, dtype=np.float64)
    s = 0.0
    for veh in range(nVeh):
        for i in range(y.shape[1] - 1):
            acc = 0.0
            for j in range(dim):
                t = y[veh * dim + j, i + 1] - y[veh * dim + j, i]
                acc += t * t
            s += np.sqrt(acc)
    return s


def accel_objective(y, nVeh, dim, tf, elev_R):
    """_minAccelObjective (optimization.py:503-519)."""
    s = 0.0
    for i in range(nVeh):
        a = diff(diff(y[i * dim:(i + 1) * dim], tf), tf)
        s = s + elev(norm_square(a), elev_R).sum()
    return s


# --------------------------------------------------------------------------
# finite-difference Jacobian exactly as SciPy's SLSQP does it
#   scipy/optimize/_slsqp_py.py:349-367 -> approx_derivative(fun, x,
#   method='2-point', abs_step=1.4901161193847656e-08)
#   scipy/optimize/_numdiff.py:585-596 (step), 683-712 (dense difference)
# --------------------------------------------------------------------------

SLSQP_EPS = 1.4901161193847656e-08


def fd_steps(x0, abs_step=SLSQP_EPS):
    """h and the divisor dx=(x0+h)-x0 SciPy uses for every variable.
    _numdiff.py:585-596: h = abs_step, falling back to the relative step
    sqrt(eps)*sign(x0)*max(1,|x0|) where x0 + abs_step == x0."""
    x0 = np.asarray(x0, dtype=np.float64)
    h = np.full_like(x0, abs_step)
    dx = (x0 + h) - x0
    sign_x0 = (x0 >= 0).astype(float) * 2 - 1
    rel = np.finfo(np.float64).eps ** 0.5
    h = np.where(dx == 0, rel * sign_x0 * np.maximum(1.0, np.abs(x0)), h)
    dx = (x0 + h) - x0
    return h, dx


def fd_jacobian(fun, x0, abs_step=SLSQP_EPS):
    """Dense 2-point forward Jacobian, _numdiff.py:683-712."""
    x0 = np.asarray(x0, dtype=np.float64)
    f0 = np.atleast_1d(fun(x0))
    h, dx = fd_steps(x0, abs_step)
    J = np.empty((f0.size, x0.size))
    for k in range(x0.size):
        x1 = x0.copy()
        x1[k] = x0[k] + h[k]
        J[:, k] = (np.atleast_1d(fun(x1)) - f0) / dx[k]
    return J


def fd_jacobian_exact(fun_x, x0, abs_step=SLSQP_EPS, dtype=object):
    """The same quotient with f evaluated in exact rational arithmetic
    (dtype=object: fractions.Fraction; the fp64 tables and inputs are taken as
    exact rationals) on the *fp64* perturbed points, i.e. the exactly rounded
    value of SciPy's formula.  dtype=np.longdouble is a cheaper variant that is
    only good to ~|f| * 1e-19 / h.  ``fun_x`` comes from make_callables(...,
    dtype=dtype)."""
    x0 = np.asarray(x0, dtype=np.float64)
    f0 = np.atleast_1d(fun_x(_cast(x0, dtype)))
    h, dx = fd_steps(x0, abs_step)
    J = np.empty((f0.size, x0.size))
    for k in range(x0.size):
        x1 = x0.copy()
        x1[k] = x0[k] + h[k]
        f1 = np.atleast_1d(fun_x(_cast(x1, dtype)))
        q = (f1 - f0) / _num(float(dx[k]), dtype)
        J[:, k] = np.array([float(v) for v in q]) if dtype is object else q.astype(np.float64)
    return J
