"""Drop-in mirror of the reference's ``bezier.py`` (class ``Bezier``) whose
arithmetic and distance routines run on the B200 through libbezgpu.so.

Same constructor, methods, properties and error behaviour as the reference
(bezier.py:34-889): ``Bezier(cpts=None, t0=0.0, tf=1.0, tau=None)``; ``elev``,
``diff``, ``mul``/``*``, ``add``/``+``, ``sub``/``-``, ``div``/``/``,
``normSquare``, ``split``, ``min``, ``max``, ``minDist``, ``minDist2Poly``,
``collCheck``, ``collCheck2Poly``, ``integrate``, ``copy``, ``__call__``,
``cpts, deg, degree, dim, dimension, t0, tf, tau, curve, x, y, z``.

Differences, all documented in DESIGN.md:
  * ``mul`` is the correct product for unequal degrees too (the reference is
    only right for equal degrees);
  * ``min``/``max`` implement the intended subdivision (the reference
    extrapolates beyond depth 1 and may not terminate);
  * ``minDist`` & co. work at all (they raise at the reference's HEAD, SURVEY
    Q5) and report a depth-limit status instead of RecursionError (Q6);
  * plotting is not part of the hot path and is not provided.

Each call moves a few hundred bytes to the device and back; the batched,
device-resident path for optimisation loops is ``optimization.BezOptimization``.
There is no CPU fallback: without the CUDA library these methods raise.
"""
import ctypes

import numpy as np
import torch

from . import _capi, _tables
from .engine import F64, _ptr, _require_cuda, _stream

_dev_tables = {}


def _device():
    _require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _table(kind, *key):
    """Device copy of a host table, cached per (kind, parameters, device)."""
    dev = _device()
    k = (kind,) + key + (dev.index,)
    t = _dev_tables.get(k)
    if t is None:
        if kind == "elev":
            host = _tables.elev_matrix(*key)
        elif kind == "prod":
            host = _tables.prod_weights(*key)
        else:
            raise KeyError(kind)
        t = torch.as_tensor(np.array(host, dtype=np.float64, copy=True), device=dev)
        _dev_tables[k] = t
    return t


def _to_dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=_device())


def _iptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class BezierParams:
    """Parameter storage of bezier.py:34-146 (control points, time window, tau)."""

    def __init__(self, cpts=None, tau=None, t0=0.0, tf=1.0):
        self._curve = None
        if cpts is not None:
            if not isinstance(cpts, np.ndarray):        # the reference needs .ndim (Q5b)
                cpts = np.array(cpts, dtype=float)
            if cpts.ndim == 1:
                self._cpts = np.atleast_2d(cpts)
                self._dim = 1
                self._deg = cpts.size - 1
            else:
                self._cpts = cpts
                self._dim = self._cpts.shape[0]
                self._deg = self._cpts.shape[1] - 1
        else:
            self._dim = None
            self._deg = None
        if tau is not None:
            self._t0 = tau[0]
            self._tf = tau[-1]
        else:
            self._t0 = float(t0)
            self._tf = float(tf)
        self._tau = tau

    @property
    def cpts(self):
        return self._cpts

    @cpts.setter
    def cpts(self, value):
        self._curve = None
        if isinstance(value, np.ndarray) and value.ndim == 2 and value.dtype == 'float64':
            newCpts = value
        else:
            newCpts = np.array(value, ndmin=2, dtype=float)
        self._dim = newCpts.shape[0]
        self._deg = newCpts.shape[1] - 1
        self._cpts = newCpts

    @property
    def deg(self):
        return self._deg

    @property
    def degree(self):
        return self._deg

    @property
    def dim(self):
        return self._dim

    @property
    def dimension(self):
        return self._dim

    @property
    def t0(self):
        return self._t0

    @t0.setter
    def t0(self, value):
        self._t0 = float(value)
        self._tau = None

    @property
    def tf(self):
        return self._tf

    @tf.setter
    def tf(self, value):
        self._tf = float(value)
        self._tau = None

    @property
    def tau(self):
        if self._tau is None:
            self._tau = np.linspace(self._t0, self._tf, 1001)
        elif not isinstance(self._tau, np.ndarray):
            self._tau = np.array(self._tau)
        return self._tau

    @tau.setter
    def tau(self, val):
        self._curve = None
        self._t0 = val[0]
        self._tf = val[-1]
        self._tau = np.array(val)


class Bezier(BezierParams):
    """Bezier curve for trajectory generation (bezier.py:148-889), GPU backed."""

    def __init__(self, cpts=None, t0=0.0, tf=1.0, tau=None):
        super().__init__(cpts=cpts, tau=tau, t0=t0, tf=tf)

    def __add__(self, curve):
        return self.add(curve)

    def __sub__(self, curve):
        return self.sub(curve)

    def __mul__(self, curve):
        return self.mul(curve)

    def __truediv__(self, curve):
        return self.div(curve)

    def __repr__(self):
        return 'Bezier({}, {}, {}, {})'.format(self.cpts, self.tau, self.t0, self.tf)

    def _f64(self):
        return np.ascontiguousarray(self.cpts, dtype=np.float64)

    # -- evaluation ------------------------------------------------------
    def _eval(self, tau):
        tau = np.ascontiguousarray(np.atleast_1d(tau), dtype=np.float64)
        c = _to_dev(self._f64())
        d_tau = _to_dev(tau)
        out = torch.empty((self.dim, tau.size), dtype=F64, device=c.device)
        _capi.call("bez_curve_eval", _ptr(c), _ptr(d_tau), self.dim, self.deg, tau.size,
                   float(self.t0), float(self.tf), _ptr(out), _stream())
        return out.cpu().numpy()

    def __call__(self, t):
        """bezier.py:187-203"""
        return self._eval(t)

    @property
    def curve(self):
        """bezier.py:240-262 (cached)"""
        if self._curve is None:
            self._curve = self._eval(self.tau)
        return self._curve

    @property
    def x(self):
        return Bezier(self.cpts[0], t0=self.t0, tf=self.tf)

    @property
    def y(self):
        if self.dim > 1:
            return Bezier(self.cpts[1], t0=self.t0, tf=self.tf)
        return None

    @property
    def z(self):
        if self.dim > 2:
            return Bezier(self.cpts[2], t0=self.t0, tf=self.tf)
        return None

    def copy(self):
        return Bezier(self.cpts, self.t0, self.tf)

    def plot(self, *args, **kwargs):
        raise NotImplementedError("plotting is outside the GPU hot path (SURVEY section 2: out of scope)")

    # -- arithmetic ---------------------------------------------------------
    def _aligned(self, other):
        if self.t0 == other.t0 and self.tf == other.tf:
            return self.cpts, other.cpts, self.t0, self.tf
        c1, c2 = _temporalAlignment(self, other)
        return c1.cpts, c2.cpts, c1.t0, c1.tf

    def add(self, other):
        """bezier.py:318-345"""
        a, b, t0, tf = self._aligned(other)
        if t0 >= tf:
            return None
        return Bezier(a + b, t0=t0, tf=tf)

    def sub(self, other):
        """bezier.py:347-374"""
        a, b, t0, tf = self._aligned(other)
        if t0 >= tf:
            return None
        return Bezier(a - b, t0=t0, tf=tf)

    def mul(self, multiplicand):
        """bezier.py:376-432"""
        if not isinstance(multiplicand, Bezier):
            msg = 'The multiplicand must be a {} object, not a {}'.format(Bezier, type(multiplicand))
            raise TypeError(msg)
        dim = self.dim
        if multiplicand.dim != dim:
            msg = ('The dimension of both Bezier curves must be the same.\n'
                   'The first dimension is {} and the second is {}'.format(dim, multiplicand.dim))
            raise ValueError(msg)
        m, n = self.deg, multiplicand.deg
        a, b = _to_dev(self._f64()), _to_dev(multiplicand._f64())
        out = torch.empty((dim, m + n + 1), dtype=F64, device=a.device)
        _capi.call("bez_curve_mul", _ptr(a), _ptr(b), _ptr(_table("prod", m, n)), dim, m, n, _ptr(out), _stream())
        newCurve = self.copy()
        newCurve.cpts = out.cpu().numpy()
        return newCurve

    def div(self, denominator):
        """bezier.py:434-467 (control-point-wise ratio -> RationalBezier)"""
        if not isinstance(denominator, Bezier):
            msg = ('The denominator must be a Bezier object, not a {}. '
                   'Or the module has been reloaded.').format(type(denominator))
            raise TypeError(msg)
        num, den = self._f64(), denominator._f64()
        with np.errstate(divide='ignore', invalid='ignore'):
            cpts = np.where(num == 0, 0.0, np.where(den == 0, np.inf, num / den))
        return RationalBezier(cpts.astype(np.float64), den.astype(np.float64), tau=self.tau, tf=self.tf)

    def elev(self, R=1):
        """bezier.py:469-495"""
        c = _to_dev(self._f64())
        out = torch.empty((self.dim, self.deg + R + 1), dtype=F64, device=c.device)
        _capi.call("bez_curve_elev", _ptr(c), _ptr(_table("elev", self.deg, int(R))), self.dim, self.deg,
                   int(R), _ptr(out), _stream())
        curveElev = self.copy()
        curveElev.cpts = out.cpu().numpy()
        return curveElev

    def diff(self):
        """bezier.py:497-519: derivative, elevated back to the same degree (Q3)."""
        c = _to_dev(self._f64())
        T = _to_dev(np.array([self.tf - self.t0]))
        out = torch.empty((self.dim, self.deg + 1), dtype=F64, device=c.device)
        _capi.call("bez_curve_diff", _ptr(c), _ptr(_table("elev", self.deg - 1, 1)), _ptr(T), self.dim,
                   self.dim, self.deg, _ptr(out), _stream())
        curveDot = self.copy()
        curveDot.cpts = out.cpu().numpy()
        return curveDot

    def integrate(self):
        """bezier.py:521-531"""
        areas = np.empty(self.dim)
        for d in range(self.dim):
            areas[d] = self.tf * sum(self.cpts[d]) / (self.deg + 1)
        return areas

    def normSquare(self):
        """bezier.py:869-889 (includes the dim/2 factor, Q1)"""
        c = _to_dev(self._f64())
        out = torch.empty((1, 2 * self.deg + 1), dtype=F64, device=c.device)
        _capi.call("bez_curve_normsq", _ptr(c), _ptr(_table("prod", self.deg, self.deg)), 1, self.dim,
                   self.deg, _ptr(out), _stream())
        newCurve = self.copy()
        newCurve.cpts = out.cpu().numpy()
        return newCurve

    # -- subdivision ------------------------------------------------------------
    def split(self, tDiv):
        """bezier.py:533-572"""
        c1, c2 = self.copy(), self.copy()
        if np.isnan(tDiv):
            print('[!] Warning, tDiv is {}, changing to 0.'.format(tDiv))
            tDiv = 0
        c = _to_dev(self._f64())
        # deCasteljauSplit(cpts, tDiv - t0, tf - t0) divides inside (bezier.py:1007)
        tl = _to_dev(np.array([(tDiv - self.t0) / (self.tf - self.t0)]))
        left, right = torch.empty_like(c), torch.empty_like(c)
        _capi.call("bez_split", _ptr(c), _ptr(tl), 1, self.dim, self.deg, _ptr(left), _ptr(right), _stream())
        c1.cpts = left.cpu().numpy()
        c1.t0 = self.t0
        c1.tf = tDiv
        c2.cpts = right.cpu().numpy()
        c2.t0 = tDiv
        c2.tf = self.tf
        return c1, c2

    def _extreme(self, dim, tol, maximum, glob, max_depth=64):
        # root call of the recursion: the caller's bound is already within tol of the extreme
        # control point (bezier.py:651-656 / 747-752)
        vals = self._f64()[dim]
        ext = vals.max() if maximum else vals.min()
        if np.abs(glob - ext) < tol:
            return float(ext)
        row = _to_dev(self._f64()[dim])
        n = self.deg
        scratch = torch.empty(int(_capi.lib.bez_extrema_scratch_doubles(1, n, max_depth)), dtype=F64,
                              device=row.device)
        out = torch.empty(1, dtype=F64, device=row.device)
        status = torch.zeros(1, dtype=torch.int32, device=row.device)
        _capi.call("bez_extrema", _ptr(row), 1, n, float(tol), int(maximum), max_depth, _ptr(scratch),
                   _ptr(out), _iptr(status), _stream())
        self.last_status = int(status.item())
        if self.last_status != 0:
            # the subdivision did not converge within max_depth levels; the reference's
            # recursion has no bound and ends in Python's RecursionError
            raise RecursionError('Bezier.%s: subdivision depth limit (%d) reached'
                                 % ('max' if maximum else 'min', max_depth))
        return float(out.item())

    def min(self, dim=0, globMin=-np.inf, tol=1e-6):
        """bezier.py:631-667"""
        return self._extreme(dim, tol, False, globMin)

    def max(self, dim=0, globMax=np.inf, tol=1e-6):
        """bezier.py:727-763"""
        return self._extreme(dim, tol, True, globMax)

    # -- distance routines ----------------------------------------------------
    def minDist(self, otherCurve, max_depth=200):
        """bezier.py:840-852 -> _minDist: (alpha, t1, t2)."""
        if (self.dim < 2 or self.dim > 3 or otherCurve.dim < 2 or otherCurve.dim > 3):
            err = ('Both curves must be either 2D or 3D, not {}D and {}D.').format(self.dim, otherCurve.dim)
            raise ValueError(err)
        out, status = min_dist_batch(self._f64()[None], otherCurve._f64()[None], max_depth=max_depth)
        self.last_status = int(status[0])
        return float(out[0, 0]), float(out[0, 1]), float(out[0, 2])

    def minDist2Poly(self, poly, max_depth=200):
        """bezier.py:854-857 -> _minDist2Poly: (alpha, t1, closest point)."""
        out, status = min_dist2poly_batch(self._f64()[None], np.asarray(poly, dtype=np.float64)[None],
                                          max_depth=max_depth)
        self.last_status = int(status[0])
        pt = -1 if (self.last_status & 2) else out[0, 2:5].copy()
        return float(out[0, 0]), float(out[0, 1]), pt

    def collCheck(self, otherCurve):
        """bezier.py:859-862 -> _collCheckBez2Bez"""
        v = float(coll_check_batch(self._f64()[None], otherCurve._f64()[None])[0])
        return 1 if v == 1.0 else (-1 if v == -1.0 else v)

    def collCheck2Poly(self, poly, max_nodes=200000):
        """bezier.py:864-867 -> _collCheckBez2Poly"""
        out, status = coll_check2poly_batch(self._f64()[None], np.asarray(poly, dtype=np.float64)[None],
                                            max_nodes=max_nodes)
        self.last_status = int(status[0])
        return int(out[0])


class RationalBezier(BezierParams):
    """bezier.py:894-900"""

    def __init__(self, cpts=None, weights=None, tau=None, tf=1.0):
        super().__init__(cpts=cpts, tau=tau, tf=tf)
        self._weights = np.array(weights, ndmin=2)


def _temporalAlignment(c1, c2):
    """Restricts two curves to the time window they share, [max t0, min tf], by
    subdividing whichever curve sticks out at either end (bezier.py:903-941; used by
    add/sub when the windows differ)."""
    t0 = max(c1.t0, c2.t0)
    tf = min(c1.tf, c2.tf)
    pieces = []
    for curve in (c1, c2):
        piece = curve.copy()
        if curve.t0 < t0:                 # starts early: keep the part after t0
            piece = piece.split(t0)[1]
        if curve.tf > tf:                 # ends late: keep the part before tf
            piece = piece.split(tf)[0]
        piece.t0, piece.tf = t0, tf
        pieces.append(piece)
    return pieces[0], pieces[1]


# ---------------------------------------------------------------------------
# batched entry points (additive): one warp per item, one launch per batch
def min_dist_batch(c1, c2, eps=1e-9, max_depth=200, max_nodes=1 << 20):
    """c1 [count, dim1, n1+1], c2 [count, dim2, n2+1] -> (out [count,3], status [count])."""
    c1 = np.ascontiguousarray(c1, dtype=np.float64)
    c2 = np.ascontiguousarray(c2, dtype=np.float64)
    count, dim1, n1 = c1.shape[0], c1.shape[1], c1.shape[2] - 1
    dim2, n2 = c2.shape[1], c2.shape[2] - 1
    a, b = _to_dev(c1), _to_dev(c2)
    scratch = torch.empty(int(_capi.lib.bez_mindist_scratch_doubles(count, n1, n2, max_depth)), dtype=F64,
                          device=a.device)
    out = torch.empty((count, 3), dtype=F64, device=a.device)
    status = torch.zeros(count, dtype=torch.int32, device=a.device)
    _capi.call("bez_mindist", _ptr(a), _ptr(b), count, dim1, dim2, n1, n2, float(eps), int(max_depth),
               int(max_nodes), _ptr(scratch), _ptr(out), _iptr(status), _stream())
    return out.cpu().numpy(), status.cpu().numpy()


def min_dist2poly_batch(c1, polys, eps=1e-6, max_depth=200, max_nodes=1 << 20):
    """c1 [count, dim, n+1], polys [count, m, 3] -> (out [count,5], status)."""
    c1 = np.ascontiguousarray(c1, dtype=np.float64)
    polys = np.ascontiguousarray(polys, dtype=np.float64)
    count, dim1, n1 = c1.shape[0], c1.shape[1], c1.shape[2] - 1
    a, p = _to_dev(c1), _to_dev(polys)
    scratch = torch.empty(int(_capi.lib.bez_mindist2poly_scratch_doubles(count, n1, max_depth)), dtype=F64,
                          device=a.device)
    out = torch.empty((count, 5), dtype=F64, device=a.device)
    status = torch.zeros(count, dtype=torch.int32, device=a.device)
    _capi.call("bez_mindist2poly", _ptr(a), _ptr(p), ctypes.c_void_p(0), count, dim1, n1, polys.shape[1],
               float(eps), int(max_depth), int(max_nodes), _ptr(scratch), _ptr(out), _iptr(status), _stream())
    return out.cpu().numpy(), status.cpu().numpy()


def coll_check_batch(c1, c2, eps=1e-9):
    c1 = np.ascontiguousarray(c1, dtype=np.float64)
    c2 = np.ascontiguousarray(c2, dtype=np.float64)
    count, dim1, n1 = c1.shape[0], c1.shape[1], c1.shape[2] - 1
    dim2, n2 = c2.shape[1], c2.shape[2] - 1
    a, b = _to_dev(c1), _to_dev(c2)
    scratch = torch.empty(int(_capi.lib.bez_collcheck_scratch_doubles(count, n1, n2)), dtype=F64, device=a.device)
    out = torch.empty(count, dtype=F64, device=a.device)
    _capi.call("bez_collcheck", _ptr(a), _ptr(b), count, dim1, dim2, n1, n2, float(eps), _ptr(scratch),
               _ptr(out), _stream())
    return out.cpu().numpy()


def coll_check2poly_batch(c1, polys, max_nodes=200000):
    c1 = np.ascontiguousarray(c1, dtype=np.float64)
    polys = np.ascontiguousarray(polys, dtype=np.float64)
    count, dim1, n1 = c1.shape[0], c1.shape[1], c1.shape[2] - 1
    a, p = _to_dev(c1), _to_dev(polys)
    scratch = torch.empty(int(_capi.lib.bez_collcheck2poly_scratch_doubles(count, n1)), dtype=F64, device=a.device)
    out = torch.empty(count, dtype=F64, device=a.device)
    status = torch.zeros(count, dtype=torch.int32, device=a.device)
    _capi.call("bez_collcheck2poly", _ptr(a), _ptr(p), ctypes.c_void_p(0), count, dim1, n1, polys.shape[1],
               int(max_nodes), _ptr(scratch), _ptr(out), _iptr(status), _stream())
    return out.cpu().numpy(), status.cpu().numpy()


def extrema_batch(rows, tol=1e-6, maximum=False, max_depth=64):
    """rows [count, n+1] -> (values [count], status [count])."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    count, n = rows.shape[0], rows.shape[1] - 1
    r = _to_dev(rows)
    scratch = torch.empty(int(_capi.lib.bez_extrema_scratch_doubles(count, n, max_depth)), dtype=F64,
                          device=r.device)
    out = torch.empty(count, dtype=F64, device=r.device)
    status = torch.zeros(count, dtype=torch.int32, device=r.device)
    _capi.call("bez_extrema", _ptr(r), count, n, float(tol), int(maximum), max_depth, _ptr(scratch), _ptr(out),
               _iptr(status), _stream())
    return out.cpu().numpy(), status.cpu().numpy()


def split_batch(cpts, t_local):
    """cpts [count, dim, n+1], t_local [count] -> (left, right)."""
    cpts = np.ascontiguousarray(cpts, dtype=np.float64)
    c = _to_dev(cpts)
    tl = _to_dev(np.asarray(t_local, dtype=np.float64))
    left, right = torch.empty_like(c), torch.empty_like(c)
    _capi.call("bez_split", _ptr(c), _ptr(tl), cpts.shape[0], cpts.shape[1], cpts.shape[2] - 1, _ptr(left),
               _ptr(right), _stream())
    return left.cpu().numpy(), right.cpu().numpy()
