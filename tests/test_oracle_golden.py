"""Pins the numpy oracle (oracle/bezier_oracle.py) against golden vectors the
UNMODIFIED reference produced (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import relerr
from oracle import bezier_oracle as O
from oracle.make_golden import dubins_problem_args, synthetic_swarm_args

TOL = 1e-12     # oracle vs reference values (summation order differs from BLAS)


def test_tables_bit_equal(golden):
    g = golden("algebra")
    assert np.array_equal(O.elev_matrix(20, 100), g["elevMatrix_20_100"])
    assert np.array_equal(O.elev_matrix(10, 3), g["elevMatrix_10_3"])
    # prodMatrix(N)[j, N*i+j] == W[i, j-i]   (bezier.py:1174)
    N = 10
    P = g["prodMatrix_10"]
    W = O.prod_weights(N)
    dense = np.zeros_like(P)
    for i in range(N + 1):
        for j in range(N + 1):
            dense[i + j, i * (N + 1) + j] = W[i, j]
    assert np.array_equal(dense, P)
    C = g["bezProductCoefficients_7"]
    W = O.prod_weights(7, 7)
    dense = np.zeros_like(C)
    for i in range(8):
        for j in range(8):
            dense[i * 8 + j, i + j] = W[i, j]
    assert np.array_equal(dense, C)


def test_algebra(golden):
    g = golden("algebra")
    for ci in range(int(g["ncases"])):
        k = "c%02d_" % ci
        c, tf, other = g[k + "cpts"], float(g[k + "tf"]), g[k + "other"]
        assert relerr(O.norm_square(c)[None, :], g[k + "normsq"]) < TOL
        for R in (0, 1, 7, 30):
            assert relerr(O.elev(c, R), g[k + "elev%d" % R]) < TOL
        assert relerr(O.diff(c, tf), g[k + "diff"]) < TOL
        # degree-1 curves have an exactly-zero 2nd derivative; the reference's BLAS
        # leaves O(ulp) residue there, so judge against the scale of the inputs
        n = c.shape[1] - 1
        scale = np.abs(c).max() * (n / tf) ** 2
        assert np.abs(O.diff(O.diff(c, tf), tf) - g[k + "diff2"]).max() < 1e-13 * scale
        prod = np.vstack([O.mul(c[d], other[d]) for d in range(c.shape[0])])
        assert relerr(prod, g[k + "mul"]) < TOL
        l, r = O.split(c, float(g[k + "tdiv"]), 0.0, tf)
        assert np.array_equal(l, g[k + "split_l"])          # no-FMA de Casteljau: bit exact (Q13)
        assert np.array_equal(r, g[k + "split_r"])
        assert np.array_equal(O.de_casteljau_eval(c, g[k + "tau"], 0.0, tf), g[k + "eval"])


def _ex1_model():
    return O.Model(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1,
                   maxSpeed=5, maxAngRate=1, initPoints=[(0, 5), (3, 0)],
                   finalPoints=[(8, 4), (7, 10)], initSpeeds=[1, 1], finalSpeeds=[1, 1],
                   initAngs=[0, np.pi / 2], finalAngs=[0, np.pi / 2],
                   pointObstacles=[[3, 2], [6, 7]])


def test_example1_constraints(golden):
    g = golden("constraints")
    m = _ex1_model()
    assert np.array_equal(O.reshape_vector(m, g["ex1_x0"]), g["ex1_y0"])
    for xi in (0, 1):
        x = g["ex1_x%d" % xi]
        for E in (0, 30, 100):
            own = O.temporal_separation(O.reshape_vector(m, x), 2, 2, 1, E)
            assert relerr(own, g["ex1_x%d_ownsep_E%d" % (xi, E)]) < TOL
        for E in (0, 10, 100):
            f = O.make_callables(m, E)
            for name in ("sep", "maxspeed", "minspeed", "angrate"):
                key = "ex1_x%d_%s_E%d" % (xi, name, E)
                if key in g.files:
                    assert relerr(f[name](x), g[key]) < 1e-11, key


def test_swarm_constraints(golden):
    g = golden("constraints")
    m = O.Model(numVeh=36, dimension=3, degree=5, minimizeGoal='Euclidean', maxSep=0.9,
                initPoints=g["swarm_initPts"], finalPoints=g["swarm_finalPts"])
    assert m.nvar == 432
    for xi in (0, 1):
        x = g["swarm_x%d" % xi]
        for E in (0, 10, 100):
            f = O.make_callables(m, E)
            for name in ("sep", "maxspeed", "minspeed"):
                assert relerr(f[name](x), g["swarm_x%d_%s_E%d" % (xi, name, E)]) < TOL
    # SURVEY section 4 golden minima at x0
    assert O.make_callables(m, 0)["sep"](g["swarm_x0"]).min() == pytest.approx(-4.809999999999999, rel=1e-13)
    assert O.make_callables(m, 100)["sep"](g["swarm_x0"]).min() == pytest.approx(-1.031434528773978, rel=1e-12)
    y0 = O.reshape_vector(m, g["swarm_x0"])
    assert O.euclidean_objective(y0, 36, 3) == pytest.approx(float(g["swarm_objective_x0"]), rel=1e-13)


def test_c4_like(golden):
    g = golden("constraints")
    for N in (2, 16, 33):
        args, x = synthetic_swarm_args(N)
        assert np.array_equal(x, g["c4_N%d_x" % N])
        f = O.make_callables(O.Model(**args), 100)
        for name in ("sep", "maxspeed", "minspeed"):
            assert relerr(f[name](x), g["c4_N%d_%s_E100" % (N, name)]) < TOL


def test_c5_like(golden):
    g = golden("constraints")
    for seed in (0, 1, 2):
        m = O.Model(**dubins_problem_args(seed))
        x = g["c5_s%d_x" % seed]
        assert np.array_equal(O.reshape_vector(m, x), g["c5_s%d_y" % seed])
        for E in ((100, 0, 5) if seed == 0 else (100,)):
            f = O.make_callables(m, E)
            for name in ("sep", "maxspeed", "minspeed", "angrate"):
                assert relerr(f[name](x), g["c5_s%d_%s_E%d" % (seed, name, E)]) < 1e-10, (seed, E, name)


def test_sequential_swarm_pickle(golden):
    g = golden("constraints")
    out = O.temporal_separation(g["seq_y"], 121, 3, 0.9, 10)
    assert relerr(out, g["seq_sep_E10"]) < TOL


def test_objectives(golden):
    g = golden("constraints")
    args, x = synthetic_swarm_args(5, deg=6, seed=3)
    m = O.Model(**args)
    y = O.reshape_vector(m, x)
    assert O.euclidean_objective(y, 5, 3) == pytest.approx(float(g["obj_Euclidean"]), rel=1e-13)
    assert O.accel_objective(y, 5, 3, m.tf, 0) == pytest.approx(float(g["obj_Accel"]), rel=1e-12)


def test_fd_jacobian_vs_reference_noise_floor(golden):
    """The reference's FD Jacobian is only reproducible to its own FD noise
    (ulp(f)/h ~ 1e-16*|f|/1.5e-8); the oracle's literal FD must agree with it
    to that level and both must sit equally close to the exactly rounded
    quotient (hard part 1)."""
    g = golden("jacobian")
    args, x = synthetic_swarm_args(6, deg=5, seed=11)
    assert np.array_equal(x, g["sw6_x"])
    m = O.Model(**args)
    f64 = O.make_callables(m, 10)
    fld = O.make_callables(m, 10, dtype=object)
    for name in ("sep", "maxspeed"):
        Jref = g["sw6_J_%s_E10" % name]
        Jlit = O.fd_jacobian(f64[name], x)
        Jex = O.fd_jacobian_exact(fld[name], x)
        scale = np.abs(Jex).max()
        assert np.abs(Jlit - Jref).max() / scale < 5e-6
        assert np.abs(Jref - Jex).max() / scale < 5e-6
        # structural zeros are exact zeros everywhere
        assert np.array_equal(Jref == 0, Jex == 0)
        assert np.array_equal(Jlit == 0, Jex == 0)


def test_fd_steps_follow_scipy_rule_for_batches():
    """engine.fd_steps (host logic, used by batch.ProblemBatch.fd_points for [M, nvar] arrays):
    h = sqrt(eps) unless x + h == x, dx = (x + h) - x  (scipy/optimize/_numdiff.py:585-596)."""
    import importlib.util
    import os
    import sys
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # engine imports the ctypes binding at import time; only the static helper is needed here
    src = open(os.path.join(root, "optimalbeziertrajectorygeneration_b200", "engine.py")).read()
    start = src.index("    @staticmethod\n    def fd_steps")
    end = src.index("    def _direction_rows")
    ns = {"np": np}
    exec("class E:\n" + src[start:end], ns)
    X = np.array([[0.0, 1.0, -3.5, 1e9, -1e9], [2.5e-9, 7.0, 1e17, -1e17, 5.0]])
    h, dx = ns["E"].fd_steps(X)
    assert h.shape == X.shape and dx.shape == X.shape
    rel = 1.4901161193847656e-08
    small = np.abs(X) < 1e7
    assert np.all(h[small] == rel) and np.all(dx[small] == (X[small] + rel) - X[small])
    assert np.all(dx != 0.0)
    big = ~small & (np.abs(X) > 1e16)
    assert np.all(np.abs(h[big]) >= rel * np.abs(X[big]) * 0.999)
