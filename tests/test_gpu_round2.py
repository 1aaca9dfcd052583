"""Round-2 GPU tests: temporal alignment of add/sub, the engine's buffer / device contract
(ADVICE r01), A13 as a batched constraint with a Jacobian and polytope obstacles, objective
gradients."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def gopt():
    import torch
    assert torch.cuda.is_available()
    from optimalbeziertrajectorygeneration_b200 import optimization
    yield optimization
    optimization.DEG_ELEV = 0


def test_add_sub_across_time_windows(golden):
    """Bezier.add / sub with differing [t0, tf] (bezier.py:318-374 -> _temporalAlignment
    :903-941, rewritten here): bit-exact control points and windows, None when disjoint."""
    from optimalbeziertrajectorygeneration_b200 import bezier as gbez
    g = golden("round2")
    for k in range(int(g["nalign"])):
        a0, a1, b0, b1 = g["align%d_win" % k]
        A = gbez.Bezier(g["align%d_a" % k].copy(), t0=a0, tf=a1)
        Bc = gbez.Bezier(g["align%d_b" % k].copy(), t0=b0, tf=b1)
        for name, res in (("add", A + Bc), ("sub", A - Bc)):
            assert np.array_equal(res.cpts, g["align%d_%s" % (k, name)]), (k, name)
            assert np.array_equal([res.t0, res.tf], g["align%d_%s_win" % (k, name)])
    A = gbez.Bezier(np.zeros((2, 4)), t0=0.0, tf=1.0)
    Bc = gbez.Bezier(np.ones((2, 4)), t0=2.0, tf=3.0)
    assert A.add(Bc) is None and A.sub(Bc) is None            # bezier.py:342-343


def test_zero_copy_results_of_different_closures_do_not_alias(gopt):
    """SLSQP collects [con['fun'](x) for con in cons] before concatenating: with
    zero_copy_results every closure must own its pinned buffer."""
    from oracle.make_golden import dubins_problem_args
    b = gopt.BezOptimization(**dubins_problem_args(1, nobs=3))
    x = b.generateGuess(std=0.3, seed=4)
    gopt.DEG_ELEV = 6
    names = ["temporalSeparationConstraints", "minSpeedConstraints", "maxSpeedConstraints",
             "maxAngularRateConstraints"]
    b.zero_copy_results = False
    want = [getattr(b, n)(x) for n in names]
    wantJ = [getattr(b, n + "_jac")(x) for n in names]
    b.zero_copy_results = True
    got = [getattr(b, n)(x) for n in names]                   # all four alive at once
    gotJ = [getattr(b, n + "_jac")(x) for n in names]
    for w, g_ in zip(want + wantJ, got + gotJ):
        assert np.array_equal(w, g_)
    assert np.array_equal(np.concatenate(got), np.concatenate(want))
    assert np.array_equal(np.vstack(gotJ), np.vstack(wantJ))


def test_back_to_back_uploads_and_workspace_validation(gopt):
    import torch
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(12)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    X1 = x[None, :] + 1.0
    X2 = x[None, :] - 1.0
    d1 = eng.upload(X1)
    d2 = eng.upload(X2)                                       # must not overwrite the source of d1's copy
    d3 = eng.upload(X1 * 2)
    torch.cuda.synchronize()
    assert np.array_equal(d1.cpu().numpy(), X1) and np.array_equal(d2.cpu().numpy(), X2)
    assert np.array_equal(d3.cpu().numpy(), X1 * 2)
    cpts, tf = eng.assemble(d1, 100)
    P = 12 * 11 // 2
    ok = torch.empty((1, P, 121), dtype=torch.float64, device=eng.device)
    eng.separation(cpts, 100, 0.9, out=ok)
    for bad in (torch.empty((1, P, 120), dtype=torch.float64, device=eng.device),      # wrong size
                torch.empty((1, P, 121), dtype=torch.float32, device=eng.device),      # wrong dtype
                torch.empty((1, P, 242), dtype=torch.float64, device=eng.device)[:, :, ::2],   # strided
                torch.empty((1, P, 121), dtype=torch.float64)):                        # host tensor
        with pytest.raises(ValueError):
            eng.separation(cpts, 100, 0.9, out=bad)
    with pytest.raises(ValueError):
        eng.separation(cpts, 100, 0.9, pairmin=torch.empty((1, P + 1), dtype=torch.float64, device=eng.device))
    with pytest.raises(ValueError):
        eng.speed(cpts, tf, 100, -1.0, 25.0, out=torch.empty((1, 12, 100), dtype=torch.float64, device=eng.device))
    with pytest.raises(ValueError):                           # peer stores need the full pair list
        eng.separation(cpts, 100, 0.9, pair_begin=1, npairs=P - 1,
                       pairmin=torch.empty((1, P - 1), dtype=torch.float64, device=eng.device),
                       peer_ptrs=[ok.data_ptr()])
    with pytest.raises(ValueError):
        eng.separation(cpts, 100, 0.9, pair_begin=3, npairs=P)


def test_model_edit_rebuilds_device_state(gopt):
    gopt.DEG_ELEV = 0
    b = gopt.BezOptimization(numVeh=1, dimension=2, degree=5, initPoints=[(0, 0)], finalPoints=[(4, 4)],
                             pointObstacles=[[1.0, 2.0]])
    x = b.generateGuess(std=0.2, seed=1)
    c1 = b.temporalSeparationConstraints(x)
    b.pointObstacles = [[1.0, 2.0], [3.0, 1.0]]               # the reference re-reads this every call
    c2 = b.temporalSeparationConstraints(x)
    assert c1.shape == (11,) and c2.shape == (33,)
    assert np.array_equal(c2[:11], c1)


def test_bezier_extrema_status_and_glob_bounds():
    from optimalbeziertrajectorygeneration_b200 import bezier as gbez
    c = gbez.Bezier(np.array([0.0, 3.0, -2.0, 1.0, 0.5]))
    true_min = c.min()
    assert true_min > -2.0
    # root call: |globMin - min control point| < tol returns the control point (bezier.py:651-656)
    assert c.min(globMin=-2.0 + 1e-9) == -2.0
    assert c.max(globMax=3.0 - 1e-9) == 3.0
    assert c.last_status == 0


# --------------------------------------------------------------------------
# A13 widened: Jacobian, polytope obstacles, mixed degrees
def _spatial_model(gopt, shapes, numVeh=1):
    return gopt.BezOptimization(numVeh=numVeh, dimension=2, degree=6, minimizeGoal='Euclidean', maxSep=0.5,
                                initPoints=[(0.0, 0.0), (0.0, 3.0)][:numVeh],
                                finalPoints=[(10.0, 1.0), (10.0, 4.0)][:numVeh], shapeObstacles=shapes)


def test_spatial_jacobian_is_the_batched_fd_of_the_closure(gopt):
    """spatialSeparationConstraints_jac = SciPy's 2-point FD of spatialSeparationConstraints,
    but all (nvar+1) x pairs go through one launch per pair group."""
    from scipy.optimize._numdiff import approx_derivative
    from optimalbeziertrajectorygeneration_b200 import bezier as gbez
    rng = np.random.default_rng(5)
    obs = gbez.Bezier(np.array([np.linspace(2, 8, 5), np.full(5, 6.0)]) + rng.normal(size=(2, 5)) * 0.2)   # degree 4
    b = _spatial_model(gopt, [obs], numVeh=2)
    x = b.generateGuess(std=0.1, seed=2)
    f0 = b.spatialSeparationConstraints(x)
    assert f0.shape == (3, 3) and (b.last_status == 0).all()
    J = b.spatialSeparationConstraints_jac(x)
    Jref = approx_derivative(lambda v: b.spatialSeparationConstraints(v).ravel(), x, method='2-point',
                             abs_step=1.4901161193847656e-08)
    assert J.shape == Jref.shape == (9, x.size)
    assert np.array_equal(J, Jref)


def test_spatial_polytope_and_mixed_degree_obstacles(gopt):
    """shapeObstacles may mix Bezier curves of other degrees and convex polytopes (vertex
    arrays): curve-curve pairs -> bez_mindist, curve-polytope -> bez_mindist2poly,
    polytope-polytope -> bez_gjk; rows keep the lexicographic pair order."""
    from oracle import gjk_oracle as G
    from optimalbeziertrajectorygeneration_b200 import bezier as gbez
    curve = gbez.Bezier(np.array([[2.0, 4.0, 6.0, 8.0], [6.0, 7.0, 6.5, 6.0]]))               # degree 3
    poly_a = np.array([[3.0, 3.0], [4.0, 3.0], [4.0, 4.0], [3.0, 4.0]])                        # 2-D square
    poly_b = np.array([[7.0, -3.0, 0.0], [8.0, -3.0, 0.0], [7.5, -2.0, 0.0]])
    b = _spatial_model(gopt, [curve, poly_a, poly_b])
    x = b.generateGuess(std=0.05, seed=3)
    got = b.spatialSeparationConstraints(x) + 0.5
    assert got.shape == (6, 3)
    y = b.reshapeVector(x)
    pa3 = np.hstack([poly_a, np.zeros((4, 1))])
    a, t1, t2, st = G.min_dist(y, curve.cpts)
    assert st == 0 and np.array_equal(got[0], [a, t1, t2])                                    # vehicle-curve
    for row, poly in ((1, pa3), (2, poly_b)):                                                 # vehicle-polytope
        al, tt, pt, st = G.min_dist2poly(y, poly)
        assert st == 0 and np.array_equal(got[row], [al, tt, al])
        assert gbez.Bezier(y).minDist2Poly(poly)[0] == al
    for row, poly in ((3, pa3), (4, poly_b)):                                                 # curve-polytope
        al, tt, pt, st = G.min_dist2poly(curve.cpts, poly)
        assert st == 0 and np.array_equal(got[row], [al, tt, al])
    f, p1, p2, d = G.gjk_new(pa3, poly_b)                                                     # polytope-polytope
    assert f == 1 and np.array_equal(got[5], [d, d, d])
    J = b.spatialSeparationConstraints_jac(x)
    assert J.shape == (18, x.size)
    assert not J[9:].any()                                    # obstacle-obstacle rows are constants
    assert J[:9].any()
    with pytest.raises(TypeError):
        _spatial_model(gopt, [np.zeros((3, 4))]).spatialSeparationConstraints(x)


# --------------------------------------------------------------------------
# (f)2: objective gradients
def test_objective_gradients(gopt, golden):
    """objectiveFunction_jac == SciPy's 2-point gradient of the callable: within the FD noise
    floor of the reference's own gradient (ulp(f)/h ~ 1e-6 absolute here), and <= 1e-9 of the
    exactly rounded quotient (long double for the Euclidean length, exact rationals for the
    quadratic accel objective)."""
    from fractions import Fraction
    from oracle import bezier_oracle as O
    from oracle.make_golden import synthetic_swarm_args
    g, gc = golden("round2"), golden("constraints")
    args, x = synthetic_swarm_args(5, deg=6, seed=3)
    assert np.array_equal(x, gc["obj_Euclidean_x"])
    h, dx = O.fd_steps(x)

    def fd(fun):
        f0 = fun(x)
        out = []
        for k in range(x.size):
            x1 = x.copy()
            x1[k] = x[k] + h[k]
            out.append((fun(x1) - f0) / dx[k])
        return np.array([float(v) for v in out])

    def euclid_ld(xx):
        m = O.Model(**args)
        y = O.reshape_vector(m, xx).astype(np.longdouble)
        d = np.diff(y.reshape(5, 3, 7), axis=2)
        return np.sqrt((d * d).sum(axis=1)).sum()

    def accel_exact(E):
        def f(xx):
            m = O.Model(**args)
            y = O._cast(O.reshape_vector(m, xx), object)
            tf = Fraction(float(m.tf))
            s = Fraction(0)
            for i in range(5):
                a = O.diff(O.diff(y[3 * i:3 * i + 3], tf, dtype=object), tf, dtype=object)
                s = s + O.elev(O.norm_square(a, dtype=object), E, dtype=object).sum()
            return s
        return f

    for goal, E, key, exact in (("Euclidean", 0, "obj_Euclidean_grad", euclid_ld),
                                ("Accel", 0, "obj_Accel_grad", accel_exact(0)),
                                ("Accel", 7, "obj_Accel_E7_grad", accel_exact(7))):
        a = dict(args)
        a["minimizeGoal"] = goal
        b = gopt.BezOptimization(**a)
        gopt.DEG_ELEV = E
        grad = b.objectiveFunction_jac(x)
        ref = g[key]
        scale = np.abs(ref).max()
        assert grad.shape == ref.shape == (x.size,)
        assert np.abs(grad - ref).max() / scale < 5e-6, goal            # reference FD noise floor
        ex = fd(exact)
        assert np.abs(grad - ex).max() / scale < (1e-9 if goal == "Accel" else 1e-8), goal
        # batched objective values: one launch for many x
        X = np.stack([x, x + 0.1, x - 0.2])
        vals = b.objectiveFunction(X)
        assert vals.shape == (3,) and vals[0] == b.objectiveFunction(x)
        assert vals[1] == b.objectiveFunction(X[1]) and vals[2] == b.objectiveFunction(X[2])
    gopt.DEG_ELEV = 7
    assert b.objectiveFunction(x) == pytest.approx(float(g["obj_Accel_E7"]), rel=1e-12)
    gopt.DEG_ELEV = 0
    t = gopt.BezOptimization(numVeh=1, dimension=2, degree=5, minimizeGoal='TimeOpt', initPoints=[(0, 0)],
                             finalPoints=[(1, 1)])
    gt = t.objectiveFunction_jac(np.arange(9.0))
    assert gt.shape == (9,) and gt[-1] == 1.0 and not gt[:-1].any()
    with pytest.raises(ValueError):
        a = dict(args)
        a["minimizeGoal"] = "nonsense"
        gopt.BezOptimization(**a).objectiveFunction_jac
