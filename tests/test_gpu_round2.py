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
        # the reference's literal FD carries the rounding noise of f amplified by 1/h
        noise = 64 * np.spacing(abs(b.objectiveFunction(x))) / 1.4901161193847656e-08
        assert np.abs(grad - ref).max() < max(noise, 5e-6 * scale), goal
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


# --------------------------------------------------------------------------
# reduced results: packed active bitmask, compacted list, minima-only launches
@pytest.mark.parametrize("deg,E,dim", [(10, 100, 3), (5, 10, 3), (10, 30, 2), (10, 300, 3), (4, 6, 1)])
def test_active_bitmask_and_compacted_list(gopt, deg, E, dim):
    """bez_pair_sepsq_elev_ex: mask bit f = (pairmin[f] < threshold); list = exactly those
    (f, pairmin[f]); an overflowing list is reported by its count, the mask stays complete.
    Shapes on the tensor path fuse this into the epilogue; the others run a post-pass."""
    import torch
    from optimalbeziertrajectorygeneration_b200.engine import ActiveSet
    rng = np.random.default_rng(deg * 7 + E)
    N, B = 61, 3
    args = dict(numVeh=N, dimension=dim, degree=deg, minimizeGoal='Euclidean', maxSep=6.0, maxSpeed=1.0, tf=10.0,
                initPoints=rng.uniform(0, 30, size=(N, dim)), finalPoints=rng.uniform(0, 30, size=(N, dim)))
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    X = rng.uniform(0, 30, size=(B, b.nvar))
    cpts, tf = eng.assemble(eng.upload(X), E)
    P = N * (N - 1) // 2
    pm = torch.empty((B, P), dtype=torch.float64, device=eng.device)
    sep = eng.separation(cpts, E, 6.0, pairmin=pm)
    assert torch.equal(pm, sep.min(dim=2).values)
    ref = pm.cpu().numpy().ravel()
    for thr in (0.0, 50.0):
        want = ref < thr
        act = ActiveSet(B * P, capacity=B * P, device=eng.device, threshold=thr)
        pm2 = torch.empty_like(pm)
        sep2 = eng.separation(cpts, E, 6.0, pairmin=pm2, active=act)
        torch.cuda.synchronize()
        flags, idx, val, overflow = ActiveSet.decode(act.buf.cpu().numpy(), B * P, act.capacity)
        assert torch.equal(sep2, sep) and torch.equal(pm2, pm)
        assert not overflow and 0 < want.sum() < B * P
        assert np.array_equal(flags, want)
        assert np.array_equal(idx, np.nonzero(want)[0]) and np.array_equal(val, ref[want])
    # overflow: capacity smaller than the number of active items
    small = ActiveSet(B * P, capacity=5, device=eng.device, threshold=50.0)
    eng.separation(cpts, E, 6.0, pairmin=pm2, active=small)
    flags, idx, val, overflow = ActiveSet.decode(small.buf.cpu().numpy(), B * P, 5)
    assert overflow and idx.size == 5 and np.array_equal(flags, ref < 50.0)
    assert np.array_equal(val, ref[idx])
    # a second launch after reset() starts a fresh list
    small.reset()
    small.threshold = -1e300
    eng.separation(cpts, E, 6.0, pairmin=pm2, active=small)
    flags, idx, val, overflow = ActiveSet.decode(small.buf.cpu().numpy(), B * P, 5)
    assert not overflow and idx.size == 0 and not flags.any()
    # mask only (no list)
    mo = ActiveSet(B * P, capacity=0, device=eng.device, threshold=0.0)
    eng.separation(cpts, E, 6.0, pairmin=pm2, active=mo)
    flags, idx, val, overflow = ActiveSet.decode(mo.buf.cpu().numpy(), B * P, 0)
    assert np.array_equal(flags, ref < 0.0) and idx.size == 0
    # speed rows: per-vehicle minima + active vehicles
    vm = torch.empty((B, N), dtype=torch.float64, device=eng.device)
    va = ActiveSet(B * N, capacity=B * N, device=eng.device, threshold=0.0)
    spd = eng.speed(cpts, tf, E, -1.0, 1.0, vehmin=vm, active=va)
    assert torch.equal(vm, spd.min(dim=2).values)
    flags, idx, val, overflow = ActiveSet.decode(va.buf.cpu().numpy(), B * N, B * N)
    vref = vm.cpu().numpy().ravel()
    assert np.array_equal(flags, vref < 0) and np.array_equal(val, vref[vref < 0])
    assert 0 < (vref < 0).sum()


def test_minima_only_launch_and_pitched_pair_ranges(gopt):
    """d_out = NULL: the pair kernel produces minima without writing rows; pair sub-ranges
    with min_pitch write into one [B, P] matrix (what strong-scaling ranks do)."""
    import torch
    from oracle.make_golden import synthetic_swarm_args
    for deg, E in ((10, 100), (5, 0), (10, 30)):
        args, x = synthetic_swarm_args(53, deg=deg)
        b = gopt.BezOptimization(**args)
        eng = b._engine(True)
        X = x[None, :] + np.random.default_rng(1).normal(size=(2, x.size)) * 0.1
        cpts, _ = eng.assemble(eng.upload(X), E)
        P = 53 * 52 // 2
        pm = torch.empty((2, P), dtype=torch.float64, device=eng.device)
        sep = eng.separation(cpts, E, 0.9, pairmin=pm)
        only = torch.full((2, P), float("nan"), dtype=torch.float64, device=eng.device)
        assert eng.separation(cpts, E, 0.9, pairmin=only, rows=False) is None
        assert torch.equal(only, pm)
        whole = torch.full((2, P), float("nan"), dtype=torch.float64, device=eng.device)
        cuts = [0, 7, 300, 301, 1000, P]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            eng.separation(cpts, E, 0.9, pair_begin=lo, npairs=hi - lo, pairmin=whole[:, lo:], min_pitch=P, rows=False)
        assert torch.equal(whole, pm)
        with pytest.raises(ValueError):
            eng.separation(cpts, E, 0.9, rows=False)
    # shapes outside the tensor path cannot skip the rows
    args, x = synthetic_swarm_args(9, deg=10)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    cpts, _ = eng.assemble(eng.upload(x), 300)
    with pytest.raises(Exception):
        eng.separation(cpts, 300, 0.9, pairmin=torch.empty((1, 36), dtype=torch.float64, device=eng.device), rows=False)


def test_evaluate_sweep_active_matches_minima(gopt):
    """The reduced sweep (bitmask + compacted list + per-vehicle speed minima) against the
    per-pair-minimum sweep of round 1, incl. a ragged last chunk, a threshold that makes the
    first list overflow (recompute path) and the capacity carried over to the next call."""
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(40)
    b = gopt.BezOptimization(**args)
    X = x[None, :] + np.random.default_rng(5).normal(size=(11, x.size)) * 0.02
    ref = b.evaluate_sweep(X, elev=100, chunk=4)
    pm, sp = ref["pairmin"].copy(), ref["maxspeed"].copy().reshape(11, 40, 121)
    P = 40 * 39 // 2
    for thr, rows in ((0.0, True), (5000.0, True), (5000.0, False), (-1.0, True)):
        res = b.evaluate_sweep_active(X, elev=100, chunk=4, threshold=thr, rows=rows)
        assert res.nchunks == 3
        for k in range(3):
            lo, hi = 4 * k, min(11, 4 * k + 4)
            flags, ev, pair, val = res.pairs(k)
            want = pm[lo:hi] < thr
            assert np.array_equal(flags, want)
            assert np.array_equal(ev * P + pair, np.nonzero(want.ravel())[0])
            assert np.array_equal(val, pm[lo:hi][want])
            vmin, vflags = res.vehicles(k)
            assert np.array_equal(vmin, sp[lo:hi].min(axis=2))
            assert np.array_equal(vflags, vmin < 0)
        if rows:
            import torch
            assert torch.equal(b.workspace['sep'][:3].min(dim=2).values.cpu(), torch.as_tensor(pm[8:11]))
    assert (pm < 5000.0).mean() > 0.05                       # the overflow path really ran


def test_active_list_post_pass_equals_fused_append(gopt, monkeypatch):
    """Large launches (>= 65536 items) build the compacted list in a post-pass from the bitmask
    (the epilogue only packs the mask); BEZGPU_MMA_FLAGS=64 forces the fused append.  Both must
    give the same set of (index, minimum) entries."""
    import torch
    from optimalbeziertrajectorygeneration_b200.engine import ActiveSet
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(200)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    B, P = 4, 200 * 199 // 2
    X = x[None, :] + np.random.default_rng(2).normal(size=(B, x.size)) * 0.1
    cpts, _ = eng.assemble(eng.upload(X), 100)
    pm = torch.empty((B, P), dtype=torch.float64, device=eng.device)
    got = {}
    for flags in ("0", "64"):
        monkeypatch.setenv("BEZGPU_MMA_FLAGS", flags)
        act = ActiveSet(B * P, capacity=B * P, device=eng.device, threshold=400.0)
        eng.separation(cpts, 100, 0.9, pairmin=pm, active=act, rows=False)
        torch.cuda.synchronize()
        got[flags] = ActiveSet.decode(act.buf.cpu().numpy(), B * P, B * P)
    ref = pm.cpu().numpy().ravel()
    for flags in got:
        fl, idx, val, over = got[flags]
        assert not over and np.array_equal(fl, ref < 400.0)
        assert np.array_equal(idx, np.nonzero(ref < 400.0)[0]) and np.array_equal(val, ref[ref < 400.0])
    assert 100 < (ref < 400.0).sum() < B * P
