"""GPU parity of the fused constraint kernels (through the C-ABI) against the
numpy oracle and the committed golden vectors from the reference.
Tolerance: north_star's fp64 relative error <= 1e-9 per constraint block
(max |gpu-ref| / max |ref|)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def gopt():
    import torch
    assert torch.cuda.is_available()
    from optimalbeziertrajectorygeneration_b200 import optimization
    yield optimization
    optimization.DEG_ELEV = 0


def _check_blocks(gopt, bezopt, x, g, prefix, E, names):
    gopt.DEG_ELEV = E
    fns = {"sep": bezopt.temporalSeparationConstraints, "maxspeed": bezopt.maxSpeedConstraints,
           "minspeed": bezopt.minSpeedConstraints}
    if "angrate" in names:
        fns["angrate"] = bezopt.maxAngularRateConstraints
    for name in names:
        key = "%s_%s_E%d" % (prefix, name, E)
        if key not in g.files:
            continue
        got = fns[name](x)
        assert got.shape == g[key].shape
        assert relerr(got, g[key]) < RTOL, key


def test_swarm_golden(gopt, golden):
    g = golden("constraints")
    b = gopt.BezOptimization(numVeh=36, dimension=3, degree=5, minimizeGoal='Euclidean', maxSep=0.9,
                             initPoints=g["swarm_initPts"], finalPoints=g["swarm_finalPts"])
    for xi in (0, 1):
        for E in (0, 10, 100):
            _check_blocks(gopt, b, g["swarm_x%d" % xi], g, "swarm_x%d" % xi, E,
                          ("sep", "maxspeed", "minspeed"))


def test_c4_like_golden(gopt, golden):
    from oracle.make_golden import synthetic_swarm_args
    g = golden("constraints")
    for N in (2, 16, 33):
        args, x = synthetic_swarm_args(N)
        b = gopt.BezOptimization(**args)
        _check_blocks(gopt, b, x, g, "c4_N%d" % N, 100, ("sep", "maxspeed", "minspeed"))


def test_example1_golden(gopt, golden):
    g = golden("constraints")
    b = gopt.BezOptimization(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1,
                             maxSpeed=5, maxAngRate=1, initPoints=[(0, 5), (3, 0)],
                             finalPoints=[(8, 4), (7, 10)], initSpeeds=[1, 1], finalSpeeds=[1, 1],
                             initAngs=[0, np.pi / 2], finalAngs=[0, np.pi / 2],
                             pointObstacles=[[3, 2], [6, 7]])
    assert np.array_equal(b.reshapeVector(g["ex1_x0"]), g["ex1_y0"])     # bit exact A0
    for xi in (0, 1):
        for E in (0, 10, 100):
            _check_blocks(gopt, b, g["ex1_x%d" % xi], g, "ex1_x%d" % xi, E,
                          ("sep", "maxspeed", "minspeed", "angrate"))


def test_c5_like_golden(gopt, golden):
    from oracle.make_golden import dubins_problem_args
    g = golden("constraints")
    for seed in (0, 1, 2):
        b = gopt.BezOptimization(**dubins_problem_args(seed))
        x = g["c5_s%d_x" % seed]
        assert np.array_equal(b.reshapeVector(x), g["c5_s%d_y" % seed])
        for E in ((100, 0, 5) if seed == 0 else (100,)):
            _check_blocks(gopt, b, x, g, "c5_s%d" % seed, E, ("sep", "maxspeed", "minspeed", "angrate"))


@pytest.mark.parametrize("dim,deg,N,E", [(1, 3, 5, 0), (2, 1, 4, 3), (3, 7, 9, 21), (2, 12, 7, 64),
                                         (3, 16, 5, 40), (3, 10, 40, 300), (2, 4, 300, 1)])
def test_random_models_vs_oracle(gopt, dim, deg, N, E):
    from oracle import bezier_oracle as O
    rng = np.random.default_rng(deg * 100 + N)
    args = dict(numVeh=N, dimension=dim, degree=deg, minimizeGoal='Euclidean', maxSep=0.7,
                minSpeed=0.3, maxSpeed=4.0, tf=float(rng.uniform(1, 30)),
                initPoints=rng.uniform(-5, 5, size=(N, dim)), finalPoints=rng.uniform(-5, 5, size=(N, dim)))
    b = gopt.BezOptimization(**args)
    x = rng.normal(size=b.nvar) * 3
    f = O.make_callables(O.Model(**args), E)
    gopt.DEG_ELEV = E
    assert np.array_equal(b.reshapeVector(x), O.reshape_vector(O.Model(**args), x))
    assert relerr(b.temporalSeparationConstraints(x), f["sep"](x)) < RTOL
    assert relerr(b.maxSpeedConstraints(x), f["maxspeed"](x)) < RTOL
    assert relerr(b.minSpeedConstraints(x), f["minspeed"](x)) < RTOL


def test_batched_and_pair_ranges(gopt):
    """evaluate_batch == per-x calls; a split pair range == the full range
    (what the multi-GPU sharding relies on)."""
    import torch
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(33)
    b = gopt.BezOptimization(**args)
    rng = np.random.default_rng(0)
    X = x[None, :] + rng.normal(size=(5, x.size)) * 0.1
    gopt.DEG_ELEV = 100
    res = b.evaluate_batch(X, which=("sep", "maxspeed", "minspeed"))
    for i in range(5):
        assert np.array_equal(res["sep"][i].cpu().numpy(), b.temporalSeparationConstraints(X[i]))
        assert np.array_equal(res["maxspeed"][i].cpu().numpy(), b.maxSpeedConstraints(X[i]))
    eng = b._engine(True)
    cpts, _ = eng.assemble(eng.upload(X), 100)
    full = eng.separation(cpts, 100, 0.9)
    P = full.shape[1]
    cuts = [0, 1, 17, 255, 256, 300, P]
    parts = [eng.separation(cpts, 100, 0.9, pair_begin=a, npairs=c - a) for a, c in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat(parts, dim=1), full)
    pm = torch.empty((5, P), dtype=torch.float64, device=full.device)
    eng.separation(cpts, 100, 0.9, pairmin=pm)
    assert torch.equal(pm, full.min(dim=2).values)


def test_single_vehicle_returns_none(gopt):
    b = gopt.BezOptimization(numVeh=1, dimension=2, degree=5, initPoints=[(0, 0)], finalPoints=[(1, 1)])
    assert b.temporalSeparationConstraints(np.zeros(b.nvar)) is None


# --------------------------------------------------------------------------
# A7: finite-difference Jacobians.  Judged against the exactly rounded quotient
# (oracle evaluated in exact rational arithmetic on the fp64 points), tolerance 1e-9 relative
# to the largest entry; the reference's own literal FD sits ~1e-7 away from
# that value (its noise floor) and is checked at that level.
JTOL = 1e-9


def _jac_check(gopt, args, x, E, golden_J=None, names=("sep", "maxspeed", "minspeed")):
    from oracle import bezier_oracle as O
    b = gopt.BezOptimization(**args)
    gopt.DEG_ELEV = E
    fld = O.make_callables(O.Model(**args), E, dtype=object)
    fns = {"sep": b.temporalSeparationConstraints_jac, "maxspeed": b.maxSpeedConstraints_jac,
           "minspeed": b.minSpeedConstraints_jac}
    for name in names:
        J = fns[name](x)
        Jex = O.fd_jacobian_exact(fld[name], x)
        assert J.shape == Jex.shape
        scale = np.abs(Jex).max()
        assert np.abs(J - Jex).max() / scale < JTOL, name
        # sparsity: rows that do not depend on x_k are exact zeros in both; entries that
        # are zero only by cancellation (e.g. the fixed end speed of a Dubins curve w.r.t.
        # tf) may be O(ulp) instead of 0
        assert np.all(np.abs(Jex[J == 0]) <= 1e-13 * scale), name
        assert np.all(np.abs(J[Jex == 0]) <= 1e-13 * scale), name
        if golden_J is not None and name in golden_J:
            Jref = golden_J[name]
            assert np.abs(J - Jref).max() / scale < 5e-6        # reference FD noise floor
            assert np.all(np.abs(Jref[J == 0]) <= 5e-6 * scale), name
            # columns of variables that move no curve of a row are exact zeros in both
            untouched = np.all(J == 0, axis=0)
            assert np.all(Jref[:, untouched] == 0), name


def test_jacobian_swarm(gopt, golden):
    from oracle.make_golden import synthetic_swarm_args
    g = golden("jacobian")
    args, x = synthetic_swarm_args(6, deg=5, seed=11)
    _jac_check(gopt, args, x, 10, {"sep": g["sw6_J_sep_E10"], "maxspeed": g["sw6_J_maxspeed_E10"]})


def test_jacobian_dubins_timeopt_obstacles(gopt, golden):
    from oracle.make_golden import dubins_problem_args
    g = golden("jacobian")
    args = dubins_problem_args(5, nobs=4, deg=6)
    _jac_check(gopt, args, g["dub_x"], 8, {"sep": g["dub_J_sep_E8"], "maxspeed": g["dub_J_maxspeed_E8"]})


def test_jacobian_example1(gopt, golden):
    g = golden("jacobian")
    args = dict(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1, maxSpeed=5,
                maxAngRate=1, initPoints=[(0, 5), (3, 0)], finalPoints=[(8, 4), (7, 10)],
                initSpeeds=[1, 1], finalSpeeds=[1, 1], initAngs=[0, np.pi / 2],
                finalAngs=[0, np.pi / 2], pointObstacles=[[3, 2], [6, 7]])
    for E in (0, 10):
        _jac_check(gopt, args, g["ex1_x"], E, {"sep": g["ex1_J_sep_E%d" % E],
                                               "maxspeed": g["ex1_J_maxspeed_E%d" % E]})


def test_jacobian_sweep_layout_matches_dense(gopt):
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(9, deg=7, seed=4)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    E = 33
    JT = eng.jac_separation(x, E, dense=True).cpu().numpy()          # [nvar, P*L]
    sw = eng.jac_separation(x, E, dense=False).cpu().numpy()         # [nvar*(N-1), L]
    N, L = eng.N, 2 * 7 + E + 1
    sw = sw.reshape(eng.nvar, N - 1, L)
    for k in range(eng.nvar):
        v = k // (eng.dim * eng.ncols)
        others = [u for u in range(N) if u != v]
        for uu, u in enumerate(others):
            i, j = min(v, u), max(v, u)
            p = i * (2 * N - i - 1) // 2 + (j - i - 1)
            assert np.array_equal(JT[k, p * L:(p + 1) * L], sw[k, uu])


def test_angrate_rejects_3d(gopt):
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(3)
    b = gopt.BezOptimization(**args)
    with pytest.raises(ValueError):
        b.maxAngularRateConstraints(x)


def test_jacobian_angrate_literal_fd(gopt, golden):
    """The angular-rate Jacobian is the batched literal quotient: it must agree
    with the reference's own FD Jacobian to the FD noise floor."""
    from oracle.make_golden import dubins_problem_args
    g = golden("jacobian")
    b = gopt.BezOptimization(**dubins_problem_args(5, nobs=4, deg=6))
    gopt.DEG_ELEV = 8
    J = b.maxAngularRateConstraints_jac(g["dub_x"])
    Jref = g["dub_J_angrate_E8"]
    assert J.shape == Jref.shape
    assert np.abs(J - Jref).max() / np.abs(Jref).max() < 1e-5


def test_objectives(gopt, golden):
    """A14: cost callables (swarm Euclidean objective at x0 from SURVEY section 4)."""
    from oracle.make_golden import synthetic_swarm_args
    g = golden("constraints")
    gopt.DEG_ELEV = 0
    b = gopt.BezOptimization(numVeh=36, dimension=3, degree=5, minimizeGoal='Euclidean', maxSep=0.9,
                             initPoints=g["swarm_initPts"], finalPoints=g["swarm_finalPts"])
    assert b.objectiveFunction(g["swarm_x0"]) == pytest.approx(382.0101330471013, rel=1e-13)
    args, x = synthetic_swarm_args(5, deg=6, seed=3)
    for goal in ("Euclidean", "Accel"):
        a = dict(args)
        a["minimizeGoal"] = goal
        bb = gopt.BezOptimization(**a)
        assert bb.objectiveFunction(x) == pytest.approx(float(g["obj_%s" % goal]), rel=1e-12)
    a = dict(args)
    a["minimizeGoal"] = "nonsense"
    with pytest.raises(ValueError):
        gopt.BezOptimization(**a).objectiveFunction
    tb = gopt.BezOptimization(numVeh=1, dimension=2, degree=5, minimizeGoal='TimeOpt', initPoints=[(0, 0)],
                              finalPoints=[(1, 1)])
    assert tb.objectiveFunction(np.arange(9.0)) == 8.0
