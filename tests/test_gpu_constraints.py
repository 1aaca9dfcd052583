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
                          ("sep", "maxspeed", "minspeed"))


def test_c5_like_golden(gopt, golden):
    from oracle.make_golden import dubins_problem_args
    g = golden("constraints")
    for seed in (0, 1, 2):
        b = gopt.BezOptimization(**dubins_problem_args(seed))
        x = g["c5_s%d_x" % seed]
        assert np.array_equal(b.reshapeVector(x), g["c5_s%d_y" % seed])
        for E in ((100, 0, 5) if seed == 0 else (100,)):
            _check_blocks(gopt, b, x, g, "c5_s%d" % seed, E, ("sep", "maxspeed", "minspeed"))


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
