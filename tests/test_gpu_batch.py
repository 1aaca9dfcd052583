"""GPU parity of the problem-batch path (BASELINE configs[4], "C5": independent Dubins
problems with 16 point obstacles each, FD sweeps) against the reference's golden vectors
and against the single-problem closures."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gopt():
    import torch
    assert torch.cuda.is_available()
    from optimalbeziertrajectorygeneration_b200 import optimization
    yield optimization
    optimization.DEG_ELEV = 0


def _batch(seeds):
    from oracle.make_golden import dubins_problem_args
    from optimalbeziertrajectorygeneration_b200.batch import ProblemBatch
    sets = np.array([dubins_problem_args(s)["pointObstacles"] for s in seeds])
    template = {k: v for k, v in dubins_problem_args(seeds[0]).items() if k != "pointObstacles"}
    return ProblemBatch(template, sets)


def test_batch_values_match_reference_golden(gopt, golden):
    import torch
    g = golden("constraints")
    seeds = (0, 1, 2)
    pb = _batch(seeds)
    X = np.stack([g["c5_s%d_x" % s] for s in seeds])
    L = 2 * 10 + 100 + 1
    res = pb.evaluate(torch.as_tensor(X, device=pb.eng.device), 1, elev=100)
    assert pb.npairs_x == 16
    for i, s in enumerate(seeds):
        want_sep = g["c5_s%d_sep_E100" % s][:16 * L]          # vehicle-obstacle pairs come first
        assert relerr(res["sep"][i].cpu().numpy(), want_sep) < 1e-9
        assert relerr(res["maxspeed"][i].cpu().numpy(), g["c5_s%d_maxspeed_E100" % s]) < 1e-9
        assert relerr(res["angrate"][i].cpu().numpy(), g["c5_s%d_angrate_E100" % s]) < 1e-9


def test_batch_sweep_matches_single_problem_jacobians(gopt, golden):
    """M sweeps in one go == the per-problem Jacobian closures (closed form for separation and
    speed: agreement at the FD noise floor; literal FD for the angular rate: same formula)."""
    from oracle.make_golden import dubins_problem_args
    g = golden("constraints")
    seeds = (2, 0, 1, 0)
    pb = _batch(seeds)
    X = np.stack([g["c5_s%d_x" % s] for s in seeds])
    X[3] = X[3] + 0.01                                           # same obstacles as problem 1, other x
    gopt.DEG_ELEV = 100
    out = pb.sweep(X, elev=100)
    L = 121
    for i, s in enumerate(seeds):
        b = gopt.BezOptimization(**dubins_problem_args(s))
        f0, JT = out["sep"]
        assert np.array_equal(f0[i].cpu().numpy(), b.temporalSeparationConstraints(X[i])[:16 * L])
        Jref = b.temporalSeparationConstraints_jac(X[i])[:16 * L]            # [m, nvar]
        assert relerr(JT[i].cpu().numpy().T, Jref) < 5e-6
        f0, JT = out["maxspeed"]
        assert np.array_equal(f0[i].cpu().numpy(), b.maxSpeedConstraints(X[i]))
        assert relerr(JT[i].cpu().numpy().T, b.maxSpeedConstraints_jac(X[i])) < 5e-6
        f0, JT = out["angrate"]
        assert np.array_equal(f0[i].cpu().numpy(), b.maxAngularRateConstraints(X[i]))
        assert relerr(JT[i].cpu().numpy().T, b.maxAngularRateConstraints_jac(X[i])) < 1e-12


def test_batch_rejects_bad_shapes(gopt):
    pb = _batch((0, 1))
    with pytest.raises(ValueError):
        pb.fd_points(np.zeros((3, pb.nvar)))
    with pytest.raises(ValueError):
        from optimalbeziertrajectorygeneration_b200.batch import ProblemBatch
        ProblemBatch({}, np.zeros((4, 2)))
