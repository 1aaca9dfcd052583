"""GPU parity on the reference-generated fixtures the CUDA path did not see in round 1
(VERDICT r01 "What's weak" 1):

  * the 121-vehicle sequential-swarm pickle (degree 3, L = 17) through
    ``bez_pair_sepsq_elev`` and through ``FrozenSwarm``, plus the pickle's NaN vehicles;
  * Example1's own vehicle-only separation closure at E = 0 / 30 / 100 (L = 21 / 51 / 121);
  * **active-pair flags**: ``pairmin < 0`` from the fused epilogue must be ``array_equal`` to
    ``row.min() < 0`` of the reference's vector (north_star: bit-identical flags), pairs with
    ``|min| <= 1e-9`` excluded and counted;
  * A13 on the ``Examples/ComplexObstacles.py`` / ``DrivingOnATrack.py`` setups.

Values: fp64, max|gpu-ref| / max|ref| <= 1e-9 per block.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, relerr

pytestmark = pytest.mark.gpu

RTOL = 1e-9
TIE = 1e-9
_report = {}


@pytest.fixture(scope="module")
def gopt():
    import torch
    assert torch.cuda.is_available()
    from optimalbeziertrajectorygeneration_b200 import optimization
    yield optimization
    optimization.DEG_ELEV = 0
    # near-tie bookkeeping of the flag tests ("report any excluded near-ties")
    out = os.path.join(ROOT, "gpurun_out")
    if _report and os.path.isdir(out):
        with open(os.path.join(out, "active_pair_flags_report.json"), "w") as f:
            json.dump(_report, f, indent=1, sort_keys=True)


def _raw_pair_eval(y, N, dim, deg, E, max_sep, B_rows=None):
    """[N*dim, deg+1] control-point matrix (the reference's `y`) -> (values [P, L], pairmin [P])
    straight through the C-ABI pair kernel (no BezOptimization model: the pickle is raw data)."""
    import torch
    from optimalbeziertrajectorygeneration_b200 import _capi, engine
    dev = torch.device("cuda", torch.cuda.current_device())
    plan = engine.Plan.get(deg, dim, E, dev.index)
    S = (dim * (deg + 1) + 1) // 2 * 2
    rows = np.zeros((1, N, S))
    rows[0, :, :dim * (deg + 1)] = np.asarray(y, dtype=np.float64).reshape(N, dim * (deg + 1))
    cpts = torch.as_tensor(rows, device=dev)
    P = N * (N - 1) // 2
    out = torch.full((1, P, plan.L), 7.0, dtype=torch.float64, device=dev)
    pm = torch.full((1, P), 7.0, dtype=torch.float64, device=dev)
    _capi.call("bez_pair_sepsq_elev", plan.handle, engine._ptr(cpts), 1, N, 0, P, float(max_sep) ** 2,
               engine._ptr(out), engine._ptr(pm), engine._stream())
    torch.cuda.synchronize()
    return out[0].cpu().numpy(), pm[0].cpu().numpy()


def _flags_equal(name, pairmin_gpu, ref_rows):
    """(pairmin < 0) vs (row.min() < 0) of the reference, near-ties excluded and counted."""
    ref_min = ref_rows.min(axis=1)
    tie = np.abs(ref_min) <= TIE
    got = pairmin_gpu < 0
    want = ref_min < 0
    assert np.array_equal(got[~tie], want[~tie]), name
    _report[name] = {"pairs": int(ref_min.size), "active": int(want.sum()), "excluded_near_ties": int(tie.sum())}
    return int(tie.sum())


# --------------------------------------------------------------------------
def test_pickle_swarm_pair_kernel_and_flags(gopt, golden):
    """Examples/SequentialSwarmLONG_...pickle: 121 finite degree-3 3-D trajectories, E = 10."""
    g = golden("constraints")
    N, L = 121, 2 * 3 + 10 + 1
    want = g["seq_sep_E10"].reshape(-1, L)
    vals, pm = _raw_pair_eval(g["seq_y"], N, 3, 3, 10, 0.9)
    assert relerr(vals, want) < RTOL
    assert np.array_equal(pm, vals.min(axis=1))
    _flags_equal("pickle_121_E10", pm, want)


def test_pickle_swarm_through_frozen_swarm(gopt, golden):
    """The same pickle through the sequential-planning caller: vehicle v against the v earlier
    trajectories = rows (i, v), i < v, of the reference's vector, one minimum each
    (Examples/SequentialSwarm.py:62-67)."""
    from optimalbeziertrajectorygeneration_b200.sequential import FrozenSwarm
    g = golden("constraints")
    N, L = 121, 17
    want = g["seq_sep_E10"].reshape(-1, L).min(axis=1)
    traj = g["seq_y"].reshape(N, 3, 4)

    def pidx(i, j):
        return i * (2 * N - i - 1) // 2 + (j - i - 1)
    for v in (1, 2, 40, 120):
        fs = FrozenSwarm(traj[:v], elev=10)
        got = fs.separation_minima(traj[v], 0.9)
        ref = np.array([want[pidx(i, v)] for i in range(v)])
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= RTOL * np.abs(g["seq_sep_E10"]).max()
        tie = np.abs(ref) <= TIE
        assert np.array_equal((got < 0)[~tie], (ref < 0)[~tie])


def test_pickle_nan_vehicles_propagate(gopt, golden):
    """229 of the pickle's 350 vehicles are NaN.  With NaN vehicles in the
    list every pair that contains one is NaN in all L values and in its minimum (the flag is
    then False, like `nan < 0` of the reference); finite pairs keep their bits."""
    g = golden("constraints")
    gs = golden("round2")
    N0, L = 121, 17
    ref_nan_rows = gs["seq_nan_rows"]                       # reference, 121 finite + 4 NaN vehicles
    yn = gs["seq_nan_y"]
    N = yn.shape[0] // 3
    # the pickle's unplanned vehicles have NaN interior control points (their end points are set)
    assert N > N0 and np.isnan(yn[3 * N0:]).any(axis=1).all() and np.isfinite(yn[:3 * N0]).all()
    vals, pm = _raw_pair_eval(yn, N, 3, 3, 10, 0.9)
    base_vals, base_pm = _raw_pair_eval(g["seq_y"], N0, 3, 3, 10, 0.9)
    iu, ju = np.triu_indices(N, 1)
    has_nan = ju >= N0
    assert np.array_equal(ref_nan_rows, has_nan)            # the reference: NaN rows = pairs with a NaN vehicle
    assert np.isnan(vals[has_nan]).all()
    assert np.isnan(pm[has_nan]).all()
    assert not (pm[has_nan] < 0).any()
    # finite pairs: same bits as the run without the NaN vehicles, and parity with the reference
    assert np.array_equal(vals[~has_nan], base_vals)
    assert np.array_equal(pm[~has_nan], base_pm)
    # (the reference's finite rows are bit-identical to seq_sep_E10, checked when the fixture was made)
    assert relerr(vals[~has_nan], g["seq_sep_E10"].reshape(-1, L)) < RTOL


@pytest.mark.parametrize("deg,E", [(10, 100), (5, 10), (10, 300)])
def test_nan_vehicle_on_every_kernel_family(gopt, deg, E):
    """The same property on the tensor-path shapes (L = 121), the small-L shapes and L > 128:
    one NaN control point makes all L values and the minimum of every pair of that vehicle NaN,
    all other pairs keep their bits."""
    from oracle.make_golden import synthetic_swarm_args
    from oracle import bezier_oracle as O
    N = 41
    args, x = synthetic_swarm_args(N, deg=deg)
    y = O.reshape_vector(O.Model(**args), x)
    base_vals, base_pm = _raw_pair_eval(y, N, 3, deg, E, 0.9)
    yn = y.copy()
    yn[3 * 17 + 1, 2] = np.nan                              # one interior control point of vehicle 17
    vals, pm = _raw_pair_eval(yn, N, 3, deg, E, 0.9)
    iu, ju = np.triu_indices(N, 1)
    bad = (iu == 17) | (ju == 17)
    assert np.isnan(vals[bad]).all() and np.isnan(pm[bad]).all()
    assert np.array_equal(vals[~bad], base_vals[~bad]) and np.array_equal(pm[~bad], base_pm[~bad])
    assert np.array_equal(base_pm, base_vals.min(axis=1))


def test_example1_own_separation_closure(gopt, golden):
    """Examples/Example1_DubinsCarTimeOptimal.py:19-49: the example's own vehicle-only
    separation function with its own `elev` argument (L = 21 / 51 / 121)."""
    g = golden("constraints")
    b = gopt.BezOptimization(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1,
                             maxSpeed=5, maxAngRate=1, initPoints=[(0, 5), (3, 0)],
                             finalPoints=[(8, 4), (7, 10)], initSpeeds=[1, 1], finalSpeeds=[1, 1],
                             initAngs=[0, np.pi / 2], finalAngs=[0, np.pi / 2])      # no point obstacles
    for xi in (0, 1):
        for E in (0, 30, 100):
            gopt.DEG_ELEV = E
            got = b.temporalSeparationConstraints(g["ex1_x%d" % xi])
            want = g["ex1_x%d_ownsep_E%d" % (xi, E)]
            assert got.shape == want.shape == (21 + E,)
            assert relerr(got, want) < RTOL, (xi, E)
            eng = b._engine(True)
            import torch
            cpts, _ = eng.assemble(eng.upload(g["ex1_x%d" % xi]), E)
            pm = torch.empty((1, 1), dtype=torch.float64, device=eng.device)
            eng.separation(cpts, E, 1, pairmin=pm)
            assert float(pm[0, 0]) == got.min()
            _flags_equal("ex1_x%d_ownsep_E%d" % (xi, E), pm[0].cpu().numpy(), want[None, :])


def test_active_pair_flags_vs_reference(gopt, golden):
    """north_star: bit-identical active-pair flags.  Reference flag of pair p = min_k c[p,k] < 0
    (what SequentialSwarm.py:65-67 reduces to); GPU flag = sign of the fused per-pair minimum."""
    import torch
    from oracle.make_golden import synthetic_swarm_args
    g = golden("constraints")
    excluded = 0
    # C3: the 36-vehicle swarm at x0 / x1, E in {0, 10, 100}
    b = gopt.BezOptimization(numVeh=36, dimension=3, degree=5, minimizeGoal='Euclidean', maxSep=0.9,
                             initPoints=g["swarm_initPts"], finalPoints=g["swarm_finalPts"])
    eng = b._engine(True)
    P = 36 * 35 // 2
    for xi in (0, 1):
        for E in (0, 10, 100):
            L = 11 + E
            cpts, _ = eng.assemble(eng.upload(g["swarm_x%d" % xi]), E)
            pm = torch.empty((1, P), dtype=torch.float64, device=eng.device)
            eng.separation(cpts, E, 0.9, pairmin=pm)
            want = g["swarm_x%d_sep_E%d" % (xi, E)].reshape(P, L)
            excluded += _flags_equal("swarm_x%d_E%d" % (xi, E), pm[0].cpu().numpy(), want)
    # the known minima of SURVEY section 4 at x0 make sure active pairs exist in the fixtures
    assert _report["swarm_x0_E0"]["active"] > 0 and _report["swarm_x0_E100"]["active"] > 0
    # C4-like, N = 33
    args, x = synthetic_swarm_args(33)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    P = 33 * 32 // 2
    cpts, _ = eng.assemble(eng.upload(x), 100)
    pm = torch.empty((1, P), dtype=torch.float64, device=eng.device)
    eng.separation(cpts, 100, args["maxSep"], pairmin=pm)
    excluded += _flags_equal("c4_N33_E100", pm[0].cpu().numpy(), g["c4_N33_sep_E100"].reshape(P, 121))
    # C4-like with a separation radius large enough that a good share of the pairs is active
    want = g["c4_N33_sep_E100"].reshape(P, 121) + 0.9 ** 2 - 40.0 ** 2
    eng.separation(cpts, 100, 40.0, pairmin=pm)
    excluded += _flags_equal("c4_N33_E100_maxSep40", pm[0].cpu().numpy(), want)
    assert 0 < _report["c4_N33_E100_maxSep40"]["active"] < P
    print("active-pair flags: %d near-tie pairs excluded in total" % excluded)


# --------------------------------------------------------------------------
# A13 on the two example setups that use it
_TRACKS = {
    # Examples/ComplexObstacles.py:19-39
    "ComplexObstacles": dict(
        t1=[[8, 9, 10, 11, 12, 13, 12, 11, 10, 9, 8], [8, 10, 12, 14, 20, 14, 12, 10, 10, 9, 8]],
        t2=[[18, 13, 9, 6, 4, 3, 4, 6, 9, 13, 18], [3, 3, 4, 4, 4, 5, 5, 5, 7, 8, 3]], final=(15, 15)),
    # Examples/DrivingOnATrack.py:18-38
    "DrivingOnATrack": dict(
        t1=[[0, 0, 0, 3, 4, 5, 6, 7, 10, 10, 10], [0, 3, 4, 5, 6, 6, 6, 6, 7, 8, 10]],
        t2=[[4, 4, 4, 7, 8, 9, 10, 11, 14, 14, 14], [0, 3, 4, 4, 4, 5, 5, 5, 7, 8, 10]], final=(12, 9)),
}


@pytest.mark.parametrize("name", sorted(_TRACKS))
def test_spatial_separation_on_example_setups(gopt, golden, name):
    """optimization.py:109-133 with the shipped setups: one Dubins vehicle (degree 10, 2-D) and
    two degree-10 track curves as shapeObstacles -> 3 pairs x (alpha, t1, t2) - maxSep.
    Checked pair by pair against the pure-Python restatement of the reference's _minDist
    (same depth / node budgets, so statuses must agree) and against the values frozen from the
    reference itself where it terminates (tests/golden/round2.npz)."""
    from oracle import gjk_oracle as G
    from optimalbeziertrajectorygeneration_b200 import bezier as gbez
    s = _TRACKS[name]
    tracks = [gbez.Bezier(s["t1"]), gbez.Bezier(s["t2"])]
    b = gopt.BezOptimization(numVeh=1, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=0.5, maxSpeed=5,
                             maxAngRate=0.5, initPoints=(2, 1), finalPoints=s["final"], initSpeeds=1,
                             finalSpeeds=1, initAngs=np.pi / 2, finalAngs=np.pi / 2, shapeObstacles=tracks)
    gs = golden("round2")
    x = b.generateGuess()
    x[-1] = 10
    assert np.array_equal(x, gs[name + "_x"])                    # generateGuess parity (host logic)
    b.spatial_max_nodes = 1 << 14
    b.spatial_on_limit = "nan"
    got = b.spatialSeparationConstraints(x)
    assert got.shape == (3, 3)
    y = b.reshapeVector(x)
    assert np.array_equal(y, gs[name + "_y"])
    curves = [y] + [t.cpts for t in tracks]
    k = 0
    for i in range(3):
        for j in range(i + 1, 3):
            a, t1, t2, st = G.min_dist(curves[i], curves[j], max_nodes=1 << 14)
            assert st == int(b.last_status[k]), (name, i, j)
            if st == 0:
                assert np.array_equal(got[k], np.array([a, t1, t2]) - 0.5), (name, i, j)
            else:
                assert np.all(np.isnan(got[k]))
            ref = gs["%s_ref_%d%d" % (name, i, j)]
            if np.isfinite(ref).all():                           # the reference terminated on this pair
                assert st == 0 and np.array_equal(got[k] + 0.5, ref), (name, i, j)
            k += 1
    # the default drop-in behaviour: a pair over budget raises like the reference's RecursionError
    b.spatial_on_limit = "raise"
    if (b.last_status != 0).any():
        with pytest.raises(RecursionError):
            b.spatialSeparationConstraints(x)
