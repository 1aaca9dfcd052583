"""TEST INFRASTRUCTURE ONLY -- round-2 golden vectors from the UNMODIFIED reference:

  tests/golden/round2.npz
    * <setup>_x, <setup>_y        generateGuess() (x[-1] = 10) and reshapeVector(x) of the
                                  Examples/ComplexObstacles.py:19-42 / DrivingOnATrack.py:18-41 setups
    * <setup>_ref_<i><j>          Bezier.minDist of pair (i, j) among [vehicle, track1, track2]
                                  (what optimization.py:109-133 collects) where the reference
                                  terminates within the time limit with a real answer; NaN where it
                                  does not (timeout, or its depth sentinel (-1,-1,-1))
    * seq_nan_y, seq_nan_rows     the sequential-swarm pickle with 4 of its NaN vehicles kept: which rows
                                  of the reference's vector are NaN (the finite rows equal seq_sep_E10)
    * guess_*                     generateGuess(std, seed) draws (RNG order pin for the rewrite)
    * obj_*_grad                  SciPy's 2-point gradient of the reference's objectiveFunction
    * align<k>_*                  Bezier.add / Bezier.sub of curves whose time windows differ
                                  (_temporalAlignment, bezier.py:903-941): inputs and results

--reuse-mindist keeps the (slow, minutes) minDist entries of the existing file.

Run in the authoring container only:  python -m oracle.make_golden_round2 [--reuse-mindist]
"""
import os
import pickle
import sys

import numpy as np

from oracle import ref_loader
from oracle.make_golden import Timeout, dubins_problem_args, save, synthetic_swarm_args, with_timeout

sys.setrecursionlimit(20000)

SETUPS = {
    "ComplexObstacles": dict(
        t1=[[8, 9, 10, 11, 12, 13, 12, 11, 10, 9, 8], [8, 10, 12, 14, 20, 14, 12, 10, 10, 9, 8]],
        t2=[[18, 13, 9, 6, 4, 3, 4, 6, 9, 13, 18], [3, 3, 4, 4, 4, 5, 5, 5, 7, 8, 3]], final=(15, 15)),
    "DrivingOnATrack": dict(
        t1=[[0, 0, 0, 3, 4, 5, 6, 7, 10, 10, 10], [0, 3, 4, 5, 6, 6, 6, 6, 7, 8, 10]],
        t2=[[4, 4, 4, 7, 8, 9, 10, 11, 14, 14, 14], [0, 3, 4, 4, 4, 5, 5, 5, 7, 8, 10]], final=(12, 9)),
}


def main(argv=()):
    ref = ref_loader.load()
    bez = ref.bezier
    out = {}
    reuse = None
    if "--reuse-mindist" in argv:
        reuse = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                     "round2.npz"))
    for name, s in SETUPS.items():
        tracks = [bez.Bezier(s["t1"]), bez.Bezier(s["t2"])]
        b = ref.optimization.BezOptimization(
            numVeh=1, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=0.5, maxSpeed=5, maxAngRate=0.5,
            initPoints=(2, 1), finalPoints=s["final"], initSpeeds=1, finalSpeeds=1, initAngs=np.pi / 2,
            finalAngs=np.pi / 2, shapeObstacles=tracks)
        x = b.generateGuess()
        x[-1] = 10
        y = b.reshapeVector(x)
        out[name + "_x"], out[name + "_y"] = x, y
        curves = [bez.Bezier(y)] + tracks
        for i in range(3):
            for j in range(i + 1, 3):
                val = np.full(3, np.nan)
                if reuse is not None:
                    out["%s_ref_%d%d" % (name, i, j)] = reuse["%s_ref_%d%d" % (name, i, j)]
                    continue
                try:
                    r = with_timeout(150, curves[i].minDist, curves[j])
                    if r[0] >= 0:
                        val = np.array([float(v) for v in r])
                    print(name, i, j, r)
                except BaseException as e:        # Timeout lands inside numba dispatch as SystemError
                    print(name, i, j, "no answer:", type(e).__name__)
                out["%s_ref_%d%d" % (name, i, j)] = val

    # pickle with NaN vehicles
    p = os.path.join(ref_loader.REFERENCE_ROOT, "Examples",
                     "SequentialSwarmLONG_MinDistBetweenPtsCost_1-5-20.pickle")
    with open(p, "rb") as f:
        traj = np.asarray(pickle.load(f), dtype=float)
    y = traj[:125 * 3]
    out["seq_nan_y"] = y
    ref.optimization.DEG_ELEV = 10
    try:
        full = np.asarray(ref.optimization._temporalSeparationConstraints(y, 125, 3, 0.9)).reshape(-1, 17)
        base = np.asarray(ref.optimization._temporalSeparationConstraints(y[:121 * 3], 121, 3, 0.9)).reshape(-1, 17)
    finally:
        ref.optimization.DEG_ELEV = 0
    # stored compactly: which rows the reference returns as NaN (all 17 values), after checking
    # that its finite rows are bit-identical to the run without the NaN vehicles (= seq_sep_E10)
    rownan = np.isnan(full).all(axis=1)
    assert np.array_equal(rownan, np.isnan(full).any(axis=1))
    iu, ju = np.triu_indices(125, 1)
    assert np.array_equal(rownan, ju >= 121)
    assert np.array_equal(full[~rownan], base)
    out["seq_nan_rows"] = rownan

    # generateGuess draws
    args, _ = synthetic_swarm_args(5, deg=6, seed=3)
    b = ref.optimization.BezOptimization(**args)
    out["guess_swarm_std0"] = b.generateGuess()
    out["guess_swarm_std07_seed5"] = b.generateGuess(std=0.7, seed=5)
    b = ref.optimization.BezOptimization(**dubins_problem_args(3))
    out["guess_dubins_std05_seed3"] = b.generateGuess(std=0.5, seed=3)
    b = ref.optimization.BezOptimization(
        numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1, maxSpeed=5, maxAngRate=1,
        initPoints=[(0, 5), (3, 0)], finalPoints=[(8, 4), (7, 10)], initSpeeds=[1, 1], finalSpeeds=[1, 1],
        initAngs=[0, np.pi / 2], finalAngs=[0, np.pi / 2], pointObstacles=[[3, 2], [6, 7]])
    out["guess_ex1_std1_seed9"] = b.generateGuess(std=1.0, seed=9)
    # add / sub across different time windows
    rng = np.random.default_rng(31)
    windows = [((0.0, 1.0), (0.25, 1.5)), ((0.5, 2.0), (0.0, 1.25)), ((0.0, 3.0), (1.0, 2.0)),
               ((1.0, 2.0), (0.0, 3.0)), ((0.0, 1.0), (0.0, 1.0)), ((0.0, 1.0), (0.0, 0.5))]
    for k, ((a0, a1), (b0, b1)) in enumerate(windows):
        dim, deg = (2, 4) if k % 2 == 0 else (3, 6)
        ca, cb = rng.normal(size=(dim, deg + 1)), rng.normal(size=(dim, deg + 1))
        A, Bc = bez.Bezier(ca.copy(), t0=a0, tf=a1), bez.Bezier(cb.copy(), t0=b0, tf=b1)
        out["align%d_a" % k], out["align%d_b" % k] = ca, cb
        out["align%d_win" % k] = np.array([a0, a1, b0, b1])
        for opname, res in (("add", A + Bc), ("sub", A - Bc)):
            out["align%d_%s" % (k, opname)] = np.asarray(res.cpts, dtype=float)
            out["align%d_%s_win" % (k, opname)] = np.array([res.t0, res.tf], dtype=float)
    out["nalign"] = np.array(len(windows))

    # objective gradients exactly as SLSQP forms them from the reference's callables
    # (scipy/optimize/_slsqp_py.py:424-426: approx_derivative(fun, x, '2-point', abs_step=eps))
    from scipy.optimize._numdiff import approx_derivative
    args, x = synthetic_swarm_args(5, deg=6, seed=3)
    for goal in ("Euclidean", "Accel"):
        a = dict(args)
        a["minimizeGoal"] = goal
        b = ref.optimization.BezOptimization(**a)
        out["obj_%s_grad" % goal] = approx_derivative(b.objectiveFunction, x, method='2-point',
                                                      abs_step=1.4901161193847656e-08)
    ref.optimization.DEG_ELEV = 7
    try:
        a = dict(args)
        a["minimizeGoal"] = "Accel"
        b = ref.optimization.BezOptimization(**a)
        out["obj_Accel_E7"] = np.array(b.objectiveFunction(x))
        out["obj_Accel_E7_grad"] = approx_derivative(b.objectiveFunction, x, method='2-point',
                                                     abs_step=1.4901161193847656e-08)
    finally:
        ref.optimization.DEG_ELEV = 0
    save("round2", **out)


if __name__ == "__main__":
    main(sys.argv[1:])
