"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the
UNMODIFIED reference (imported from /root/reference through ref_loader).

Run in the authoring container:   python -m oracle.make_golden [section ...]
The fixtures are committed; the GPU box never has /root/reference.

Sections
  algebra      random curves -> normSquare / elev / diff / mul / split
  constraints  the BezOptimization closures at the BASELINE configs' x0 (+noise)
  jacobian     scipy approx_derivative ('2-point', SLSQP's abs_step) over them
  gjk          gjkNew on the demo polygons + random hulls           (wave 2)
  mindist      minDist / minDist2Poly / collCheck known answers     (wave 2)
  extrema      Bezier.min / max on depth<=1 inputs                  (wave 2)
"""
import os
import pickle
import signal
import sys

import numpy as np

from oracle import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def save(name, **arrays):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, "(%d arrays, %d bytes)" % (len(arrays), os.path.getsize(path)))


class Timeout(Exception):
    pass


def with_timeout(seconds, fn, *a, **k):
    def handler(signum, frame):
        raise Timeout()
    old = signal.signal(signal.SIGALRM, handler)
    signal.alarm(seconds)
    try:
        return fn(*a, **k)
    finally:
        signal.alarm(0)
        signal.signal(signal.SIGALRM, old)


# ---------------------------------------------------------------------------
def section_algebra(ref):
    bez = ref.bezier
    rng = np.random.default_rng(1)
    out = {}
    cases = []
    for dim in (1, 2, 3):
        for deg in (1, 2, 3, 5, 8, 10):
            cases.append((dim, deg))
    for ci, (dim, deg) in enumerate(cases):
        c = rng.normal(size=(dim, deg + 1)) * 3.0
        tf = float(rng.uniform(0.5, 20.0))
        b = bez.Bezier(c.copy(), tf=tf)
        k = "c%02d_" % ci
        out[k + "cpts"] = c
        out[k + "tf"] = np.array(tf)
        out[k + "normsq"] = np.asarray(b.normSquare().cpts)
        for R in (0, 1, 7, 30):
            out[k + "elev%d" % R] = np.asarray(b.elev(R).cpts)
        out[k + "diff"] = np.asarray(b.diff().cpts)
        out[k + "diff2"] = np.asarray(b.diff().diff().cpts)
        other = rng.normal(size=(dim, deg + 1))
        out[k + "other"] = other
        out[k + "mul"] = np.asarray((b * bez.Bezier(other.copy(), tf=tf)).cpts)
        out[k + "sub"] = np.asarray((b - bez.Bezier(other.copy(), tf=tf)).cpts)
        out[k + "add"] = np.asarray((b + bez.Bezier(other.copy(), tf=tf)).cpts)
        tdiv = float(rng.uniform(0.05, 0.95)) * tf
        l, r = b.split(tdiv)
        out[k + "tdiv"] = np.array(tdiv)
        out[k + "split_l"] = np.asarray(l.cpts)
        out[k + "split_r"] = np.asarray(r.cpts)
        tau = np.linspace(0, tf, 17)
        out[k + "tau"] = tau
        out[k + "eval"] = np.asarray(b(tau))
        out[k + "integrate"] = np.asarray(b.integrate())
    out["ncases"] = np.array(len(cases))
    # the two C4 tables, straight from the reference builders (Q14)
    out["elevMatrix_20_100"] = bez.elevMatrix(20, 100)
    out["elevMatrix_10_3"] = bez.elevMatrix(10, 3)
    out["prodMatrix_10"] = bez.prodMatrix(10)
    out["bezProductCoefficients_7"] = bez.bezProductCoefficients(7)
    save("algebra", **out)


# ---------------------------------------------------------------------------
def _example1_problem(ref):
    numVeh, dim, deg = 2, 2, 10
    return ref.optimization.BezOptimization(
        numVeh=numVeh, dimension=dim, degree=deg, minimizeGoal='TimeOpt',
        maxSep=1, maxSpeed=5, maxAngRate=1,
        initPoints=[(0, 5), (3, 0)], finalPoints=[(8, 4), (7, 10)],
        initSpeeds=[1] * numVeh, finalSpeeds=[1] * numVeh,
        initAngs=[0, np.pi / 2], finalAngs=[0, np.pi / 2],
        pointObstacles=[[3, 2], [6, 7]])


def _swarm_problem(ref):
    ex = ref_loader.load_example("SwarmOfAerialVehicles")
    numVeh, initPts, finalPts = ex.generatePointsFromImage(ex.CAS_IMG)
    bezopt = ref.optimization.BezOptimization(
        numVeh=numVeh, dimension=3, degree=5, minimizeGoal='Euclidean',
        maxSep=0.9, initPoints=initPts, finalPoints=finalPts)
    x0 = ex.generate3DGuess(initPts, finalPts, 5)
    return bezopt, x0, initPts, finalPts


def dubins_problem_args(seed, nobs=16, deg=10):
    """C5-style problem (Examples/DubinsCarExample2.py:81-103)."""
    rng = np.random.default_rng(seed)
    obs = rng.uniform(1.0, 11.0, size=(nobs, 2))
    return dict(numVeh=1, dimension=2, degree=deg, minimizeGoal='TimeOpt',
                maxSep=1, maxSpeed=3, maxAngRate=np.pi / 2,
                initPoints=[(0, 0)], finalPoints=[(12, 8)],
                initSpeeds=[1], finalSpeeds=[1], tf=8,
                initAngs=[np.pi / 2], finalAngs=[0],
                pointObstacles=[list(o) for o in obs])


def synthetic_swarm_args(N, deg=10, seed=20261018):
    """C4-style problem: N vehicles, 3-D, straight-line guess +
    N(0,1) noise; end points uniform in [0,100]^3."""
    rng = np.random.default_rng(seed)
    init = rng.uniform(0, 100, size=(N, 3))
    final = rng.uniform(0, 100, size=(N, 3))
    args = dict(numVeh=N, dimension=3, degree=deg, minimizeGoal='Euclidean',
                maxSep=0.9, maxSpeed=5, tf=20.0, initPoints=init, finalPoints=final)
    x = np.empty((N * 3, deg - 1))
    for i in range(N):
        for d in range(3):
            x[3 * i + d] = np.linspace(init[i, d], final[i, d], deg + 1)[1:-1]
    x = x + rng.normal(size=x.shape)
    return args, x.ravel()


def _eval_all(ref, bezopt, x, E, with_ang):
    ref.optimization.DEG_ELEV = E
    try:
        d = {"sep": np.asarray(bezopt.temporalSeparationConstraints(x)),
             "maxspeed": np.asarray(bezopt.maxSpeedConstraints(x)),
             "minspeed": np.asarray(bezopt.minSpeedConstraints(x))}
        if with_ang:
            d["angrate"] = np.asarray(bezopt.maxAngularRateConstraints(x))
    finally:
        ref.optimization.DEG_ELEV = 0
    return d


def section_constraints(ref):
    out = {}
    rng = np.random.default_rng(7)

    # C2 Example1 (Examples/Example1_DubinsCarTimeOptimal.py:95-131)
    bezopt = _example1_problem(ref)
    x0 = bezopt.generateGuess(std=0)
    out["ex1_x0"] = x0
    out["ex1_y0"] = bezopt.reshapeVector(x0)
    ex1 = ref_loader.load_example("Example1_DubinsCarTimeOptimal")
    for xi, x in enumerate([x0, x0 + rng.normal(size=x0.size) * 0.3]):
        out["ex1_x%d" % xi] = x
        for E in (0, 30, 100):
            # the example's own separation closure (vehicles only, elev=E)
            out["ex1_x%d_ownsep_E%d" % (xi, E)] = np.asarray(
                ex1._temporalSeparationConstraints(bezopt.reshapeVector(x), 2, 2, 1, E))
        for E in (0, 10, 100):
            for k, v in _eval_all(ref, bezopt, x, E, with_ang=(E <= 10 or xi == 0)).items():
                out["ex1_x%d_%s_E%d" % (xi, k, E)] = v

    # C3 swarm (Examples/SwarmOfAerialVehicles.py:137-150)
    bezopt, x0, initPts, finalPts = _swarm_problem(ref)
    out["swarm_initPts"] = initPts
    out["swarm_finalPts"] = finalPts
    out["swarm_x0"] = x0
    out["swarm_objective_x0"] = np.array(bezopt.objectiveFunction(x0))
    xs = [x0, x0 + rng.normal(size=x0.size) * 0.5]
    for xi, x in enumerate(xs):
        out["swarm_x%d" % xi] = x
        for E in (0, 10, 100):
            for k, v in _eval_all(ref, bezopt, x, E, with_ang=False).items():
                out["swarm_x%d_%s_E%d" % (xi, k, E)] = v

    # C4-like, small N (synthetic 3-D swarm, deg 10, E=100)
    for N in (2, 16, 33):
        args, x = synthetic_swarm_args(N)
        b = ref.optimization.BezOptimization(**args)
        out["c4_N%d_x" % N] = x
        for k, v in _eval_all(ref, b, x, 100, with_ang=False).items():
            out["c4_N%d_%s_E100" % (N, k)] = v

    # C5-like Dubins + 16 point obstacles, deg 10, E=100 (ang-rate at m=110)
    for seed in (0, 1, 2):
        args = dubins_problem_args(seed)
        b = ref.optimization.BezOptimization(**args)
        x = b.generateGuess(std=0.5, seed=seed)
        out["c5_s%d_x" % seed] = x
        out["c5_s%d_y" % seed] = b.reshapeVector(x)
        for E in ((100, 0, 5) if seed == 0 else (100,)):
            for k, v in _eval_all(ref, b, x, E, with_ang=True).items():
                out["c5_s%d_%s_E%d" % (seed, k, E)] = v

    # sequential-swarm pickle: 121 finite deg-3 3-D trajectories
    p = os.path.join(ref_loader.REFERENCE_ROOT, "Examples",
                     "SequentialSwarmLONG_MinDistBetweenPtsCost_1-5-20.pickle")
    with open(p, "rb") as f:
        traj = np.asarray(pickle.load(f), dtype=float)
    y = traj[:121 * 3]
    out["seq_y"] = y
    ref.optimization.DEG_ELEV = 10
    try:
        out["seq_sep_E10"] = np.asarray(
            ref.optimization._temporalSeparationConstraints(y, 121, 3, 0.9))
    finally:
        ref.optimization.DEG_ELEV = 0

    # objectives
    args, x = synthetic_swarm_args(5, deg=6, seed=3)
    for goal in ("Euclidean", "Accel"):
        a = dict(args)
        a["minimizeGoal"] = goal
        b = ref.optimization.BezOptimization(**a)
        out["obj_%s_x" % goal] = x
        out["obj_%s" % goal] = np.array(b.objectiveFunction(x))
    save("constraints", **out)


# ---------------------------------------------------------------------------
def section_jacobian(ref):
    from scipy.optimize._numdiff import approx_derivative
    eps = 1.4901161193847656e-08           # _slsqp_py.py: _epsilon = sqrt(finfo.eps)
    out = {}

    def jac(fun, x):
        return approx_derivative(fun, x, method='2-point', abs_step=eps)

    bezopt = _example1_problem(ref)
    x0 = bezopt.generateGuess(std=0) + np.random.default_rng(3).normal(size=29) * 0.2
    out["ex1_x"] = x0
    for E in (0, 10):
        ref.optimization.DEG_ELEV = E
        try:
            out["ex1_J_sep_E%d" % E] = jac(bezopt.temporalSeparationConstraints, x0)
            out["ex1_J_maxspeed_E%d" % E] = jac(bezopt.maxSpeedConstraints, x0)
            out["ex1_J_angrate_E%d" % E] = jac(bezopt.maxAngularRateConstraints, x0)
        finally:
            ref.optimization.DEG_ELEV = 0

    args, x = synthetic_swarm_args(6, deg=5, seed=11)
    b = ref.optimization.BezOptimization(**args)
    out["sw6_x"] = x
    ref.optimization.DEG_ELEV = 10
    try:
        out["sw6_J_sep_E10"] = jac(b.temporalSeparationConstraints, x)
        out["sw6_J_maxspeed_E10"] = jac(b.maxSpeedConstraints, x)
    finally:
        ref.optimization.DEG_ELEV = 0

    args = dubins_problem_args(5, nobs=4, deg=6)
    b = ref.optimization.BezOptimization(**args)
    x = b.generateGuess(std=0.5, seed=5)
    out["dub_x"] = x
    out["dub_obs"] = np.asarray(args["pointObstacles"])
    ref.optimization.DEG_ELEV = 8
    try:
        out["dub_J_sep_E8"] = jac(b.temporalSeparationConstraints, x)
        out["dub_J_maxspeed_E8"] = jac(b.maxSpeedConstraints, x)
        out["dub_J_angrate_E8"] = jac(b.maxAngularRateConstraints, x)
    finally:
        ref.optimization.DEG_ELEV = 0
    save("jacobian", **out)


SECTIONS = {"algebra": section_algebra, "constraints": section_constraints,
            "jacobian": section_jacobian}


def main(argv):
    ref = ref_loader.load()
    try:
        from oracle import make_golden_geometry as geo   # wave-2 sections
        SECTIONS.update(geo.SECTIONS)
    except ImportError:
        pass
    names = argv or list(SECTIONS)
    for n in names:
        print("== section", n)
        SECTIONS[n](ref)


if __name__ == "__main__":
    main(sys.argv[1:])
