"""Host-side logic of the drop-in mirror that needs no GPU (the package imports the C-ABI
library, which must be built; no kernel is launched)."""
import numpy as np

from oracle.make_golden import dubins_problem_args, synthetic_swarm_args
from oracle.make_golden_round2 import SETUPS


def test_generate_guess_matches_reference_draws(golden):
    """generateGuess (optimization.py:189-240) was rewritten in this repo's own words: the
    seeded draws must still equal the reference's bit for bit (same RNG order)."""
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    g, gc = golden("round2"), golden("constraints")
    args, _ = synthetic_swarm_args(5, deg=6, seed=3)
    b = gopt.BezOptimization(**args)
    assert np.array_equal(b.generateGuess(), g["guess_swarm_std0"])
    assert np.array_equal(b.generateGuess(std=0.7, seed=5), g["guess_swarm_std07_seed5"])
    b = gopt.BezOptimization(**dubins_problem_args(3))
    assert np.array_equal(b.generateGuess(std=0.5, seed=3), g["guess_dubins_std05_seed3"])
    for seed in (0, 1, 2):                                   # the C5-like golden inputs
        b = gopt.BezOptimization(**dubins_problem_args(seed))
        assert np.array_equal(b.generateGuess(std=0.5, seed=seed), gc["c5_s%d_x" % seed])
    ex1 = gopt.BezOptimization(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1, maxSpeed=5,
                               maxAngRate=1, initPoints=[(0, 5), (3, 0)], finalPoints=[(8, 4), (7, 10)],
                               initSpeeds=[1, 1], finalSpeeds=[1, 1], initAngs=[0, np.pi / 2],
                               finalAngs=[0, np.pi / 2], pointObstacles=[[3, 2], [6, 7]])
    assert np.array_equal(ex1.generateGuess(std=1.0, seed=9), g["guess_ex1_std1_seed9"])
    assert np.array_equal(ex1.generateGuess(), gc["ex1_x0"])
    for name, s in SETUPS.items():
        b = gopt.BezOptimization(numVeh=1, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=0.5,
                                 maxSpeed=5, maxAngRate=0.5, initPoints=(2, 1), finalPoints=s["final"],
                                 initSpeeds=1, finalSpeeds=1, initAngs=np.pi / 2, finalAngs=np.pi / 2)
        x = b.generateGuess()
        x[-1] = 10
        assert np.array_equal(x, g[name + "_x"])
    bad = gopt.BezOptimization(numVeh=1, dimension=3, degree=5, initPoints=[(0, 0, 0)], finalPoints=[(1, 1, 1)],
                               initSpeeds=[1], finalSpeeds=[1], initAngs=[0], finalAngs=[0])
    try:
        bad.generateGuess()
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def test_engine_cache_follows_model_edits():
    """The reference re-reads model / pointObstacles on every call; the device state here is
    rebuilt when an entry is replaced (signature check, no GPU needed for the check itself)."""
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    b = gopt.BezOptimization(numVeh=2, dimension=2, degree=5, initPoints=[(0, 0), (1, 1)],
                             finalPoints=[(2, 2), (3, 3)], pointObstacles=[[1, 2]])
    s0 = b._model_signature()
    b.pointObstacles = [[1, 2], [3, 4]]
    s1 = b._model_signature()
    b.model['tf'] = 7.0
    s2 = b._model_signature()
    b.model['initPoints'] = np.array([[5.0, 5.0], [1.0, 1.0]])
    s3 = b._model_signature()
    assert len({s0, s1, s2, s3}) == 4
