"""Pins the plain-C oracle (oracle/bezier_oracle.c) against the golden vectors of
the unmodified reference and against the numpy oracle.  CPU only."""
import numpy as np

from conftest import relerr
from oracle import bezier_oracle as O
from oracle import c_oracle as C
from oracle.make_golden import synthetic_swarm_args

TOL = 1e-12


def test_c_oracle_c4_like(golden):
    g = golden("constraints")
    for N in (2, 16, 33):
        args, x = synthetic_swarm_args(N)
        m = O.Model(**args)
        y = O.reshape_vector(m, x)
        sep = C.temporal_separation(y, N, 3, m.maxSep, 100)
        assert relerr(sep, g["c4_N%d_sep_E100" % N]) < TOL
        spd = C.speed(y, N, 3, m.tf, 100, -1.0, m.maxSpeed ** 2)
        assert relerr(spd, g["c4_N%d_maxspeed_E100" % N]) < TOL
        # a split pair range equals the full list; threads do not change results
        P = N * (N - 1) // 2
        if P > 3:
            a = C.temporal_separation(y, N, 3, m.maxSep, 100, 0, 3, nthreads=1)
            b = C.temporal_separation(y, N, 3, m.maxSep, 100, 3, P - 3, nthreads=2)
            assert np.array_equal(np.concatenate([a, b]), sep)


def test_c_oracle_swarm_and_pickle(golden):
    g = golden("constraints")
    m = O.Model(numVeh=36, dimension=3, degree=5, minimizeGoal='Euclidean', maxSep=0.9,
                initPoints=g["swarm_initPts"], finalPoints=g["swarm_finalPts"])
    for E in (0, 10, 100):
        y = O.reshape_vector(m, g["swarm_x1"])
        assert relerr(C.temporal_separation(y, 36, 3, 0.9, E), g["swarm_x1_sep_E%d" % E]) < TOL
    assert relerr(C.temporal_separation(g["seq_y"], 121, 3, 0.9, 10), g["seq_sep_E10"]) < TOL
