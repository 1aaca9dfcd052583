"""Development check: the angular-rate kernels (generation 4 = DMMA, default; BEZGPU_ANGRATE_GEN=3 / 2
= the warp-per-item DFMA kernels; BEZGPU_ANGRATE_V1=1 = first generation) against each other and
the numpy oracle, over several (degree, elevation) pairs incl. every block count of generation 4."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import bezier_oracle as O
from oracle.make_golden import dubins_problem_args
from optimalbeziertrajectorygeneration_b200 import optimization as gopt

worst = 0.0
for deg, E in ((10, 100), (10, 0), (10, 30), (5, 7), (3, 1), (12, 115), (7, 60), (16, 111), (4, 0), (6, 1),
               (10, 6), (10, 13), (10, 22), (10, 38), (10, 45), (10, 53), (10, 62), (10, 70), (10, 77), (10, 86),
               (10, 93), (10, 101), (10, 102), (10, 110), (16, 120)):
    args = dubins_problem_args(1, nobs=3, deg=deg)
    b = gopt.BezOptimization(**args)
    x = b.generateGuess(std=0.3, seed=deg)
    gopt.DEG_ELEV = E
    os.environ["BEZGPU_ANGRATE_V1"] = "0"
    new = b.maxAngularRateConstraints(x)
    os.environ["BEZGPU_ANGRATE_GEN"] = "3"
    g3 = b.maxAngularRateConstraints(x)
    os.environ.pop("BEZGPU_ANGRATE_GEN")
    os.environ["BEZGPU_ANGRATE_V1"] = "1"
    old = b.maxAngularRateConstraints(x)
    want = O.make_callables(O.Model(**args), E)["angrate"](x)
    e1 = np.abs(new - want).max() / np.abs(want).max()
    e2 = np.abs(new - old).max() / np.abs(want).max()
    print("deg %2d E %3d m %3d: vs oracle %.2e  vs v1 %.2e  (v1 vs oracle %.2e, gen 3 vs oracle %.2e)" %
          (deg, E, deg + E, e1, e2, np.abs(old - want).max() / np.abs(want).max(),
           np.abs(g3 - want).max() / np.abs(want).max()))
    worst = max(worst, e1)
gopt.DEG_ELEV = 0
os.environ["BEZGPU_ANGRATE_V1"] = "0"
assert worst < 1e-9, worst
print("check_angrate OK, worst %.2e" % worst)
