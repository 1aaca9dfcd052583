"""Development check: tensor-path (DMMA) fused kernels vs the numpy oracle on
degree-10 / elev-100 swarms of several sizes (full tiles, ragged tails, odd pair
counts -> unaligned bulk-store fallback), separation + per-pair min + speed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import bezier_oracle as O
from oracle.make_golden import synthetic_swarm_args
from optimalbeziertrajectorygeneration_b200 import optimization as gopt

worst = 0.0
for N, E in ((16, 100), (3, 100), (11, 100), (23, 50), (40, 107), (9, 44)):
    args, x = synthetic_swarm_args(N)
    gopt.DEG_ELEV = E
    b = gopt.BezOptimization(**args)
    sep = b.temporalSeparationConstraints(x)
    spd = b.maxSpeedConstraints(x)
    red = b.evaluate_reduced(np.stack([x, x + 1e-3]))
    f = O.make_callables(O.Model(**args), E)
    want = f['sep'](x)
    e1 = np.abs(sep - want).max() / np.abs(want).max()
    ws = f['maxspeed'](x)
    e2 = np.abs(spd - ws).max() / np.abs(ws).max()
    L = 2 * 10 + E + 1
    pm = red["pairmin"][0]
    e3 = np.abs(pm - want.reshape(-1, L).min(axis=1)).max() / np.abs(want).max()
    print("N=%3d E=%3d L=%3d: sep %.2e  speed %.2e  pairmin %.2e" % (N, E, L, e1, e2, e3))
    worst = max(worst, e1, e2, e3)
gopt.DEG_ELEV = 0
assert worst < 1e-9, worst
print("check_mma OK, worst %.2e" % worst)
