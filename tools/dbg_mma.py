import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import bezier_oracle as O
from oracle.make_golden import synthetic_swarm_args
from optimalbeziertrajectorygeneration_b200 import optimization as gopt
np.set_printoptions(linewidth=250, precision=2)
N, E = 16, 100
args, x = synthetic_swarm_args(N)
gopt.DEG_ELEV = E
b = gopt.BezOptimization(**args)
sep = b.temporalSeparationConstraints(x).reshape(-1, 121)
f = O.make_callables(O.Model(**args), E)
want = f['sep'](x).reshape(-1, 121)
err = np.abs(sep - want) / np.abs(want).max()
print("per-item max err (first 40):", err.max(axis=1)[:40])
print("per-column max err:", err.max(axis=0))
print("ratio row0:", (sep[0] / want[0])[:70])
print("got row0", sep[0][:12]); print("want row0", want[0][:12])
best = [int(np.argmin(np.abs(want - sep[i]).max(axis=1))) for i in range(sep.shape[0])]
print("best matching want row per got row:", best)
print("residual of best:", [float(np.abs(want[best[i]] - sep[i]).max() / np.abs(want).max()) for i in range(0, 40)])
