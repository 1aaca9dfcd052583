"""Runs only the fused pair kernel on the C4 workload (for ncu / quick timing).
usage: python tools/prof_pair.py [B] [reps] [withmin]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import WORKLOAD, fd_batch, synthetic_swarm
from optimalbeziertrajectorygeneration_b200 import optimization as gopt
from optimalbeziertrajectorygeneration_b200.engine import num_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
withmin = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
N, deg, E = WORKLOAD["N"], WORKLOAD["deg"], WORKLOAD["elev"]
args, x = synthetic_swarm(N, deg)
bezopt = gopt.BezOptimization(**args)
eng = bezopt._engine(True)
P, L = num_pairs(N), 2 * deg + E + 1
d_x = eng.upload(fd_batch(x, B))
out = torch.empty((B, P, L), dtype=torch.float64, device=eng.device)
pm = torch.empty((B, P), dtype=torch.float64, device=eng.device) if withmin else None
cpts, tf = eng.assemble(d_x, E)
for _ in range(3):
    eng.separation(cpts, E, args["maxSep"], out=out, pairmin=pm)
torch.cuda.synchronize()
evs = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.separation(cpts, E, args["maxSep"], out=out, pairmin=pm)
    b.record()
    evs.append((a, b))
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in evs])
gb = 8.0 * B * P * L / 1e9
print("pair kernel B=%d withmin=%s: mean %.3f ms  min %.3f ms  -> %.0f GB/s (min: %.0f GB/s)" %
      (B, withmin, ms.mean(), ms.min(), gb / (ms.mean() * 1e-3), gb / (ms.min() * 1e-3)))
