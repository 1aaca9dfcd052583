"""Kernel timeline (CUPTI via torch.profiler) of evaluate_sweep_active on the C4 workload: which kernel runs when,
on which stream -- to see what happens at the boundaries between the persistent pair kernels.
usage: python tools/timeline_e2e.py [chunks]"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from optimalbeziertrajectorygeneration_b200 import optimization as gopt   # noqa: E402

nS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = 4
W = bench.WORKLOAD
N, deg, E = W["N"], W["deg"], W["elev"]
args, x = bench.synthetic_swarm(N, deg)
Xs = np.concatenate([bench.fd_batch(x, B)] * nS, axis=0)
bezopt = gopt.BezOptimization(**args)
for _ in range(2):
    bezopt.evaluate_sweep_active(Xs, elev=E, chunk=B)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    bezopt.evaluate_sweep_active(Xs, elev=E, chunk=B)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    s = e.time_range.start - t0
    if 3 * 360 < s < 6 * 360:                                  # a few steps from the middle
        print("%9.1f us  +%7.1f us  stream %3s  %s" % (s, e.time_range.end - e.time_range.start,
                                                       getattr(e, "stream", "?"), e.name[:70]))
