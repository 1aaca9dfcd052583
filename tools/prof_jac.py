"""Times the closed-form FD Jacobian sweep of the C4 separation block (bench.py key `jacobian_sweep`)
for a list of BEZGPU_MMA_FLAGS values and checks that they agree bit for bit.
usage: AB_FLAGS=0,128 python tools/prof_jac.py [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import WORKLOAD, synthetic_swarm
from optimalbeziertrajectorygeneration_b200 import optimization as gopt

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N, deg, E = WORKLOAD["N"], WORKLOAD["deg"], WORKLOAD["elev"]
args, x = synthetic_swarm(N, deg)
bezopt = gopt.BezOptimization(**args)
eng = bezopt._engine(True)
flags = os.environ.get("AB_FLAGS", "0,128").split(",")
os.environ["BEZGPU_MMA_FLAGS"] = flags[0]
J = eng.jac_separation(x, E, dense=False)
torch.cuda.synchronize()
ref = J[:64].clone()
gb = J.numel() * 8 / 1e9
for rnd in range(2):
    for fl in flags:
        os.environ["BEZGPU_MMA_FLAGS"] = fl
        time.sleep(0.5)
        eng.jac_separation(x, E, dense=False, out=J)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.jac_separation(x, E, dense=False, out=J)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        ms = np.array(ms)
        print("flags=%s  %s ms  -> %.0f GB/s = %.3f of 6484.6   identical to flags=%s: %s" %
              (fl, " ".join("%.3f" % v for v in ms), gb / (ms.mean() * 1e-3), gb / (ms.mean() * 1e-3) / 6484.6,
               flags[0], bool(torch.equal(J[:64], ref))), flush=True)
