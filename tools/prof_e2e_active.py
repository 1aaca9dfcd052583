"""Where evaluate_sweep_active spends its time (C4 workload): variants of the sweep against the
bare device loop.  Usage: python tools/prof_e2e_active.py [B] [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from optimalbeziertrajectorygeneration_b200 import optimization as gopt   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nS = int(sys.argv[2]) if len(sys.argv) > 2 else 40
W = bench.WORKLOAD
N, deg, E = W["N"], W["deg"], W["elev"]
args, x = bench.synthetic_swarm(N, deg)
X = bench.fd_batch(x, B)
Xs = np.concatenate([X] * nS, axis=0)


def run(label, **kw):
    bezopt = gopt.BezOptimization(**args)
    for k, v in kw.pop("attrs", {}).items():
        setattr(bezopt, k, v)
    for _ in range(2):
        bezopt.evaluate_sweep_active(Xs, elev=E, chunk=B, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bezopt.evaluate_sweep_active(Xs, elev=E, chunk=B, **kw)
    dt = time.perf_counter() - t0
    print("%-46s %.4f ms/step  %.0f evals/s" % (label, dt / nS * 1e3, nS * B / dt), flush=True)
    del bezopt
    torch.cuda.empty_cache()


run("default (graphs, 2 streams, threshold 0)")
run("no CUDA graphs", attrs={"sweep_cuda_graphs": False})
run("threshold -inf (no active pairs, empty list)", threshold=-1e300)
run("rows=False (minima only kernel)", rows=False)
