"""Where the end-to-end sweep spends its time (C4 workload): host enqueue time of
BezOptimization.evaluate_sweep vs its wall time, for several sweep lengths, and the bare
PCIe copy times of one step's results.  Usage: python tools/prof_e2e.py [B]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                    # noqa: E402
from optimalbeziertrajectorygeneration_b200 import optimization as gopt   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
W = bench.WORKLOAD
N, deg, E = W["N"], W["deg"], W["elev"]
args, x = bench.synthetic_swarm(N, deg)
X = bench.fd_batch(x, B)
bezopt = gopt.BezOptimization(**args)
gopt.DEG_ELEV = E
eng = bezopt._engine(True)

enq = {}
_orig = torch.cuda.Stream.synchronize


def _sync(self):
    enq.setdefault("t", time.perf_counter())
    return _orig(self)


torch.cuda.Stream.synchronize = _sync

for nS in (24, 96, 384):
    Xs = np.concatenate([X] * nS, axis=0)
    for _ in range(2):
        bezopt.evaluate_sweep(Xs, elev=E, chunk=B)
    torch.cuda.synchronize()
    enq.clear()
    t0 = time.perf_counter()
    bezopt.evaluate_sweep(Xs, elev=E, chunk=B)
    t1 = time.perf_counter()
    print("steps %4d: wall %.3f ms/step, host enqueue %.3f ms/step -> %.0f evals/s"
          % (nS, (t1 - t0) / nS * 1e3, (enq["t"] - t0) / nS * 1e3, nS * B / (t1 - t0)), flush=True)

# bare copies of one step's results
P = N * (N - 1) // 2
L = 2 * deg + E + 1
d_pm = torch.empty((B, P), dtype=torch.float64, device="cuda")
d_sp = torch.empty((B, N * L), dtype=torch.float64, device="cuda")
h_pm = torch.empty((B, P), dtype=torch.float64).pin_memory()
h_sp = torch.empty((B, N * L), dtype=torch.float64).pin_memory()
for name, d, h in (("pairmin", d_pm, h_pm), ("maxspeed", d_sp, h_sp)):
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        h.copy_(d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("D2H %s: %.1f MB in %.3f ms = %.1f GB/s" % (name, d.numel() * 8 / 1e6, ms, d.numel() * 8 / ms / 1e6))
