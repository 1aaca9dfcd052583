"""Throughput of the geometry kernels (per-warp GJK, minDist DFS,
collCheck) on batches of random degree-5 3-D curve pairs (BASELINE configs[0] style), next to
the pure-Python oracle (a restatement of the reference's Python/numba code) on a bounded
sample.  One JSON line; not part of bench.py's contract.
usage: python tools/bench_geometry.py [pairs]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import gjk_oracle as G
from optimalbeziertrajectorygeneration_b200 import bezier as gbez
from optimalbeziertrajectorygeneration_b200.gjk import gjk as ggjk

count = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rng = np.random.default_rng(2026)
A = np.cumsum(rng.normal(size=(count, 3, 6)), axis=2)
B = np.cumsum(rng.normal(size=(count, 3, 6)), axis=2) + rng.normal(size=(count, 3, 1)) * 3
P1 = np.ascontiguousarray(A.transpose(0, 2, 1))          # control polygons [count, 6, 3]
P2 = np.ascontiguousarray(B.transpose(0, 2, 1))


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


t_gjk, (flag, a, b, dist) = timed(lambda: ggjk.gjk_batch(P1, P2))
t_md, (out, status) = timed(lambda: gbez.min_dist_batch(A, B, max_nodes=20000))
t_cc, cc = timed(lambda: gbez.coll_check_batch(A, B))

ns = 48                                                  # CPU sample
t0 = time.perf_counter()
for k in range(ns):
    G.gjk_new(P1[k], P2[k])
t_gjk_cpu = (time.perf_counter() - t0) / ns
t0 = time.perf_counter()
agree = 0
for k in range(ns):
    alpha, t1, t2, st = G.min_dist(A[k], B[k], max_nodes=20000)
    agree += int(st == status[k] and (st != 0 or (alpha, t1, t2) == tuple(out[k])))
t_md_cpu = (time.perf_counter() - t0) / ns
print(json.dumps({
    "workload": "%d random degree-5 3-D curve pairs (host arrays in, host results out)" % count,
    "gjk_pairs_per_s": count / t_gjk, "mindist_pairs_per_s": count / t_md, "collcheck_pairs_per_s": count / t_cc,
    "mindist_ok_fraction": float(np.mean(np.asarray(status) == 0)),
    "cpu_oracle": {"kind": "port (pure-Python restatement of gjk/gjk.py + bezier._minDist)", "cores": 1,
                   "sample_pairs": ns, "gjk_pairs_per_s": 1.0 / t_gjk_cpu, "mindist_pairs_per_s": 1.0 / t_md_cpu,
                   "bit_identical_on_sample": agree == ns}}))
