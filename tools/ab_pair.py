"""A/B timing of the fused pair kernel on the C4 workload: kernel flags (BEZGPU_MMA_FLAGS:
1 = L1 prefetch of the next tile's rows, 2 = round-1 strided tile order, 4 = partner rows by per-lane
global loads instead of the TMA row fetch (development builds only), 8 = proxy fence + bulk store right
behind each m-tile's epilogue instead of deferred) x output variants.
usage: python tools/ab_pair.py [B] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bench import WORKLOAD, fd_batch, synthetic_swarm
from optimalbeziertrajectorygeneration_b200 import optimization as gopt
from optimalbeziertrajectorygeneration_b200.engine import ActiveSet, num_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N, deg, E = WORKLOAD["N"], WORKLOAD["deg"], WORKLOAD["elev"]
args, x = synthetic_swarm(N, deg)
bezopt = gopt.BezOptimization(**args)
eng = bezopt._engine(True)
P, L = num_pairs(N), 2 * deg + E + 1
d_x = eng.upload(fd_batch(x, B))
out = torch.empty((B, P, L), dtype=torch.float64, device=eng.device)
pm = torch.empty((B, P), dtype=torch.float64, device=eng.device)
act = ActiveSet(B * P, capacity=1 << 17, device=eng.device)
cpts, tf = eng.assemble(d_x, E)
gb = 8.0 * B * (P * L + P + 34 * N) / 1e9


def timeit(fn):
    if not os.environ.get("AB_HOT"):
        time.sleep(0.4)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in evs])


variants = {
    "rows+min": lambda: eng.separation(cpts, E, args["maxSep"], out=out, pairmin=pm),
    "rows only": lambda: eng.separation(cpts, E, args["maxSep"], out=out),
    "rows+min+mask+list": lambda: (act.reset(), eng.separation(cpts, E, args["maxSep"], out=out, pairmin=pm, active=act)),
    "min only (no rows)": lambda: eng.separation(cpts, E, args["maxSep"], pairmin=pm, rows=False),
}
# Under sustained fp64 load a B200 slows down (2-3 % over the first seconds, 15 % after 2 s of
# back-to-back launches: rows+min 0.36 -> 0.42 ms), and a kernel timed right behind a different one
# inherits its dirty L2 lines.  bench.py times short bursts (20 steps = 8 ms), so the comparison is
# made in that regime: every (flags, variant) block = idle pause, 3 warm-up launches, `reps` identical
# launches; AB_ROUNDS cycles through the combinations (their spread is the noise floor).
# AB_HOT=1 measures the sustained regime instead.
flag_list = os.environ.get("AB_FLAGS", "0,2,1").split(",")
only = os.environ.get("AB_VARIANTS")
names = [n for n in variants if not only or n in only.split(",")]
combos = [(fl, name) for fl in flag_list for name in names]
if os.environ.get("AB_HOT"):                  # sustained-load regime: ~2 s of back-to-back launches first
    os.environ["BEZGPU_MMA_FLAGS"] = flag_list[0]
    for _ in range(5000):
        variants[names[0]]()
    torch.cuda.synchronize()
rounds = int(os.environ.get("AB_ROUNDS", "3"))
res = {c: [] for c in combos}
for r in range(rounds):
    for fl, name in combos:
        os.environ["BEZGPU_MMA_FLAGS"] = fl
        res[(fl, name)].append(timeit(variants[name]).mean())
for fl, name in combos:
    ms = np.array(res[(fl, name)])
    print("flags=%s %-20s %s  mean %.4f ms -> %.0f GB/s = %.3f of 6484.6" %
          (fl, name, " ".join("%.4f" % v for v in ms), ms.mean(), gb / (ms.mean() * 1e-3), gb / (ms.mean() * 1e-3) / 6484.6), flush=True)
os.environ["BEZGPU_MMA_FLAGS"] = "0"
eng.separation(cpts, E, args["maxSep"], out=out, pairmin=pm)
print("min == row min:", bool(torch.equal(pm, out.min(dim=2).values)), " active pairs:", int((pm < 0).sum()))
# every flag set must reproduce the flags=0 bits (values, minima, mask, list)
ref_out, ref_pm = out[:1].clone(), pm.clone()
for flags in os.environ.get("AB_FLAGS", "0,2,1").split(","):
    os.environ["BEZGPU_MMA_FLAGS"] = flags
    out.fill_(float("nan")); pm.fill_(float("nan")); act.reset()
    eng.separation(cpts, E, args["maxSep"], out=out, pairmin=pm, active=act)
    torch.cuda.synchronize()
    fl, idx, val, over = ActiveSet.decode(act.buf.cpu().numpy(), B * P, act.capacity)
    pmh = ref_pm.cpu().numpy().ravel()
    ok = (torch.equal(out[:1], ref_out) and torch.equal(pm, ref_pm) and torch.equal(out.min(dim=2).values, ref_pm)
          and np.array_equal(fl, pmh < 0) and np.array_equal(val, pmh[pmh < 0]) and not over)
    pm2 = torch.full_like(pm, float("nan"))
    eng.separation(cpts, E, args["maxSep"], pairmin=pm2, rows=False)
    print("flags=%s bit-identical to flags=0: %s   minima-only identical: %s" % (flags, ok, bool(torch.equal(pm2, ref_pm))))
