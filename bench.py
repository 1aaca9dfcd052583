#!/usr/bin/env python
"""bench.py -- constraint + Jacobian evals/sec of the Bezier hot path on B200.

Workload (BASELINE.json configs[3], "C4"): synthetic 1024-vehicle 3-D swarm,
degree 10, DEG_ELEV 100: one *eval* = the full constraint vector of one x
(523 776 pairs x 121 separation values + 1024 x 121 max-speed values = 508 MB).
One *step* = one pass over a batch of B such x (the base point and B-1
finite-difference perturbations of it, which is what SLSQP's Jacobian asks for).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU).  Sharding: the batch of
perturbed x is split across ranks (each rank evaluates its own B x over all
pairs, no data-path collective); the per-pair minimum (the active-pair source,
[B, P] fp64) is all-gathered over NCCL/NVLink so every rank holds the whole
min-distance matrix.  Weak scaling: B per rank is fixed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(N=1024, dim=3, deg=10, elev=100)
SEED = 20261018


def synthetic_swarm(N, deg, seed=SEED):
    """SURVEY 8(d) 'Synthetic inputs C4' (same generator as oracle/make_golden.py)."""
    rng = np.random.default_rng(seed)
    init = rng.uniform(0, 100, size=(N, 3))
    final = rng.uniform(0, 100, size=(N, 3))
    args = dict(numVeh=N, dimension=3, degree=deg, minimizeGoal='Euclidean',
                maxSep=0.9, maxSpeed=5, tf=20.0, initPoints=init, finalPoints=final)
    x = np.empty((N * 3, deg - 1))
    for i in range(N):
        for d in range(3):
            x[3 * i + d] = np.linspace(init[i, d], final[i, d], deg + 1)[1:-1]
    x = x + rng.normal(size=x.shape)
    return args, x.ravel()


def fd_batch(x, B, seed=1):
    """base x plus B-1 single-variable forward perturbations (SciPy's h)."""
    X = np.repeat(x[None, :], B, axis=0)
    rng = np.random.default_rng(seed)
    for b, k in enumerate(rng.choice(x.size, size=B - 1, replace=False), start=1):
        X[b, k] += 1.4901161193847656e-08
    return X


# --------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled (NVML) during the timed region."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
               "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index, period=0.002):
        self.gpu, self.period = gpu_index, period
        self.sm, self.reasons, self.smax = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _open(self):
        import pynvml
        pynvml.nvmlInit()
        # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it lists indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.gpu
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            idx = int(vis.split(",")[self.gpu])
        self._nvml = pynvml
        self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))

    def _run(self):
        try:
            pynvml, h = self._nvml, self._h
            while True:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                if self._stop.wait(self.period):
                    break
        except Exception as e:            # pragma: no cover
            self.err = repr(e)

    def open(self):
        """NVML is opened ahead of the timed region (and of the barrier in front of it: it takes
        milliseconds, on rank 0 only), so that short runs are sampled from their first step."""
        try:
            self._open()
        except Exception as e:            # pragma: no cover
            self.err = repr(e)

    def start(self):
        if self.err is None and not hasattr(self, "_h"):
            self.open()
        if self.err is not None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=5)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


# --------------------------------------------------------------------------
def cpu_baseline(args, x, E, target_seconds=12.0, threads=0):
    """Times the plain-C restatement of the reference (oracle/bezier_oracle.c, all
    host threads) on a bounded, contiguous sample of the C4 pair list and
    extrapolates linearly to evals/s (pairs are independent and of equal cost).
    The Python reference itself cannot travel to the GPU box; in the authoring
    container it costs 23-34 us per pair per core, about
    10x more than this port."""
    from oracle import bezier_oracle as O
    from oracle import c_oracle as C
    m = O.Model(**args)
    y = O.reshape_vector(m, x)
    N, dim = m.numVeh, m.dim
    P = N * (N - 1) // 2
    L = 2 * m.deg + E + 1
    threads = threads or C.max_threads()
    cal_pairs = min(P, 65536)
    out = np.empty(cal_pairs * L)
    C.temporal_separation(y, N, dim, m.maxSep, E, 0, cal_pairs, nthreads=threads, out=out)   # warm + calibrate
    t0 = time.perf_counter()
    C.temporal_separation(y, N, dim, m.maxSep, E, 0, cal_pairs, nthreads=threads, out=out)
    per_pair = (time.perf_counter() - t0) / cal_pairs
    chunk = min(P, 262144)
    out = np.empty(chunk * L)
    reps = max(1, int(target_seconds / (per_pair * chunk)))
    done = 0
    t0 = time.perf_counter()
    for r in range(reps):
        begin = (r * chunk) % max(1, P - chunk + 1)
        C.temporal_separation(y, N, dim, m.maxSep, E, begin, chunk, nthreads=threads, out=out)
        done += chunk
    t_speed0 = time.perf_counter()
    C.speed(y, N, dim, m.tf, E, -1.0, m.maxSpeed ** 2, nthreads=threads)
    t_speed = time.perf_counter() - t_speed0
    wall = time.perf_counter() - t0 - t_speed
    sec_per_eval = wall / done * P + t_speed
    return {"value": 1.0 / sec_per_eval, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": "%d pair evaluations drawn from the %d pairs of the C4 swarm (N=%d, deg %d, elev %d) + all %d speed rows, "
                      "%d threads, %.1f s wall; plain-C restatement of the reference "
                      "(oracle/bezier_oracle.c: sub -> normSquare -> elev), extrapolated linearly in pairs"
                      % (done, P, N, m.deg, E, N, threads, wall + t_speed)}, wall + t_speed


# --------------------------------------------------------------------------
def run_reference(opts):
    """Reference arm: the reference's own algorithm for the path (C restatement, all host
    threads) on the same workload; each step is a bounded sample of the C4 pair list, sized so
    that the whole --steps/--warmup run stays within about two minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bezier_oracle as O
    from oracle import c_oracle as C
    args, x = synthetic_swarm(WORKLOAD["N"], WORKLOAD["deg"])
    E = WORKLOAD["elev"]
    m = O.Model(**args)
    y = O.reshape_vector(m, x)
    N, dim = m.numVeh, m.dim
    P = N * (N - 1) // 2
    L = 2 * m.deg + E + 1
    threads = C.max_threads()
    cal = min(P, 65536)
    out = np.empty(cal * L)
    C.temporal_separation(y, N, dim, m.maxSep, E, 0, cal, nthreads=threads, out=out)          # warm
    t0 = time.perf_counter()
    C.temporal_separation(y, N, dim, m.maxSep, E, 0, cal, nthreads=threads, out=out)
    per_pair = (time.perf_counter() - t0) / cal
    t0 = time.perf_counter()
    C.speed(y, N, dim, m.tf, E, -1.0, m.maxSpeed ** 2, nthreads=threads)
    t_speed = time.perf_counter() - t0
    total_steps = max(1, opts.steps + opts.warmup)
    per_step = max(0.2, min(15.0, 110.0 / total_steps))
    chunk = int(min(P, max(4096, per_step / per_pair)))
    out = np.empty(chunk * L)
    vals, walls, done = [], [], 0
    for s in range(total_steps):
        begin = (s * chunk) % max(1, P - chunk + 1)
        t0 = time.perf_counter()
        C.temporal_separation(y, N, dim, m.maxSep, E, begin, chunk, nthreads=threads, out=out)
        wall = time.perf_counter() - t0
        if s >= opts.warmup:
            vals.append(1.0 / (wall / chunk * P + t_speed))
            walls.append(wall)
            done += chunk
    v = float(np.mean(vals))
    cb = {"value": v, "unit": "evals/s", "cores": threads, "kind": "port",
          "sample": "%d steps x %d pair evaluations drawn from the %d pairs of the C4 swarm (N=%d, deg %d, elev %d) "
                    "+ all %d speed rows timed once, %d threads, %.1f s of CPU wall in the timed steps; plain-C "
                    "restatement of the reference (oracle/bezier_oracle.c: sub -> normSquare -> elev), extrapolated "
                    "linearly in pairs" % (len(vals), chunk, P, N, m.deg, E, N, threads, float(np.sum(walls)))}
    line = {"impl": "reference", "metric": "constraint+Jacobian evals/sec", "value": v, "unit": "evals/s",
            "n_gpus": opts.gpus, "steps": opts.steps, "warmup": opts.warmup,
            "ms_per_step": 1e3 * float(np.mean(walls)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(1), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(B, gather_mode="n/a"):
    return {"workload": "C4 synthetic swarm: N=1024 vehicles, dim 3, degree 10, DEG_ELEV 100, "
                        "all 523776 pairs x 121 separation values + 1024 x 121 max-speed values per eval",
            "evals_per_step_per_gpu": B, "sharding": "FD-perturbation batch split across ranks; "
            "all-gather of the [B,P] per-pair minimum", "gather": gather_mode, "l2_policy": "outputs (508 MB/eval) >> 126 MB L2, "
            "streaming stores; inputs 270 KB"}


# --------------------------------------------------------------------------
def run_ours(opts):
    import torch
    import torch.distributed as dist
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    from optimalbeziertrajectorygeneration_b200 import sharding
    from optimalbeziertrajectorygeneration_b200.engine import num_pairs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    N, deg, E = WORKLOAD["N"], WORKLOAD["deg"], WORKLOAD["elev"]
    B = opts.batch
    args, x = synthetic_swarm(N, deg)
    Xall = fd_batch(x, B * world)
    X = Xall[rank * B:(rank + 1) * B]
    bezopt = gopt.BezOptimization(**args)
    eng = bezopt._engine(True)
    P = num_pairs(N)
    L = 2 * deg + E + 1
    d_x = eng.upload(X)
    out_sep = torch.empty((B, P, L), dtype=torch.float64, device=eng.device)
    out_spd = torch.empty((B, N, L), dtype=torch.float64, device=eng.device)
    pairmin = torch.empty((B, P), dtype=torch.float64, device=eng.device)
    max_speed2 = float(args["maxSpeed"]) ** 2

    # The one collective of the path: every rank ends up with the whole [world*B, P] per-pair
    # minimum (active-pair) matrix.  Preferred: fused into the pair kernel (NVLink peer stores
    # into symmetric memory, sharding.PeerMinima); fallback: NCCL all-gather on its own stream.
    gatherer, peer, gather_mode = None, None, "none (1 GPU)"
    if world > 1 and not opts.nccl_gather:
        try:
            peer = sharding.PeerMinima(B, P, eng.device)
            gather_mode = "fused: in-kernel NVLink peer stores into symmetric memory + signal-pad barrier"
        except Exception as e:                      # symmetric memory not available on this box
            if rank == 0:
                print("PeerMinima unavailable (%r); falling back to NCCL all-gather" % (e,), file=sys.stderr)
    if peer is None:
        gatherer = sharding.PairMinimaGatherer(B, P, eng.device)
        if world > 1:
            gather_mode = "NCCL all_gather_into_tensor on a side stream"

    def step():
        cpts, tf = eng.assemble(d_x, E)
        if peer is not None:
            pm, peers = peer.targets()
            eng.separation(cpts, E, args["maxSep"], out=out_sep, pairmin=pm, peer_ptrs=peers)
            eng.speed(cpts, tf, E, -1.0, max_speed2, out=out_spd)
            return peer.complete()
        pm = gatherer.local_buffer()
        eng.separation(cpts, E, args["maxSep"], out=out_sep, pairmin=pm)
        eng.speed(cpts, tf, E, -1.0, max_speed2, out=out_spd)
        return gatherer.gather()
    launches_per_step = 3       # assemble, fused pair kernel (values + per-pair min), speed kernel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, opts.warmup)):
        step()
    if gatherer is not None:
        gatherer.finish()
    if peer is not None:
        peer.wait()
    sampler = ClockSampler(local, period=opts.clock_period)
    if rank == 0:
        sampler.open()
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(opts.steps):
        gathered = step()
    if gatherer is not None:
        gatherer.finish()
    if peer is not None:
        peer.wait()                 # the last step's barrier (side stream) is part of the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # self-check of the collective (outside the timed region): the gathered matrix of the
        # last step must equal a plain NCCL all-gather of the per-rank minima, bit for bit
        mine = gathered[rank * B:(rank + 1) * B].clone()
        ref = torch.empty((world * B, P), dtype=torch.float64, device=eng.device)
        dist.all_gather_into_tensor(ref, mine)
        torch.cuda.synchronize()
        assert torch.equal(ref, gathered), "gathered per-pair minima differ from the NCCL all-gather"
        assert torch.equal(mine, out_sep.min(dim=2).values), "per-pair minima differ from the row minima"
    ms = float(t.item())

    # dominant kernel alone (pair kernel), CUDA events on its stream
    cpts, tf = eng.assemble(d_x, E)
    torch.cuda.synchronize()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(opts.steps)]
    for a, b_ in kev:
        a.record()
        eng.separation(cpts, E, args["maxSep"], out=out_sep, pairmin=pairmin)
        b_.record()
    torch.cuda.synchronize()
    kms = float(np.mean([a.elapsed_time(b_) for a, b_ in kev]))
    clocks = sampler.stop() if rank == 0 else None

    # end to end through the public host API (host X in pinned memory -> H2D ->
    # kernels, every row materialised in HBM -> D2H of the step's result).
    #  (a) evaluate_reduced: result = per-pair minima [B,P] + max-speed rows
    #  (b) the reference-facing closures: result = the full constraint vector
    gopt.DEG_ELEV = E
    bezopt.zero_copy_results = True
    nE = max(3, min(opts.steps, 20))
    for _ in range(3):
        red = bezopt.evaluate_reduced(X, elev=E)
    barrier()
    t0 = time.perf_counter()
    for _ in range(nE):
        red = bezopt.evaluate_reduced(X, elev=E)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_red = float(tt.item())
    sepf, spdf = bezopt.temporalSeparationConstraints, bezopt.maxSpeedConstraints
    nF = 3
    for _ in range(2):
        r1 = sepf(X[0]); r2 = spdf(X[0])
    barrier()
    t0 = time.perf_counter()
    for s_ in range(nF):
        r1 = sepf(X[s_ % B]); r2 = spdf(X[s_ % B])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_full = float(tt.item())
    # (c) the pipelined sweep API: nS chunks of B rows, copies of chunk k overlap kernels of k+1
    nS = max(4, min(opts.steps, 24))
    Xs = np.concatenate([X] * nS, axis=0)
    for _ in range(2):
        sw = bezopt.evaluate_sweep(Xs, elev=E, chunk=B)
    barrier()
    t0 = time.perf_counter()
    sw = bezopt.evaluate_sweep(Xs, elev=E, chunk=B)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_sweep = float(tt.item())
    assert np.array_equal(sw["pairmin"][:B], red["pairmin"]) and np.array_equal(sw["maxspeed"][-B:], red["maxspeed"])
    e2e = {"value": nS * B * world / dt_sweep, "unit": "evals/s",
           "h2d_bytes_per_step": int(X.size * 8),
           "d2h_bytes_per_step": int((red["pairmin"].size + red["maxspeed"].size) * 8),
           "steps_timed": nS,
           "note": "BezOptimization.evaluate_sweep(X_host[steps*B, nvar], chunk=B): per step (chunk of B rows) pinned "
                   "H2D of X -> assemble -> fused pair kernel (all P*L values written to HBM + per-pair min) -> speed "
                   "kernel -> D2H of the per-pair minima and max-speed rows into pinned host memory; two device "
                   "workspaces, the D2H of step k overlaps the kernels of step k+1; wall clock incl. Python",
           "serial_call": {"value": nE * B * world / dt_red, "unit": "evals/s", "steps_timed": nE,
                           "note": "one evaluate_reduced(X_host[B,nvar]) call per step, H2D -> kernels -> D2H "
                                   "strictly in sequence with a host synchronisation per call"},
           "full_vector": {"value": nF * world / dt_full, "unit": "evals/s",
                           "d2h_bytes_per_eval": int((r1.size + r2.size) * 8),
                           "note": "temporalSeparationConstraints(x)+maxSpeedConstraints(x) returning the full "
                                   "508 MB constraint vector to host memory (PCIe-bound)"}}
    gopt.DEG_ELEV = 0

    # Sparse-aware FD sweep: the whole Jacobian of the separation block of ONE x
    # (what SLSQP obtains from nvar+1 = 27 649 full evals) in closed form, sweep layout
    # [variable][partner curve][L] = only the rows that depend on the variable (27.4 GB).
    sweep = None
    if not opts.no_sweep:
        torch.cuda.empty_cache()
        J = eng.jac_separation(x, E, dense=False)
        for _ in range(2):
            eng.jac_separation(x, E, dense=False, out=J)
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(3):
            eng.jac_separation(x, E, dense=False, out=J)
        b_.record()
        torch.cuda.synchronize()
        sms = a_.elapsed_time(b_) / 3
        sweep = {"value": (bezopt.nvar + 1) * world / (sms * 1e-3), "unit": "evals/s",
                 "ms_per_jacobian": sms, "bytes_per_jacobian": int(J.numel() * 8),
                 "note": "closed-form FD Jacobian of the separation block of one x (all %d variables x %d partner "
                         "curves x %d values), equivalent to nvar+1 full evals of the reference; every rank "
                         "computes the Jacobian of its own x" % (bezopt.nvar, N - 1, L)}
        del J
        torch.cuda.empty_cache()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = json.load(open(peaks_path))["hbm_gbs"]
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        alg_bytes = 8.0 * B * (P * L + P + 34 * N)      # rows + per-pair minima written, control-point rows read
        achieved = alg_bytes / (kms * 1e-3) / 1e9
        # measured DRAM traffic of the same launch shape from the committed ncu --set full capture
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("evals_per_launch") == B:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        evals = B * world * opts.steps
        line = {"metric": "constraint+Jacobian evals/sec", "value": evals / (ms * 1e-3), "unit": "evals/s",
                "n_gpus": world, "steps": opts.steps, "warmup": max(3, opts.warmup),
                "ms_per_step": ms / opts.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(B, gather_mode),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "kernel": "sq_elev_mma_kernel<10,3,PAIR,min> (DMMA.8x8x4 stage 2, TMA bulk-store epilogue)",
                             "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
                "e2e": e2e, "gpu_launches": launches_per_step * opts.steps, "clocks": clocks}
        if sweep is not None:
            sweep["hbm_frac"] = sweep["bytes_per_jacobian"] / (sweep["ms_per_jacobian"] * 1e-3) / 1e9 / peak
            line["jacobian_sweep"] = sweep
        if world == 1 and not opts.no_cpu:
            cb, _ = cpu_baseline(args, x, E, target_seconds=12.0)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------
def run_c5(opts):
    """Secondary workload (BASELINE.json configs[4], SURVEY 8(d) "C5"): independent Dubins
    time-optimal problems (1 vehicle, degree 10, DEG_ELEV 100, 16 point obstacles each);
    one step = one FD sweep (nvar+1 = 16 evals) of every problem of this rank's block.
    Problems are dealt out in contiguous blocks (no collective).  Not the headline line:
    run with `--workload c5`."""
    import torch
    import torch.distributed as dist
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    from optimalbeziertrajectorygeneration_b200 import sharding
    from optimalbeziertrajectorygeneration_b200.batch import ProblemBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Mtot = opts.problems * world                                # weak scaling: problems per rank fixed
    lo, hi = sharding.block_range(Mtot, world, rank)
    M, E, nobs, deg = hi - lo, 100, 16, 10
    template = dict(numVeh=1, dimension=2, degree=deg, minimizeGoal='TimeOpt', maxSep=1, maxSpeed=3,
                    maxAngRate=np.pi / 2, initPoints=[(0, 0)], finalPoints=[(12, 8)], initSpeeds=[1],
                    finalSpeeds=[1], tf=8, initAngs=[np.pi / 2], finalAngs=[0])
    sets = np.stack([np.random.default_rng(p).uniform(1.0, 11.0, size=(nobs, 2)) for p in range(lo, hi)])
    pb = ProblemBatch(template, sets)
    guess = gopt.BezOptimization(**template).generateGuess(std=0)
    X = guess[None, :] + np.stack([np.random.default_rng(10 ** 6 + p).normal(size=guess.size) * 0.5
                                   for p in range(lo, hi)])
    X[:, -1] = np.abs(X[:, -1]) + 4.0                            # tf stays positive
    Xp, dx = pb.fd_points(X)
    nv1 = pb.nvar + 1
    d_X = Xp.view(M * nv1, pb.nvar)

    def step():
        return pb.evaluate(d_X, nv1, elev=E)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, opts.warmup)):
        F = step()
    sampler = ClockSampler(local, period=opts.clock_period)
    if rank == 0:
        sampler.open()
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(opts.steps):
        F = step()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=pb.eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    # per-block kernel times (CUDA events) to name the dominant kernel
    cpts, tf = pb.eng.assemble(d_X, E, obst_sets=pb.d_obst, evals_per_set=nv1)
    parts = {}
    for name, fn in (("separation", lambda: pb.eng.separation(cpts, E, 1.0, pair_begin=0, npairs=pb.npairs_x)),
                     ("maxspeed", lambda: pb.eng.speed(cpts, tf, E, -1.0, 9.0)),
                     ("angrate", lambda: pb.eng.angrate(cpts, tf, E, -1.0, (np.pi / 2) ** 2))):
        fn()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            fn()
        b_.record()
        torch.cuda.synchronize()
        parts[name] = a.elapsed_time(b_) / 3
    if rank == 0:
        L, A = 2 * deg + E + 1, 4 * (deg + E) + 1
        evals = M * nv1 * world * opts.steps
        alg_eval = 8.0 * (pb.npairs_x * L + L + A + 2 * (deg + 1) + 1)       # 20 168 B
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        m_ = deg + E
        fma_eval = 4 * (m_ + 1) ** 2 + 2 * (2 * m_ + 1) ** 2                    # 146 966 MAC
        line = {"metric": "constraint+Jacobian evals/sec", "value": evals / (ms * 1e-3), "unit": "evals/s",
                "n_gpus": world, "steps": opts.steps, "warmup": max(3, opts.warmup), "ms_per_step": ms / opts.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C5 synthetic Dubins batch: %d problems per GPU x (nvar+1 = %d) FD points, "
                                       "1 vehicle + 16 point obstacles, dim 2, degree 10, DEG_ELEV 100: 16 separation "
                                       "rows + max-speed row + angular-rate row per eval" % (M, nv1),
                           "problems_per_gpu": M, "sharding": "contiguous blocks of problems per rank, no collective"},
                "kernel_ms": parts,
                "roofline": {"bound": "fp64 (angular rate) / hbm (separation rows)",
                             "hbm_frac_whole_step": alg_eval * M * nv1 / (ms / opts.steps * 1e-3) / 1e9 / peak,
                             "angrate_fp64_frac": fma_eval * M * nv1 / (parts["angrate"] * 1e-3) / (64 * 148 * 1.965e9),
                             "separation_hbm_frac": 8.0 * pb.npairs_x * L * M * nv1 / (parts["separation"] * 1e-3) / 1e9 / peak,
                             "peak": peak, "unit": "GB/s",
                             "note": "fp64 peak = 64 FMA/clk/SM x 148 SMs x 1.965 GHz (tools/pipe_bench.cu)"},
                "gpu_launches": 4 * opts.steps, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="evals (x vectors) per step per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--clock-period", type=float, default=0.002, help="NVML sampling period in seconds")
    ap.add_argument("--nccl-gather", action="store_true",
                    help="multi-GPU: use the NCCL all-gather instead of the fused in-kernel peer stores")
    ap.add_argument("--no-sweep", action="store_true", help="skip the closed-form Jacobian sweep leg")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4 = the headline swarm (default); c5 = batch of independent Dubins problems")
    ap.add_argument("--problems", type=int, default=8192, help="c5: problems per GPU")
    opts = ap.parse_args()
    if opts.impl == "reference":
        run_reference(opts)
    elif opts.workload == "c5":
        run_c5(opts)
    else:
        run_ours(opts)


if __name__ == "__main__":
    main()
