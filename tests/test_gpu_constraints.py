"""GPU parity of the fused constraint kernels (through the C-ABI) against the
numpy oracle and the committed golden vectors from the reference.
Tolerance: north_star's fp64 relative error <= 1e-9 per constraint block
(max |gpu-ref| / max |ref|)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def gopt():
    import torch
    assert torch.cuda.is_available()
    from optimalbeziertrajectorygeneration_b200 import optimization
    yield optimization
    optimization.DEG_ELEV = 0


def _check_blocks(gopt, bezopt, x, g, prefix, E, names):
    gopt.DEG_ELEV = E
    fns = {"sep": bezopt.temporalSeparationConstraints, "maxspeed": bezopt.maxSpeedConstraints,
           "minspeed": bezopt.minSpeedConstraints}
    if "angrate" in names:
        fns["angrate"] = bezopt.maxAngularRateConstraints
    for name in names:
        key = "%s_%s_E%d" % (prefix, name, E)
        if key not in g.files:
            continue
        got = fns[name](x)
        assert got.shape == g[key].shape
        assert relerr(got, g[key]) < RTOL, key


def test_swarm_golden(gopt, golden):
    g = golden("constraints")
    b = gopt.BezOptimization(numVeh=36, dimension=3, degree=5, minimizeGoal='Euclidean', maxSep=0.9,
                             initPoints=g["swarm_initPts"], finalPoints=g["swarm_finalPts"])
    for xi in (0, 1):
        for E in (0, 10, 100):
            _check_blocks(gopt, b, g["swarm_x%d" % xi], g, "swarm_x%d" % xi, E,
                          ("sep", "maxspeed", "minspeed"))


def test_c4_like_golden(gopt, golden):
    from oracle.make_golden import synthetic_swarm_args
    g = golden("constraints")
    for N in (2, 16, 33):
        args, x = synthetic_swarm_args(N)
        b = gopt.BezOptimization(**args)
        _check_blocks(gopt, b, x, g, "c4_N%d" % N, 100, ("sep", "maxspeed", "minspeed"))


def test_example1_golden(gopt, golden):
    g = golden("constraints")
    b = gopt.BezOptimization(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1,
                             maxSpeed=5, maxAngRate=1, initPoints=[(0, 5), (3, 0)],
                             finalPoints=[(8, 4), (7, 10)], initSpeeds=[1, 1], finalSpeeds=[1, 1],
                             initAngs=[0, np.pi / 2], finalAngs=[0, np.pi / 2],
                             pointObstacles=[[3, 2], [6, 7]])
    assert np.array_equal(b.reshapeVector(g["ex1_x0"]), g["ex1_y0"])     # bit exact A0
    for xi in (0, 1):
        for E in (0, 10, 100):
            _check_blocks(gopt, b, g["ex1_x%d" % xi], g, "ex1_x%d" % xi, E,
                          ("sep", "maxspeed", "minspeed", "angrate"))


def test_c5_like_golden(gopt, golden):
    from oracle.make_golden import dubins_problem_args
    g = golden("constraints")
    for seed in (0, 1, 2):
        b = gopt.BezOptimization(**dubins_problem_args(seed))
        x = g["c5_s%d_x" % seed]
        assert np.array_equal(b.reshapeVector(x), g["c5_s%d_y" % seed])
        for E in ((100, 0, 5) if seed == 0 else (100,)):
            _check_blocks(gopt, b, x, g, "c5_s%d" % seed, E, ("sep", "maxspeed", "minspeed", "angrate"))


@pytest.mark.parametrize("dim,deg,N,E", [(1, 3, 5, 0), (2, 1, 4, 3), (3, 7, 9, 21), (2, 12, 7, 64),
                                         (3, 16, 5, 40), (3, 10, 40, 300), (2, 4, 300, 1),
                                         # 33..64 column pairs: the fp64 tensor-path kernels (DMMA + TMA stores)
                                         (3, 10, 40, 100), (2, 5, 33, 60), (1, 15, 20, 97), (3, 4, 37, 56),
                                         (2, 10, 12, 43), (3, 10, 9, 44), (2, 8, 11, 75), (3, 13, 70, 101)])
def test_random_models_vs_oracle(gopt, dim, deg, N, E):
    from oracle import bezier_oracle as O
    rng = np.random.default_rng(deg * 100 + N)
    args = dict(numVeh=N, dimension=dim, degree=deg, minimizeGoal='Euclidean', maxSep=0.7,
                minSpeed=0.3, maxSpeed=4.0, tf=float(rng.uniform(1, 30)),
                initPoints=rng.uniform(-5, 5, size=(N, dim)), finalPoints=rng.uniform(-5, 5, size=(N, dim)))
    b = gopt.BezOptimization(**args)
    x = rng.normal(size=b.nvar) * 3
    f = O.make_callables(O.Model(**args), E)
    gopt.DEG_ELEV = E
    assert np.array_equal(b.reshapeVector(x), O.reshape_vector(O.Model(**args), x))
    assert relerr(b.temporalSeparationConstraints(x), f["sep"](x)) < RTOL
    assert relerr(b.maxSpeedConstraints(x), f["maxspeed"](x)) < RTOL
    assert relerr(b.minSpeedConstraints(x), f["minspeed"](x)) < RTOL


def test_batched_and_pair_ranges(gopt):
    """evaluate_batch == per-x calls; a split pair range == the full range
    (what the multi-GPU sharding relies on)."""
    import torch
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(33)
    b = gopt.BezOptimization(**args)
    rng = np.random.default_rng(0)
    X = x[None, :] + rng.normal(size=(5, x.size)) * 0.1
    gopt.DEG_ELEV = 100
    res = b.evaluate_batch(X, which=("sep", "maxspeed", "minspeed"))
    for i in range(5):
        assert np.array_equal(res["sep"][i].cpu().numpy(), b.temporalSeparationConstraints(X[i]))
        assert np.array_equal(res["maxspeed"][i].cpu().numpy(), b.maxSpeedConstraints(X[i]))
    eng = b._engine(True)
    cpts, _ = eng.assemble(eng.upload(X), 100)
    full = eng.separation(cpts, 100, 0.9)
    P = full.shape[1]
    cuts = [0, 1, 17, 255, 256, 300, P]
    parts = [eng.separation(cpts, 100, 0.9, pair_begin=a, npairs=c - a) for a, c in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat(parts, dim=1), full)
    pm = torch.empty((5, P), dtype=torch.float64, device=full.device)
    eng.separation(cpts, 100, 0.9, pairmin=pm)
    assert torch.equal(pm, full.min(dim=2).values)


@pytest.mark.parametrize("deg,E", [(10, 100),                       # C4 / C5: L = 121 (4 n-tile pairs)
                                   (5, 0), (5, 10), (5, 100),      # C3 swarm: L = 11 / 21 / 111
                                   (10, 0), (10, 30),              # C2 Example1: L = 21 / 51
                                   (3, 10), (7, 0), (2, 0), (15, 33), (12, 103),
                                   (10, 300), (5, 150), (12, 106), (4, 121)])   # L > 128: column-tiled variant
def test_tensor_path_matches_dfma_path(gopt, monkeypatch, deg, E):
    """The DMMA kernels (TMA-store variant for L <= 128, column-tiled variant above; degree <= 15)
    and the column-stationary DFMA kernels are two evaluations of the same folded sums: values
    agree to rounding, per-pair minima follow their values."""
    import torch
    from optimalbeziertrajectorygeneration_b200 import _capi, engine
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(70, deg=deg)
    b = gopt.BezOptimization(**args)
    X = x[None, :] + np.random.default_rng(3).normal(size=(3, x.size)) * 0.05
    eng = b._engine(True)
    cpts, tf = eng.assemble(eng.upload(X), E)
    outs = {}
    for force in ("0", "1"):
        monkeypatch.setenv("BEZGPU_FORCE_DFMA", force)
        sep = eng.separation(cpts, E, 0.9)
        pm = torch.empty(sep.shape[:2], dtype=torch.float64, device=sep.device)
        eng.separation(cpts, E, 0.9, pairmin=pm)
        spd = eng.speed(cpts, tf, E, -1.0, 25.0)
        assert torch.equal(pm, sep.min(dim=2).values)
        outs[force] = (sep.cpu().numpy(), spd.cpu().numpy())
    assert relerr(outs["0"][0], outs["1"][0]) < 1e-13
    assert relerr(outs["0"][1], outs["1"][1]) < 1e-13
    assert not np.array_equal(outs["0"][0], outs["1"][0])     # really two different kernels


def test_c4_full_size_properties(gopt):
    """BASELINE configs[3] at full size (1024 vehicles, 523 776 pairs x 121 values): sampled
    pair blocks against the C oracle, and size-independent properties over every pair --
    end-point identity of degree elevation (b_0 = s_0, b_M = s_2n), invariance of the
    coefficient mean under elevation, fused per-pair minimum == row minimum, FD-perturbed
    x changes exactly the rows of the perturbed vehicle."""
    import torch
    from oracle import bezier_oracle as O
    from oracle import c_oracle as C
    from oracle.make_golden import synthetic_swarm_args
    N, E, n = 1024, 100, 10
    args, x = synthetic_swarm_args(N)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    X = np.stack([x, x])
    k = 3 * 9 * 500 + 4                      # a free control point of vehicle 500
    X[1, k] += 1.4901161193847656e-08
    cpts, tf = eng.assemble(eng.upload(X), E)
    P = N * (N - 1) // 2
    L = 2 * n + E + 1
    sep = torch.empty((2, P, L), dtype=torch.float64, device=eng.device)
    pm = torch.empty((2, P), dtype=torch.float64, device=eng.device)
    eng.separation(cpts, E, args["maxSep"], out=sep, pairmin=pm)
    y = O.reshape_vector(O.Model(**args), x)
    # (1) sampled contiguous pair blocks vs the C restatement of the reference
    for begin in (0, 1023, 77777, 261888, P - 4096):
        want = C.temporal_separation(y, N, 3, args["maxSep"], E, begin, 4096).reshape(4096, L)
        got = sep[0, begin:begin + 4096].cpu().numpy()
        assert relerr(got, want) < RTOL
    # (2) fused minimum is the row minimum, bit for bit, for all 2 x 523 776 pairs
    assert torch.equal(pm, sep.min(dim=2).values)
    # (3) end points and mean over all pairs (vectorised numpy restatement of sub -> normSquare)
    c = y.reshape(N, 3, n + 1)
    iu, ju = np.triu_indices(N, 1)
    d0 = c[iu, :, 0] - c[ju, :, 0]
    dn = c[iu, :, n] - c[ju, :, n]
    s0 = 1.5 * (d0 * d0).sum(axis=1) - args["maxSep"] ** 2
    sn = 1.5 * (dn * dn).sum(axis=1) - args["maxSep"] ** 2
    got0 = sep[0, :, 0].cpu().numpy()
    gotn = sep[0, :, L - 1].cpu().numpy()
    assert np.abs(got0 - s0).max() / np.abs(s0).max() < 1e-12
    assert np.abs(gotn - sn).max() / np.abs(sn).max() < 1e-12
    W = O.prod_weights(n)
    a = c[iu[:20000]] - c[ju[:20000]]                       # [pairs, 3, n+1]
    G = np.einsum('pdi,pdj->pij', a, a) * W
    smean = np.array([np.trace(G[:, ::-1], offset=n - kk, axis1=1, axis2=2) for kk in range(2 * n + 1)]).T
    smean = 1.5 * smean.mean(axis=1) - args["maxSep"] ** 2    # mean of the 2n+1 product coefficients
    gotmean = sep[0, :20000].mean(dim=1).cpu().numpy()
    assert np.abs(gotmean - smean).max() / np.abs(smean).max() < 1e-12
    # (4) the perturbed evaluation differs from the base one only in pairs that contain vehicle 500
    changed = (sep[0] != sep[1]).any(dim=1).cpu().numpy()
    touches = (iu == 500) | (ju == 500)
    assert not changed[~touches].any()
    assert changed[touches].mean() > 0.5


def test_fused_gather_stores_on_one_gpu(gopt):
    """bez_pair_sepsq_elev_p2p with 'peer' buffers that live on the same GPU: every minimum
    lands in the local matrix and in each extra destination (the store path of the fused
    all-gather; the two-GPU version is tests/test_gpu_multigpu.py)."""
    import torch
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(45)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    X = x[None, :] + np.random.default_rng(9).normal(size=(3, x.size)) * 0.05
    cpts, _ = eng.assemble(eng.upload(X), 100)
    P = 45 * 44 // 2
    local = torch.empty((3, P), dtype=torch.float64, device=eng.device)
    extra = [torch.full((3, P), float("nan"), dtype=torch.float64, device=eng.device) for _ in range(7)]
    sep = eng.separation(cpts, 100, 0.9, pairmin=local, peer_ptrs=[t.data_ptr() for t in extra])
    ref = sep.min(dim=2).values
    assert torch.equal(local, ref)
    for t in extra:
        assert torch.equal(t, ref)
    with pytest.raises(Exception):                       # more than BEZ_MAX_PEERS destinations
        eng.separation(cpts, 100, 0.9, pairmin=local, peer_ptrs=[t.data_ptr() for t in extra] + [local.data_ptr()])
    gopt.DEG_ELEV = 0
    small = gopt.BezOptimization(numVeh=4, dimension=2, degree=3, initPoints=np.zeros((4, 2)),
                                 finalPoints=np.ones((4, 2)))
    e2 = small._engine(True)
    c2, _ = e2.assemble(e2.upload(np.zeros((1, small.nvar))), 0)
    p6 = torch.empty((1, 6), dtype=torch.float64, device=e2.device)
    x6 = torch.full((1, 6), float("nan"), dtype=torch.float64, device=e2.device)
    s7 = e2.separation(c2, 0, 0.9, pairmin=p6, peer_ptrs=[x6.data_ptr()])     # L = 7: tensor path since round 2
    assert torch.equal(p6, s7.min(dim=2).values) and torch.equal(x6, p6)
    s137 = e2.separation(c2, 130, 0.9, pairmin=p6, peer_ptrs=[x6.data_ptr()])   # L = 137: column-tiled variant
    assert torch.equal(p6, s137.min(dim=2).values) and torch.equal(x6, p6)
    d1 = gopt.BezOptimization(numVeh=4, dimension=1, degree=3, initPoints=np.zeros((4, 1)),
                              finalPoints=np.ones((4, 1)))
    e1 = d1._engine(True)
    c1, _ = e1.assemble(e1.upload(np.zeros((1, d1.nvar))), 0)
    with pytest.raises(Exception):                       # dim 1: outside the tensor-path kernels
        e1.separation(c1, 0, 0.9, pairmin=p6, peer_ptrs=[local.data_ptr()])


def test_unaligned_and_ragged_outputs(gopt):
    """Output base that is only 8-byte aligned (no TMA bulk stores possible) and item counts that
    leave ragged last tiles / odd row counts: the copy fallback must give the same bits."""
    import torch
    from oracle.make_golden import synthetic_swarm_args
    for N, B, E in ((23, 3, 50), (12, 2, 100), (7, 5, 45)):
        args, x = synthetic_swarm_args(N)
        b = gopt.BezOptimization(**args)
        eng = b._engine(True)
        X = x[None, :] + np.random.default_rng(N).normal(size=(B, x.size)) * 0.05
        cpts, _ = eng.assemble(eng.upload(X), E)
        P, L = N * (N - 1) // 2, 2 * 10 + E + 1
        ref = eng.separation(cpts, E, 0.9)
        flat = torch.full((B * P * L + 1,), float("nan"), dtype=torch.float64, device=eng.device)
        odd = flat[1:].view(B, P, L)
        assert odd.data_ptr() % 16 == 8
        pm = torch.empty((B, P), dtype=torch.float64, device=eng.device)
        eng.separation(cpts, E, 0.9, out=odd, pairmin=pm)
        assert torch.equal(odd, ref) and torch.equal(pm, ref.min(dim=2).values)
        part = eng.separation(cpts, E, 0.9, pair_begin=3, npairs=P - 5)      # ragged range
        assert torch.equal(part, ref[:, 3:P - 2])


def test_evaluate_sweep_matches_serial_calls(gopt):
    """The pipelined sweep (two workspaces, copy stream) returns what per-chunk
    evaluate_reduced calls return, including a ragged last chunk."""
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(40)
    b = gopt.BezOptimization(**args)
    X = x[None, :] + np.random.default_rng(5).normal(size=(11, x.size)) * 0.02
    sw = b.evaluate_sweep(X, elev=100, chunk=4)
    pm, sp = sw["pairmin"].copy(), sw["maxspeed"].copy()
    for lo in range(0, 11, 4):
        red = b.evaluate_reduced(X[lo:lo + 4], elev=100)
        assert np.array_equal(pm[lo:lo + 4], red["pairmin"])
        assert np.array_equal(sp[lo:lo + 4], red["maxspeed"])
    sw2 = b.evaluate_sweep(X[:3], elev=100, chunk=8)          # fewer rows than the chunk
    assert np.array_equal(sw2["pairmin"], pm[:3])


@pytest.mark.parametrize("deg,elev,K", [(3, 10, 37), (10, 100, 70), (5, 0, 1)])
def test_sequential_swarm_constraint(gopt, deg, elev, K):
    """Examples/SequentialSwarm.py:43-70: new vehicle vs frozen trajectories, one minimum each."""
    from oracle import bezier_oracle as O
    from optimalbeziertrajectorygeneration_b200.sequential import FrozenSwarm
    rng = np.random.default_rng(deg + K)
    traj = rng.uniform(0, 100, size=(K, 3, deg + 1))
    new = rng.uniform(0, 100, size=(4, 3, deg + 1))
    fs = FrozenSwarm(traj, elev=elev)
    got = fs.separation_minima(new, 1.0)
    W, T = O.prod_weights(deg), O.elev_matrix(2 * deg, elev)
    want = np.empty((4, K))
    for b in range(4):
        for i in range(K):
            a = new[b] - traj[i]
            G = np.einsum('di,dj->ij', a, a) * W
            s = 1.5 * np.array([np.trace(G[::-1], offset=k - deg) for k in range(2 * deg + 1)])
            want[b, i] = (s @ T).min() - 1.0
    assert relerr(got, want) < RTOL
    assert np.array_equal(fs.separation_minima(new[2], 1.0), got[2])
    assert FrozenSwarm(np.zeros((0, 3, deg + 1)), elev=elev).separation_minima(new[0], 1.0).shape == (0,)


def test_single_vehicle_returns_none(gopt):
    b = gopt.BezOptimization(numVeh=1, dimension=2, degree=5, initPoints=[(0, 0)], finalPoints=[(1, 1)])
    assert b.temporalSeparationConstraints(np.zeros(b.nvar)) is None


# --------------------------------------------------------------------------
# A7: finite-difference Jacobians.  Judged against the exactly rounded quotient
# (oracle evaluated in exact rational arithmetic on the fp64 points), tolerance 1e-9 relative
# to the largest entry; the reference's own literal FD sits ~1e-7 away from
# that value (its noise floor) and is checked at that level.
JTOL = 1e-9


def _jac_check(gopt, args, x, E, golden_J=None, names=("sep", "maxspeed", "minspeed")):
    from oracle import bezier_oracle as O
    b = gopt.BezOptimization(**args)
    gopt.DEG_ELEV = E
    fld = O.make_callables(O.Model(**args), E, dtype=object)
    fns = {"sep": b.temporalSeparationConstraints_jac, "maxspeed": b.maxSpeedConstraints_jac,
           "minspeed": b.minSpeedConstraints_jac}
    for name in names:
        J = fns[name](x)
        Jex = O.fd_jacobian_exact(fld[name], x)
        assert J.shape == Jex.shape
        scale = np.abs(Jex).max()
        assert np.abs(J - Jex).max() / scale < JTOL, name
        # sparsity: rows that do not depend on x_k are exact zeros in both; entries that
        # are zero only by cancellation (e.g. the fixed end speed of a Dubins curve w.r.t.
        # tf) may be O(ulp) instead of 0
        assert np.all(np.abs(Jex[J == 0]) <= 1e-13 * scale), name
        assert np.all(np.abs(J[Jex == 0]) <= 1e-13 * scale), name
        if golden_J is not None and name in golden_J:
            Jref = golden_J[name]
            assert np.abs(J - Jref).max() / scale < 5e-6        # reference FD noise floor
            assert np.all(np.abs(Jref[J == 0]) <= 5e-6 * scale), name
            # columns of variables that move no curve of a row are exact zeros in both
            untouched = np.all(J == 0, axis=0)
            assert np.all(Jref[:, untouched] == 0), name


def test_jacobian_swarm(gopt, golden):
    from oracle.make_golden import synthetic_swarm_args
    g = golden("jacobian")
    args, x = synthetic_swarm_args(6, deg=5, seed=11)
    _jac_check(gopt, args, x, 10, {"sep": g["sw6_J_sep_E10"], "maxspeed": g["sw6_J_maxspeed_E10"]})


def test_jacobian_dubins_timeopt_obstacles(gopt, golden):
    from oracle.make_golden import dubins_problem_args
    g = golden("jacobian")
    args = dubins_problem_args(5, nobs=4, deg=6)
    _jac_check(gopt, args, g["dub_x"], 8, {"sep": g["dub_J_sep_E8"], "maxspeed": g["dub_J_maxspeed_E8"]})


def test_jacobian_example1(gopt, golden):
    g = golden("jacobian")
    args = dict(numVeh=2, dimension=2, degree=10, minimizeGoal='TimeOpt', maxSep=1, maxSpeed=5,
                maxAngRate=1, initPoints=[(0, 5), (3, 0)], finalPoints=[(8, 4), (7, 10)],
                initSpeeds=[1, 1], finalSpeeds=[1, 1], initAngs=[0, np.pi / 2],
                finalAngs=[0, np.pi / 2], pointObstacles=[[3, 2], [6, 7]])
    for E in (0, 10):
        _jac_check(gopt, args, g["ex1_x"], E, {"sep": g["ex1_J_sep_E%d" % E],
                                               "maxspeed": g["ex1_J_maxspeed_E%d" % E]})


def test_jacobian_sweep_layout_matches_dense(gopt):
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(9, deg=7, seed=4)
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    E = 33
    JT = eng.jac_separation(x, E, dense=True).cpu().numpy()          # [nvar, P*L]
    sw = eng.jac_separation(x, E, dense=False).cpu().numpy()         # [nvar*(N-1), L]
    N, L = eng.N, 2 * 7 + E + 1
    sw = sw.reshape(eng.nvar, N - 1, L)
    for k in range(eng.nvar):
        v = k // (eng.dim * eng.ncols)
        others = [u for u in range(N) if u != v]
        for uu, u in enumerate(others):
            i, j = min(v, u), max(v, u)
            p = i * (2 * N - i - 1) // 2 + (j - i - 1)
            assert np.array_equal(JT[k, p * L:(p + 1) * L], sw[k, uu])


@pytest.mark.parametrize("deg,E,N,dim", [(10, 100, 9, 3), (6, 60, 7, 2), (12, 41, 5, 3)])
def test_jacobian_sweep_tensor_path_matches_dense(gopt, deg, E, N, dim):
    """65 <= L <= 128: the sweep layout runs the DMMA kernel, the dense layout the DFMA kernel;
    both evaluate the same closed-form quotient (agreement to rounding), and the sweep rows
    also match the exactly rounded FD quotient of the oracle."""
    from oracle import bezier_oracle as O
    rng = np.random.default_rng(100 * deg + N)
    args = dict(numVeh=N, dimension=dim, degree=deg, minimizeGoal='Euclidean', maxSep=0.7, maxSpeed=4.0,
                tf=12.0, initPoints=rng.uniform(-5, 5, size=(N, dim)), finalPoints=rng.uniform(-5, 5, size=(N, dim)))
    b = gopt.BezOptimization(**args)
    eng = b._engine(True)
    x = rng.normal(size=b.nvar) * 3
    JT = eng.jac_separation(x, E, dense=True).cpu().numpy()          # [nvar, P*L]   (DFMA kernel)
    sw = eng.jac_separation(x, E, dense=False).cpu().numpy()         # [nvar*(N-1), L] (DMMA kernel)
    L = 2 * deg + E + 1
    sw = sw.reshape(eng.nvar, N - 1, L)
    scale = np.abs(JT).max()
    for k in range(eng.nvar):
        v = k // (eng.dim * eng.ncols)
        others = [u for u in range(N) if u != v]
        for uu, u in enumerate(others):
            i, j = min(v, u), max(v, u)
            p = i * (2 * N - i - 1) // 2 + (j - i - 1)
            assert np.abs(JT[k, p * L:(p + 1) * L] - sw[k, uu]).max() <= 1e-13 * scale
    # rows that do not depend on the variable are exactly zero in the dense layout
    gopt.DEG_ELEV = E
    Jfun = b.temporalSeparationConstraints_jac(x)                    # [m, nvar], public closure
    assert relerr(Jfun, JT.T) < 1e-15


def test_angrate_rejects_3d(gopt):
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(3)
    b = gopt.BezOptimization(**args)
    with pytest.raises(ValueError):
        b.maxAngularRateConstraints(x)


def test_jacobian_angrate_literal_fd(gopt, golden):
    """The angular-rate Jacobian is the batched literal quotient: it must agree
    with the reference's own FD Jacobian to the FD noise floor."""
    from oracle.make_golden import dubins_problem_args
    g = golden("jacobian")
    b = gopt.BezOptimization(**dubins_problem_args(5, nobs=4, deg=6))
    gopt.DEG_ELEV = 8
    J = b.maxAngularRateConstraints_jac(g["dub_x"])
    Jref = g["dub_J_angrate_E8"]
    assert J.shape == Jref.shape
    assert np.abs(J - Jref).max() / np.abs(Jref).max() < 1e-5


def test_objectives(gopt, golden):
    """A14: cost callables (swarm Euclidean objective at x0 from SURVEY section 4)."""
    from oracle.make_golden import synthetic_swarm_args
    g = golden("constraints")
    gopt.DEG_ELEV = 0
    b = gopt.BezOptimization(numVeh=36, dimension=3, degree=5, minimizeGoal='Euclidean', maxSep=0.9,
                             initPoints=g["swarm_initPts"], finalPoints=g["swarm_finalPts"])
    assert b.objectiveFunction(g["swarm_x0"]) == pytest.approx(382.0101330471013, rel=1e-13)
    args, x = synthetic_swarm_args(5, deg=6, seed=3)
    for goal in ("Euclidean", "Accel"):
        a = dict(args)
        a["minimizeGoal"] = goal
        bb = gopt.BezOptimization(**a)
        assert bb.objectiveFunction(x) == pytest.approx(float(g["obj_%s" % goal]), rel=1e-12)
    a = dict(args)
    a["minimizeGoal"] = "nonsense"
    with pytest.raises(ValueError):
        gopt.BezOptimization(**a).objectiveFunction
    tb = gopt.BezOptimization(numVeh=1, dimension=2, degree=5, minimizeGoal='TimeOpt', initPoints=[(0, 0)],
                              finalPoints=[(1, 1)])
    assert tb.objectiveFunction(np.arange(9.0)) == 8.0
