"""The warp-specialised kernels (csrc/sq_elev_ws.cuh, jac_sq_elev_ws_kernel) against the 8-warp
tensor-path kernels they replace on the headline shapes: both evaluate the same folded sums with
the same instruction sequence per item, so every output must agree BIT FOR BIT -- rows, fused
per-pair minima, the packed active bitmask and the compacted list -- for full launches, pair
sub-ranges written into a pitched minimum matrix, minima-only launches and ragged last tiles.
BEZGPU_MMA_FLAGS bit 128 keeps a launch on the 8-warp kernel, bit 4 additionally replaces its TMA
row fetch by per-lane global loads (development builds only; ignored otherwise)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gopt():
    import torch
    assert torch.cuda.is_available()
    from optimalbeziertrajectorygeneration_b200 import optimization
    yield optimization
    optimization.DEG_ELEV = 0


def _setup(gopt, nveh, deg, E, B):
    from oracle.make_golden import synthetic_swarm_args
    args, x = synthetic_swarm_args(nveh, deg=deg)
    b = gopt.BezOptimization(**args)
    X = x[None, :] + np.random.default_rng(11).normal(size=(B, x.size)) * 0.05
    eng = b._engine(True)
    cpts, tf = eng.assemble(eng.upload(X), E)
    return b, eng, cpts, tf, x


@pytest.mark.parametrize("nveh,deg,E,B", [(97, 10, 100, 3),     # C4's shape, ragged last tile (4656 pairs x 3)
                                          (40, 10, 100, 1),     # fewer tiles than scheduler groups
                                          (64, 8, 60, 2),       # L = 77
                                          (33, 12, 100, 2)])    # degree 12, L = 125
def test_ws_pair_kernel_bit_identical_to_8warp_kernel(gopt, monkeypatch, nveh, deg, E, B):
    import torch
    from optimalbeziertrajectorygeneration_b200.engine import ActiveSet, num_pairs
    b, eng, cpts, tf, _ = _setup(gopt, nveh, deg, E, B)
    P = num_pairs(eng.N)
    L = 2 * deg + E + 1
    res = {}
    for flags in ("0", "128"):
        monkeypatch.setenv("BEZGPU_MMA_FLAGS", flags)
        out = torch.full((B, P, L), float("nan"), dtype=torch.float64, device=eng.device)
        pm = torch.full((B, P), float("nan"), dtype=torch.float64, device=eng.device)
        act = ActiveSet(B * P, capacity=B * P, device=eng.device)
        act.reset()
        eng.separation(cpts, E, 0.9, out=out, pairmin=pm, active=act)
        rows_only = eng.separation(cpts, E, 0.9)
        pm_only = torch.full((B, P), float("nan"), dtype=torch.float64, device=eng.device)
        eng.separation(cpts, E, 0.9, pairmin=pm_only, rows=False)
        # pair sub-ranges into one pitched [B, P] minimum matrix (the strong-scaling layout)
        pitched = torch.full((B, P), float("nan"), dtype=torch.float64, device=eng.device)
        cuts = [0, 33, 1000 if P > 1200 else P // 2, P]
        parts = []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            parts.append(eng.separation(cpts, E, 0.9, pair_begin=lo, npairs=hi - lo, pairmin=pitched[:, lo:],
                                        min_pitch=P))
        torch.cuda.synchronize()
        flags_, idx, val, over = ActiveSet.decode(act.buf.cpu().numpy(), B * P, act.capacity)
        assert not over
        assert torch.equal(pm, out.min(dim=2).values)
        assert torch.equal(rows_only, out) and torch.equal(pm_only, pm) and torch.equal(pitched, pm)
        assert torch.equal(torch.cat(parts, dim=1), out)
        pmh = pm.cpu().numpy().ravel()
        assert np.array_equal(flags_, pmh < 0) and np.array_equal(idx, np.nonzero(pmh < 0)[0])
        assert np.array_equal(val, pmh[pmh < 0])
        res[flags] = (out, pm)
    assert torch.equal(res["0"][0], res["128"][0])
    assert torch.equal(res["0"][1], res["128"][1])


def test_ws_kernel_propagates_nan_vehicles_like_the_8warp_kernel(gopt, monkeypatch):
    """A NaN vehicle poisons exactly its own pairs (the TMA-fetched rows share a shared-memory region
    with the staged rows: padding slots must be rewritten per tile)."""
    import torch
    from optimalbeziertrajectorygeneration_b200.engine import num_pairs
    b, eng, cpts, tf, _ = _setup(gopt, 50, 10, 100, 2)
    cpts = cpts.clone()
    cpts[:, 7] = float("nan")
    cpts[1, 31] = float("nan")
    P = num_pairs(eng.N)
    outs = []
    for flags in ("0", "128"):
        monkeypatch.setenv("BEZGPU_MMA_FLAGS", flags)
        pm = torch.empty((2, P), dtype=torch.float64, device=eng.device)
        out = eng.separation(cpts, 100, 0.9, pairmin=pm)
        outs.append((out, pm))
    ii, jj = np.triu_indices(50, 1)
    for bi, bad in ((0, {7}), (1, {7, 31})):
        want = torch.as_tensor(np.isin(ii, list(bad)) | np.isin(jj, list(bad)), device=eng.device)
        for out, pm in outs:
            assert torch.equal(torch.isnan(pm[bi]), want)
            assert torch.equal(torch.isnan(out[bi]).all(dim=1), want) and torch.equal(torch.isnan(out[bi]).any(dim=1), want)
    a, c = outs
    assert torch.equal(torch.nan_to_num(a[0]), torch.nan_to_num(c[0]))
    assert torch.equal(torch.nan_to_num(a[1]), torch.nan_to_num(c[1]))


@pytest.mark.parametrize("nveh,deg,E", [(40, 10, 100), (23, 6, 80)])
def test_ws_jacobian_sweep_bit_identical_to_8warp_kernel(gopt, monkeypatch, nveh, deg, E):
    """Warp-specialised separation sweep (TMA partner-row fetch, one-hot Bernstein product) == the 8-warp
    sweep kernel, and both == the dense J^T layout (DFMA kernel, dense double loop) row for row."""
    import torch
    from optimalbeziertrajectorygeneration_b200.engine import num_pairs
    b, eng, cpts, tf, x = _setup(gopt, nveh, deg, E, 1)
    L = 2 * deg + E + 1
    sweeps = []
    for flags in ("0", "128"):
        monkeypatch.setenv("BEZGPU_MMA_FLAGS", flags)
        sweeps.append(eng.jac_separation(x, E, dense=False).clone())
    assert torch.equal(sweeps[0], sweeps[1])
    monkeypatch.setenv("BEZGPU_MMA_FLAGS", "0")
    JT = eng.jac_separation(x, E, dense=True)                  # [nvar, P * L]
    N = eng.N
    sw = sweeps[0].view(-1, N - 1, L)                          # [variable][partner curve][L]
    JT3 = JT.view(JT.shape[0], num_pairs(N), L)
    ncols = eng.ncols
    rng = np.random.default_rng(5)
    for k in rng.choice(sw.shape[0], size=12, replace=False):
        v = int(k) // (eng.dim * ncols)
        for uu in rng.choice(N - 1, size=6, replace=False):
            u = int(uu) + (1 if uu >= v else 0)
            i, j = min(u, v), max(u, v)
            p = i * (2 * N - i - 1) // 2 + (j - i - 1)
            got, want = sw[k, uu].cpu().numpy(), JT3[k, p].cpu().numpy()
            scale = np.abs(want).max()
            assert np.abs(got - want).max() <= 1e-12 * max(scale, 1e-300)
