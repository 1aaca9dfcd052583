"""GPU parity of the geometry kernels (split, extrema, GJK, minDist family)
through the C-ABI: bit-exact against the pure-Python oracle and against the
golden vectors of the unmodified reference (collision flags, (alpha,t1,t2))."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def gbez():
    import torch
    assert torch.cuda.is_available()
    from optimalbeziertrajectorygeneration_b200 import bezier
    return bezier


def test_gjk_golden_bit_exact(golden):
    from optimalbeziertrajectorygeneration_b200.gjk import gjk as ggjk
    g = golden("geometry_gjk")
    p1 = np.nan_to_num(g["poly1"], nan=0.0)
    p2 = np.nan_to_num(g["poly2"], nan=0.0)
    flag, a, b, dist = ggjk.gjk_batch(p1, p2, g["n1"], g["n2"])
    assert np.array_equal(flag, g["flag"])                    # bit-identical collision flags
    ok = g["flag"] > 0
    assert np.array_equal(dist[ok], g["dist"][ok])
    assert np.array_equal(a[ok], g["p1"][ok])
    assert np.array_equal(b[ok], g["p2"][ok])
    # the drop-in single call
    k = int(np.argmax(ok))
    f, info = ggjk.gjkNew(g["poly1"][k][:g["n1"][k]], g["poly2"][k][:g["n2"][k]])
    assert f == 1 and info[2] == g["dist"][k]
    f, info = ggjk.gjkNew(g["poly1"][0][:g["n1"][0]], g["poly2"][0][:g["n2"][0]])
    assert f == 0 and info == ()


def test_gjk_random_vs_oracle():
    from oracle import gjk_oracle as G
    from optimalbeziertrajectorygeneration_b200.gjk import gjk as ggjk
    rng = np.random.default_rng(77)
    count = 300
    P1 = rng.normal(size=(count, 11, 3))
    P2 = rng.normal(size=(count, 7, 3)) + rng.normal(size=(count, 1, 3)) * 2
    flag, a, b, dist = ggjk.gjk_batch(P1, P2)
    for k in range(count):
        f, p1, p2, d = G.gjk_new(P1[k], P2[k])
        assert f == flag[k]
        if f > 0:
            assert d == dist[k] and np.array_equal(p1, a[k]) and np.array_equal(p2, b[k])


def test_split_and_eval_bit_exact(gbez, golden):
    g = golden("algebra")
    for ci in range(int(g["ncases"])):
        k = "c%02d_" % ci
        c, tf = g[k + "cpts"], float(g[k + "tf"])
        b = gbez.Bezier(c.copy(), tf=tf)
        l, r = b.split(float(g[k + "tdiv"]))
        assert np.array_equal(l.cpts, g[k + "split_l"]) and np.array_equal(r.cpts, g[k + "split_r"])
        assert l.tf == float(g[k + "tdiv"]) and r.t0 == float(g[k + "tdiv"])
        assert np.array_equal(b(g[k + "tau"]), g[k + "eval"])


def test_bezier_methods_golden(gbez, golden):
    from conftest import relerr
    g = golden("algebra")
    for ci in range(int(g["ncases"])):
        k = "c%02d_" % ci
        c, tf, other = g[k + "cpts"], float(g[k + "tf"]), g[k + "other"]
        b = gbez.Bezier(c.copy(), tf=tf)
        o = gbez.Bezier(other.copy(), tf=tf)
        assert relerr(b.normSquare().cpts, g[k + "normsq"]) < 1e-12
        for R in (0, 1, 7, 30):
            assert relerr(b.elev(R).cpts, g[k + "elev%d" % R]) < 1e-12
        assert relerr(b.diff().cpts, g[k + "diff"]) < 1e-12
        assert relerr((b * o).cpts, g[k + "mul"]) < 1e-12
        assert np.array_equal((b - o).cpts, g[k + "sub"]) and np.array_equal((b + o).cpts, g[k + "add"])
        assert np.allclose(b.integrate(), g[k + "integrate"], rtol=1e-14)
    with pytest.raises(TypeError):
        b.mul(3.0)
    with pytest.raises(ValueError):
        gbez.Bezier(np.zeros((2, 4))).mul(gbez.Bezier(np.zeros((3, 4))))
    with pytest.raises(ValueError):
        gbez.Bezier(np.zeros((1, 4))).minDist(gbez.Bezier(np.zeros((1, 4))))


def test_mindist_golden_bit_exact(gbez, golden):
    g = golden("geometry_mindist")
    for tag in ("named", "r33", "r35", "r24"):
        out, status = gbez.min_dist_batch(g[tag + "_a"], g[tag + "_b"])
        assert np.all(status == 0)
        assert np.array_equal(out, g[tag + "_r"]), tag          # bit-identical (alpha, t1, t2)
    # drop-in call (Examples/MinDistBez2Bez.py:87-90)
    a, b = g["named_a"], g["named_b"]
    r = gbez.Bezier(a[1].copy()).minDist(gbez.Bezier(b[1].copy()))
    assert r == (1.4142135623789327, 0.7999994253499804, 1.0)


def test_mindist2poly_and_collcheck_golden(gbez, golden):
    g = golden("geometry_mindist")
    c1 = np.array([(0, 1, 2, 3, 4, 5), (1, 2, 0, 0, 2, 1), (0, 1, 2, 3, 4, 5)], dtype=float)
    for k in range(3):
        alpha, t1, pt = gbez.Bezier(c1.copy()).minDist2Poly(g["poly%d" % k])
        assert (alpha, t1) == tuple(g["poly_r"][k])
        assert np.array_equal(pt, g["poly_pt"][k])
    out, status = gbez.min_dist2poly_batch(g["rp_a"], g["rp_poly"])
    assert np.array_equal(out[:, :2], g["rp_r"]) and np.array_equal(out[:, 2:], g["rp_pt"])
    C3 = np.array([(0, 1, 2, 3, 4, 5), (0, 1, 2, 3, 4, 5), (0, 0, 0, 0, 0, 0)], dtype=float)
    C4 = np.array([(5, 4, 3, 2, 1, 0), (-1, 0, 1, 2, 3, 4), (0, 0, 0, 0, 0, 0)], dtype=float)
    C2 = np.array([(0, 1, 2, 3, 4, 5), (3, 2, 0, 0, 2, 3), (5, 4, 3, 2, 1, 0)], dtype=float)
    assert gbez.Bezier(C3).collCheck(gbez.Bezier(C4)) == 0.0          # collision
    assert gbez.Bezier(c1).collCheck(gbez.Bezier(C2)) == 1            # none
    poly = np.array([(1, 1, 3), (1, 1, 2), (1, 2, 1), (3, 1, 3), (1, 3, 1)], dtype=float)
    assert gbez.Bezier(c1 + 3).collCheck2Poly(poly) == 1
    assert np.array_equal(gbez.coll_check_batch(g["ccr_a"], g["ccr_b"]), g["ccr_r"])
    # colliding curve <-> polytope: the reference never returns; bounded here
    poly2 = np.array([(1, 1, 3), (1, 1, 2), (1, 2, 1), (3, -1, 3), (1, 3, 1)], dtype=float)
    out, status = gbez.coll_check2poly_batch(c1[None], poly2[None], max_nodes=20000)
    assert out[0] == 0 and status[0] in (0, 1)


def test_mindist_random_vs_oracle(gbez):
    from oracle import gjk_oracle as G
    rng = np.random.default_rng(123)
    count = 24
    A = np.cumsum(rng.normal(size=(count, 3, 6)), axis=2)
    B = np.cumsum(rng.normal(size=(count, 3, 6)), axis=2) + rng.normal(size=(count, 3, 1)) * 3
    out, status = gbez.min_dist_batch(A, B, max_nodes=20000)
    for k in range(count):
        alpha, t1, t2, st = G.min_dist(A[k], B[k], max_nodes=20000)
        assert st == status[k]
        if st == 0:
            assert (alpha, t1, t2) == tuple(out[k]), k
        else:
            assert np.all(np.isnan(out[k]))


def test_extrema(gbez, golden):
    from oracle import bezier_oracle as O
    g = golden("geometry_extrema")
    for row, mn, mx in zip(g["cpts"], g["mins"], g["maxs"]):
        c = row[~np.isnan(row)]
        b = gbez.Bezier(c.copy())
        assert b.min() == mn and b.max() == mx                 # reference, depth <= 1
    rng = np.random.default_rng(8)
    rows = rng.normal(size=(200, 9))
    mins, st = gbez.extrema_batch(rows, tol=1e-9)
    maxs, st2 = gbez.extrema_batch(rows, tol=1e-9, maximum=True)
    for k in range(200):
        assert mins[k] == O.bez_extreme(rows[k], tol=1e-9)
        assert maxs[k] == O.bez_extreme(rows[k], tol=1e-9, maximum=True)


def test_spatial_separation_constraints():
    """A13: optimization.py:109-133 over vehicles + shape obstacles."""
    from oracle import gjk_oracle as G
    from optimalbeziertrajectorygeneration_b200 import bezier as gbez, optimization as gopt
    rng = np.random.default_rng(2)
    # three vehicles flying in well separated lanes + one obstacle curve above them
    init = np.array([[0., 0., 0.], [0., 4., 0.], [0., 8., 0.]])
    fin = np.array([[10., 1., 1.], [10., 5., 1.], [10., 9., 1.]])
    obs = gbez.Bezier(np.array([np.linspace(0, 10, 5), np.linspace(2, 7, 5), np.full(5, 6.0)]) +
                      rng.normal(size=(3, 5)) * 0.2)
    b = gopt.BezOptimization(numVeh=3, dimension=3, degree=4, maxSep=0.5, initPoints=init, finalPoints=fin,
                             shapeObstacles=[obs])
    x = b.generateGuess(std=0.1, seed=1)
    b.spatial_on_limit = "nan"          # keep over-budget pairs as NaN rows + status (default: raise)
    got = b.spatialSeparationConstraints(x)
    assert got.shape == (6, 3)
    y = b.reshapeVector(x)
    curves = [y[i * 3:(i + 1) * 3] for i in range(3)] + [obs.cpts]
    k = 0
    for i in range(4):
        for j in range(i + 1, 4):
            a, t1, t2, st = G.min_dist(curves[i], curves[j], max_nodes=1 << 18)
            # status 1 = a path got too deep: the reference raises RecursionError there (Q6)
            assert st == b.last_status[k]
            if st == 0:
                assert np.array_equal(got[k], np.array([a, t1, t2]) - 0.5)
            else:
                assert np.all(np.isnan(got[k]))
            k += 1
    assert (b.last_status == 0).sum() >= 4
