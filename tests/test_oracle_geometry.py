"""Pins the pure-Python geometry oracle (oracle/gjk_oracle.py: GJK state
machine, minDist / minDist2Poly / collCheck, extrema) against golden vectors of
the UNMODIFIED reference.  CPU only."""
import numpy as np
import pytest

from oracle import bezier_oracle as O
from oracle import gjk_oracle as G


def test_gjk_golden(golden):
    g = golden("geometry_gjk")
    n = len(g["flag"])
    exact = 0
    for k in range(n):
        a = g["poly1"][k][:g["n1"][k]]
        b = g["poly2"][k][:g["n2"][k]]
        flag, p1, p2, dist = G.gjk_new(a, b)
        assert flag == g["flag"][k], k                     # bit-identical collision flags
        if flag > 0:
            # bit for bit (the oracle mirrors numba's uncontracted arithmetic and the
            # FMA chain of the BLAS dot behind ndarray.dot / np.linalg.norm)
            assert dist == g["dist"][k], k
            assert np.array_equal(p1, g["p1"][k]) and np.array_equal(p2, g["p2"][k]), k
            exact += 1
    # SURVEY section 4 known answers (gjk/gjk.py __main__ demo, dyn4j case)
    assert list(g["flag"][:7]) == [0, 0, 0, 1, 1, 1, 1]
    assert g["dist"][3] == 1.0 and g["dist"][5] == 1.0
    assert g["dist"][4] == pytest.approx(2.3426064283290913, rel=1e-15)
    assert g["dist"][6] == pytest.approx(1.7179113807746667, rel=1e-15)
    assert exact == int((g["flag"] > 0).sum())


def test_mindist_golden(golden):
    g = golden("geometry_mindist")
    for tag in ("named", "r33", "r35", "r24"):
        for a, b, r in zip(g[tag + "_a"], g[tag + "_b"], g[tag + "_r"]):
            alpha, t1, t2, status = G.min_dist(a, b)
            assert status == 0
            assert (alpha, t1, t2) == tuple(r), tag          # bit-identical (alpha, t1, t2)
    # SURVEY section 4 goldens (Examples/MinDistBez2Bez.py, BezierUsageExamples.py)
    want = [(1.0, 0.0, 0.0), (1.4142135623789327, 0.7999994253499804, 1.0),
            (6.585445079830187e-10, 0.39999999990686774, 0.6),
            (2.173752805053979, 0.0, 0.6699547765929507), (0.125, 0.5, 0.5)]
    for r, w in zip(g["named_r"], want):
        assert tuple(r) == pytest.approx(w, rel=1e-13)


def test_mindist2poly_golden(golden):
    g = golden("geometry_mindist")
    c1 = np.array([(0, 1, 2, 3, 4, 5), (1, 2, 0, 0, 2, 1), (0, 1, 2, 3, 4, 5)], dtype=float)
    for k in range(3):
        alpha, t1, pt, status = G.min_dist2poly(c1, g["poly%d" % k])
        assert status == 0
        assert alpha == pytest.approx(g["poly_r"][k][0], rel=1e-11)
        assert t1 == pytest.approx(g["poly_r"][k][1], rel=1e-9)
        np.testing.assert_allclose(pt, g["poly_pt"][k], rtol=1e-9, atol=1e-12)
    for a, poly, r, pt in zip(g["rp_a"], g["rp_poly"], g["rp_r"], g["rp_pt"]):
        alpha, t1, p, status = G.min_dist2poly(a, poly)
        assert alpha == pytest.approx(r[0], rel=1e-11)
        assert t1 == pytest.approx(r[1], rel=1e-9, abs=1e-12)
        np.testing.assert_allclose(p, pt, rtol=1e-9, atol=1e-12)


def test_collcheck_golden(golden):
    g = golden("geometry_mindist")
    C3 = np.array([(0, 1, 2, 3, 4, 5), (0, 1, 2, 3, 4, 5), (0, 0, 0, 0, 0, 0)], dtype=float)
    C4 = np.array([(5, 4, 3, 2, 1, 0), (-1, 0, 1, 2, 3, 4), (0, 0, 0, 0, 0, 0)], dtype=float)
    C1 = np.array([(0, 1, 2, 3, 4, 5), (1, 2, 0, 0, 2, 1), (0, 1, 2, 3, 4, 5)], dtype=float)
    C2 = np.array([(0, 1, 2, 3, 4, 5), (3, 2, 0, 0, 2, 3), (5, 4, 3, 2, 1, 0)], dtype=float)
    assert G.coll_check_bez2bez(C3, C4) == g["cc_bez"][0] == 0.0
    assert G.coll_check_bez2bez(C1, C2) == g["cc_bez"][1] == 1
    poly = np.array([(1, 1, 3), (1, 1, 2), (1, 2, 1), (3, 1, 3), (1, 3, 1)], dtype=float)
    assert G.coll_check_bez2poly(C1 + 3, poly) == (int(g["cc_poly"][0]), 0)
    for a, b, r in zip(g["ccr_a"], g["ccr_b"], g["ccr_r"]):
        assert G.coll_check_bez2bez(a, b) == r


def test_extrema_depth1_golden(golden):
    g = golden("geometry_extrema")
    for row, mn, mx in zip(g["cpts"], g["mins"], g["maxs"]):
        c = row[~np.isnan(row)]
        assert O.bez_extreme(c) == mn
        assert O.bez_extreme(c, maximum=True) == mx


def test_extrema_intended_algorithm_brackets_true_extremum():
    """Beyond depth 1 the reference is defective; the restated
    algorithm must bracket the sampled extremum within its tolerance."""
    rng = np.random.default_rng(3)
    for _ in range(30):
        c = rng.normal(size=int(rng.integers(3, 10)))
        t = np.linspace(0, 1, 4001)
        vals = O.de_casteljau_eval(c[None, :], t)[0]
        mn = O.bez_extreme(c, tol=1e-9)
        mx = O.bez_extreme(c, tol=1e-9, maximum=True)
        assert mn <= vals.min() + 1e-9 and mn >= vals.min() - 1e-3
        assert mx >= vals.max() - 1e-9 and mx <= vals.max() + 1e-3
