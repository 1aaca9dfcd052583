"""TEST INFRASTRUCTURE ONLY -- wave-2 golden vectors (GJK, minDist family,
extrema) from the UNMODIFIED reference; see oracle/make_golden.py."""
import math
import sys

import numpy as np

from oracle.make_golden import Timeout, save, with_timeout

sys.setrecursionlimit(20000)


def _demo_polys():
    """gjk/gjk.py:690-755 and gjk/gjkTests.py:23-34 (dyn4j article case)."""
    P = {}
    P[1] = np.array([(4, 11, 0), (4, 5, 0), (9, 9, 0)], dtype=float)
    P[2] = np.array([(5, 6, 0), (10, 2, 0), (13, 1, 0), (12, 3, 0), (15, 6, 0)], dtype=float)
    P[3] = np.array([(4, 11, -1), (4, 5, -1), (9, 9, -1), (7, 8, 3)], dtype=float)
    P[4] = np.array([(4, 11, 3), (4, 5, 3), (9, 9, 3), (7, 8, -1)], dtype=float)
    P[5] = np.array([(4, 11, -3), (4, 5, -3), (9, 9, -3), (7, 8, -1)], dtype=float)
    P[6] = np.array([(4, 11, 0), (4, 5, 1), (9, 9, 2), (7, 8, 3)], dtype=float)
    P[7] = np.array([(-1, -1, 0), (1, 1, 0), (1, -1, 0), (-1, 1, 0)], dtype=float)
    P[8] = np.array([(-1, -1, -3), (1, 1, -3), (1, -1, -3), (-1, 1, -3), (0, 0, -1)], dtype=float)
    P[9] = np.array([(4, 11, 0), (4, 5, 0), (9, 9, 0)], dtype=float)               # dyn4j
    P[10] = np.array([(8, 6, 0), (10, 2, 0), (13, 1, 0), (15, 6, 0)], dtype=float)  # dyn4j
    return P


def section_gjk(ref):
    gjk = ref.gjk
    out = {}
    P = _demo_polys()
    demo = [(1, 2), (1, 3), (1, 4), (1, 5), (5, 6), (7, 8), (9, 10)]
    polys1, polys2, n1, n2, flags, p1s, p2s, dists = [], [], [], [], [], [], [], []

    def run(a, b):
        flag, info = with_timeout(20, gjk.gjkNew, a.copy(), b.copy())
        pa = np.full((16, 3), np.nan)
        pb = np.full((16, 3), np.nan)
        pa[:len(a)] = a
        pb[:len(b)] = b
        polys1.append(pa); polys2.append(pb); n1.append(len(a)); n2.append(len(b))
        flags.append(flag)
        if flag > 0:
            p1s.append(np.asarray(info[0], dtype=float)); p2s.append(np.asarray(info[1], dtype=float))
            dists.append(float(info[2]))
        else:
            p1s.append(np.full(3, np.nan)); p2s.append(np.full(3, np.nan)); dists.append(np.nan)

    for i, j in demo:
        run(P[i], P[j])
    rng = np.random.default_rng(42)
    skipped = 0
    for case in range(400):
        na, nb = rng.integers(2, 12, size=2)
        flat = case % 3 == 0
        a = rng.normal(size=(na, 3)) * rng.uniform(0.5, 3)
        b = rng.normal(size=(nb, 3)) * rng.uniform(0.5, 3) + rng.normal(size=3) * rng.uniform(0, 6)
        if flat:
            a[:, 2] = 0
            b[:, 2] = 0
        try:
            run(a, b)
        except (Timeout, RecursionError, Exception) as e:      # noqa
            skipped += 1
    print("gjk cases:", len(flags), "skipped:", skipped, "flags:", np.unique(flags, return_counts=True))
    save("geometry_gjk", poly1=np.array(polys1), poly2=np.array(polys2), n1=np.array(n1), n2=np.array(n2),
         flag=np.array(flags), p1=np.array(p1s), p2=np.array(p2s), dist=np.array(dists),
         ndemo=np.array(len(demo)))


def _curves():
    c = {}
    c[1] = np.array([(0, 1, 2, 3, 4, 5), (1, 2, 0, 0, 2, 1), (0, 1, 2, 3, 4, 5)], dtype=float)
    c[2] = np.array([(0, 1, 2, 3, 4, 5), (3, 2, 0, 0, 2, 3), (5, 4, 3, 2, 1, 0)], dtype=float)
    c[3] = np.array([(0, 1, 2, 3, 4, 5), (0, 1, 2, 3, 4, 5), (0, 0, 0, 0, 0, 0)], dtype=float)
    c[4] = np.array([(5, 4, 3, 2, 1, 0), (0, 1, 2, 3, 4, 5), (0, 0, 0, 0, 0, 0)], dtype=float)
    c[4][1, :] -= 1
    c[5] = c[1] - 3
    c[6] = np.array([(0, 1, 2, 3, 4, 5), (5, 0, 2, 5, 7, 5)], dtype=float)
    c[7] = np.array([(0, 1, 3, 5, 7, 7, 8, 9, 9), (0, 6, 9, 6, 8, 3, 7, 8, 3)], dtype=float)
    return c


def section_mindist(ref):
    bez = ref.bezier
    C = _curves()
    out = {}
    # Examples/MinDistBez2Bez.py:87-90, Examples/BezierUsageExamples.py:75
    named = [(3, 1), (3, 2), (3, 4), (3, 5), (1, 2)]
    A, B, R = [], [], []
    for i, j in named:
        r = with_timeout(300, bez.Bezier(C[i].copy()).minDist, bez.Bezier(C[j].copy()))
        A.append(C[i]); B.append(C[j]); R.append(np.array(r, dtype=float))
        print("minDist c%d c%d ->" % (i, j), r)
    out["named_a"] = np.array(A); out["named_b"] = np.array(B); out["named_r"] = np.array(R)

    rng = np.random.default_rng(5)
    for dim, deg, tag, count in ((3, 3, "r33", 12), (3, 5, "r35", 10), (2, 4, "r24", 10)):
        A, B, R = [], [], []
        tries = 0
        while len(R) < count and tries < 60:
            tries += 1
            a = np.cumsum(rng.normal(size=(dim, deg + 1)), axis=1)
            b = np.cumsum(rng.normal(size=(dim, deg + 1)), axis=1) + rng.normal(size=(dim, 1)) * 3
            try:
                r = with_timeout(60, bez.Bezier(a.copy()).minDist, bez.Bezier(b.copy()))
            except (Timeout, RecursionError, SystemError):
                continue
            if r[0] < 0:
                continue
            A.append(a); B.append(b); R.append(np.array(r, dtype=float))
        print(tag, "cases:", len(R), "of", tries)
        out[tag + "_a"] = np.array(A); out[tag + "_b"] = np.array(B); out[tag + "_r"] = np.array(R)

    # curve <-> polytope: Examples/3D_Plots.py:130-132, bezier.py:1794-1798,1860
    poly_3dplots_1 = np.array([(1, 3, 3), (1, 3, 2), (1, 4, 1), (3, 3, 3), (1, 5, 1)])
    poly_3dplots_3 = np.array([(1, 1, 0), (1, 3, 0), (2, 5, 0), (4, 4, 0)])
    poly_main_1 = np.array([(1, 1, 3), (1, 1, 2), (1, 2, 1), (3, 1, 3), (1, 3, 1)])
    polys = [poly_3dplots_1, poly_3dplots_3, poly_main_1]
    PR, PT = [], []
    for k, poly in enumerate(polys):
        r = with_timeout(300, bez.Bezier(C[1].copy()).minDist2Poly, poly.copy())
        print("minDist2Poly", k, r)
        PR.append(np.array([r[0], r[1]], dtype=float)); PT.append(np.asarray(r[2], dtype=float))
        out["poly%d" % k] = poly.astype(float)
    out["poly_r"] = np.array(PR); out["poly_pt"] = np.array(PT)
    A, Pp, R, Rp = [], [], [], []
    tries = 0
    while len(R) < 10 and tries < 60:
        tries += 1
        a = np.cumsum(rng.normal(size=(3, 5)), axis=1)
        poly = rng.normal(size=(5, 3)) + rng.normal(size=3) * 4
        try:
            r = with_timeout(60, bez.Bezier(a.copy()).minDist2Poly, poly.copy())
        except (Timeout, RecursionError, SystemError):
            continue
        if r[0] < 0 or np.ndim(r[2]) == 0:
            continue
        A.append(a); Pp.append(poly); R.append(np.array([r[0], r[1]], dtype=float)); Rp.append(np.asarray(r[2], dtype=float))
    print("random minDist2Poly cases:", len(R), "of", tries)
    out["rp_a"] = np.array(A); out["rp_poly"] = np.array(Pp); out["rp_r"] = np.array(R); out["rp_pt"] = np.array(Rp)

    # collision checks: Examples/BezierUsageExamples.py:103,118; bezier.py:1812-1820
    cc = []
    cc.append(float(with_timeout(300, bez.Bezier(C[3].copy()).collCheck, bez.Bezier(C[4].copy()))))
    cc.append(float(with_timeout(300, bez.Bezier(C[1].copy()).collCheck, bez.Bezier(C[2].copy()))))
    out["cc_bez"] = np.array(cc)
    out["cc_poly"] = np.array([float(with_timeout(300, bez.Bezier(C[1] + 3).collCheck2Poly, poly_main_1.astype(float)))])
    A, B, R = [], [], []
    for _ in range(12):
        a = np.cumsum(rng.normal(size=(3, 5)), axis=1)
        b = np.cumsum(rng.normal(size=(3, 5)), axis=1) + rng.normal(size=(3, 1)) * 2
        try:
            r = with_timeout(60, bez.Bezier(a.copy()).collCheck, bez.Bezier(b.copy()))
        except (Timeout, RecursionError, SystemError):
            continue
        A.append(a); B.append(b); R.append(float(r))
    out["ccr_a"] = np.array(A); out["ccr_b"] = np.array(B); out["ccr_r"] = np.array(R)
    print("collCheck random:", R)
    save("geometry_mindist", **out)


def section_extrema(ref):
    """Bezier.min / max on inputs whose recursion stops at depth <= 1 (Q4)."""
    bez = ref.bezier
    rng = np.random.default_rng(9)
    rows, mins, maxs = [], [], []
    tries = 0
    while len(rows) < 40 and tries < 4000:
        tries += 1
        deg = int(rng.integers(2, 9))
        c = rng.normal(size=deg + 1)
        kind = tries % 3
        if kind == 0:
            c = np.sort(c)                      # monotone: extrema at the ends
        elif kind == 1:
            c = np.sort(c)[::-1].copy()
        depth = [0]

        orig_split = bez.Bezier.split

        def counting_split(self, t):
            depth[0] += 1
            return orig_split(self, t)
        bez.Bezier.split = counting_split
        try:
            mn = with_timeout(5, bez.Bezier(c.copy()).min)
            d1 = depth[0]
            depth[0] = 0
            mx = with_timeout(5, bez.Bezier(c.copy()).max)
            d2 = depth[0]
        except (Timeout, RecursionError, SystemError):
            continue
        finally:
            bez.Bezier.split = orig_split
        if d1 <= 1 and d2 <= 1:
            pad = np.full(9, np.nan)
            pad[:deg + 1] = c
            rows.append(pad); mins.append(mn); maxs.append(mx)
    print("extrema cases:", len(rows), "of", tries)
    save("geometry_extrema", cpts=np.array(rows), mins=np.array(mins, dtype=float), maxs=np.array(maxs, dtype=float))


SECTIONS = {"gjk": section_gjk, "mindist": section_mindist, "extrema": section_extrema}
