// Geometry kernels (A8-A12 of SURVEY.md section 8) for sm_100a: de Casteljau split,
// extrema, GJK, minDist / minDist2Poly / collCheck.
//
// COMPILED WITH -fmad=false.  The reference's numba helpers are not FMA-contracted
//  and np.cross is separate multiplies and subtracts, so every a*b+c
// below must stay two roundings.  The places where the reference goes through
// ndarray.dot / np.linalg.norm on 3-vectors use BLAS ddot, whose scalar tail loop is
// contracted -- those use the explicit fma() chain of dot_blas().  With that the
// kernels reproduce the reference's (flag, points, distance) and (alpha, t1, t2)
// bit for bit on the golden vectors.
//
// Execution model: ONE WARP PER ITEM (curve pair / curve / polygon pair).  Lane i
// owns control point i of each curve (n+1 <= 32).  The data-parallel pieces are
// warp collectives:
//   support()           argmax of <p_i, d> with first-index tie break  (shuffle reduction)
//   split               one de Casteljau level per step via __shfl_down
//   control-point match ballot
//   distance weights    every lane computes its own W[i]
// while the GJK simplex state machine and the branch-and-bound bookkeeping are
// warp-uniform scalars.  Recursion of the reference becomes an explicit DFS stack in
// global memory (bounded depth -> status flag instead of RecursionError).
#include <math.h>

#include "common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kGeomThreads = 128;            // 4 warps (items) per block

struct V3 { double x, y, z; };

__device__ __forceinline__ V3 mk(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 vneg(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 vscale(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ bool veq(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// gjk/gjk.py:174-194 (numba, not contracted)
__device__ __forceinline__ double dot_plain(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// ndarray.dot / np.linalg.norm through BLAS: fma(a2,b2, fma(a1,b1, a0*b0))
__device__ __forceinline__ double dot_blas(V3 a, V3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ double norm_blas(V3 a) { return sqrt(dot_blas(a, a)); }
// np.cross
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 lerp(double t, V3 p, V3 q) { return vadd(vscale(1 - t, p), vscale(t, q)); }
__device__ __forceinline__ V3 shfl3(V3 v, int src) {
    return mk(__shfl_sync(FULL, v.x, src), __shfl_sync(FULL, v.y, src), __shfl_sync(FULL, v.z, src));
}

// A point cloud spread over the warp: lane i holds point i (i < n).
struct Cloud { V3 p; int n; };

// gjk/gjk.py:87-114: first strict maximum wins == smallest index among the maxima.
__device__ V3 support(const Cloud &c, V3 d, int lane) {
    double v = (lane < c.n) ? dot_plain(c.p, d) : -INFINITY;
    int idx = lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(FULL, v, off);
        const int oi = __shfl_xor_sync(FULL, idx, off);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    return shfl3(c.p, idx);
}

struct SPoint { V3 m, s1, s2; };            // Minkowski point + its two source points

// gjk/gjk.py:493-501
__device__ SPoint support_pts(const Cloud &c1, const Cloud &c2, V3 d, int lane) {
    SPoint r;
    r.s1 = support(c1, d, lane);
    r.s2 = support(c2, vneg(d), lane);
    r.m = vsub(r.s1, r.s2);
    return r;
}

// gjk/gjk.py:397-437
__device__ void weighted_origin_to_line(V3 A, V3 B, double &t, double &dist) {
    if (veq(A, B)) { t = 0.0; dist = sqrt(dot_plain(A, A)); return; }
    const V3 v = vsub(B, A);
    t = -dot_plain(v, A) / dot_plain(v, v);
    if (t > 1) t = 1; else if (t < 0) t = 0;
    const V3 c = lerp(t, A, B);
    dist = sqrt(dot_plain(c, c));
}

// gjk/gjk.py:440-477
__device__ void weighted_origin_to_plane(V3 A, V3 B, V3 C, double bary[3], double &dist) {
    const V3 N = cross(vsub(B, A), vsub(C, A));
    const double nn = norm_blas(N);
    const V3 n = mk(N.x / nn, N.y / nn, N.z / nn);
    const double t = (n.x * A.x + n.y * A.y + n.z * A.z) / (n.x * n.x + n.y * n.y + n.z * n.z);
    const V3 cl = vscale(t, n);
    dist = sqrt(dot_plain(cl, cl));
    const V3 PA = vsub(A, cl), PB = vsub(B, cl), PC = vsub(C, cl);
    const double area = norm_blas(N);
    bary[0] = norm_blas(cross(PB, PC)) / area;
    bary[1] = norm_blas(cross(PC, PA)) / area;
    bary[2] = 1 - bary[0] - bary[1];
}

struct Simplex {
    SPoint A, B, C, D;
    bool hasA, hasB, hasC, hasD, collision;
};
__device__ __forceinline__ void sclear(Simplex &s) { s.hasA = s.hasB = s.hasC = s.hasD = s.collision = false; }

// gjk/gjk.py:564-642
__device__ V3 simplex3(const Cloud &c1, const Cloud &c2, Simplex &s, int lane) {
    const V3 A = s.A.m, B = s.B.m, C = s.C.m;
    const V3 A0 = vneg(A), AB = vsub(B, A), AC = vsub(C, A);
    const V3 ABC = cross(AB, AC);
    V3 d;
    if (dot_blas(cross(ABC, AC), A0) > 0) {
        if (dot_blas(AC, A0) > 0) {
            d = cross(cross(AC, A0), AC);
            s.B = s.A;
        } else if (dot_blas(AB, A0) > 0) {
            d = cross(cross(AB, A0), AB);
            s.C = s.A;
        } else {
            d = A;                               // '+A' 
            sclear(s);
        }
    } else if (dot_blas(cross(AB, ABC), A0) > 0) {
        if (dot_blas(AB, A0) > 0) {
            d = cross(cross(AB, A0), AB);
            s.C = s.A;
        } else {
            d = vneg(A);
            sclear(s);
        }
    } else {
        const double h = dot_blas(ABC, A0);
        if (h == 0) {
            s.collision = true;
            d = mk(0.0, 0.0, 0.0);
        } else if (h > 0) {
            d = ABC;
            s.D = s.C; s.hasD = true;
            s.C = s.B;
            s.B = s.A;
        } else {
            d = vneg(ABC);
            s.D = s.B; s.hasD = true;
            s.B = s.A;
        }
    }
    s.A = support_pts(c1, c2, d, lane);
    s.hasA = true;
    return d;
}

// gjk/gjk.py:645-681
__device__ V3 simplex4(const Cloud &c1, const Cloud &c2, Simplex &s, int lane) {
    const V3 A = s.A.m;
    const V3 A0 = vneg(A);
    const V3 AB = vsub(s.B.m, A), AC = vsub(s.C.m, A), AD = vsub(s.D.m, A);
    const V3 ABC = cross(AB, AC), ACD = cross(AC, AD), ADB = cross(AD, AB);
    if (dot_blas(ABC, A0) > 0) {
        s.hasD = false;
        return simplex3(c1, c2, s, lane);
    }
    if (dot_blas(ACD, A0) > 0) {
        s.B = s.C; s.C = s.D; s.hasD = false;
        return simplex3(c1, c2, s, lane);
    }
    if (dot_blas(ADB, A0) > 0) {
        s.C = s.B; s.B = s.D; s.hasD = false;
        return simplex3(c1, c2, s, lane);
    }
    s.collision = true;
    return mk(0.0, 0.0, 0.0);
}

// gjk/gjk.py:504-561
__device__ V3 do_simplex(const Cloud &c1, const Cloud &c2, Simplex &s, V3 d, int lane) {
    if (!s.hasA) {
        s.A = support_pts(c1, c2, d, lane); s.hasA = true;
        return d;
    }
    if (!s.hasB) {
        s.B = s.A; s.hasB = true;
        d = vneg(d);
        s.A = support_pts(c1, c2, d, lane);
        return d;
    }
    if (!s.hasC) {
        double t, dist;
        weighted_origin_to_line(s.A.m, s.B.m, t, dist);
        d = vneg(lerp(t, s.A.m, s.B.m));
        s.C = s.A; s.hasC = true;
        s.A = support_pts(c1, c2, d, lane);
        return d;
    }
    if (!s.hasD) return simplex3(c1, c2, s, lane);
    return simplex4(c1, c2, s, lane);
}

struct GjkResult { int flag; V3 p1, p2; double dist; };

// gjkNew + minimumDistance (gjk/gjk.py:229-360).  flag: 1 distance available,
// 0 collision, -1 iteration limit (also used when minimumDistance does not
// converge within max_md iterations, where the reference would spin forever).
__device__ GjkResult gjk_new(const Cloud &c1, const Cloud &c2, int lane, int max_iter = 128,
                             int max_md = 4096) {
    GjkResult r;
    r.flag = -1; r.p1 = r.p2 = mk(0, 0, 0); r.dist = 0.0;
    Simplex s;
    sclear(s);
    V3 d = mk(1.0, 0.0, 0.0);
    for (int it = 0; it < max_iter; ++it) {
        d = do_simplex(c1, c2, s, d, lane);
        if (s.collision) { r.flag = 0; return r; }
        if (dot_blas(s.A.m, d) < 0) {
            // ---- minimumDistance ----
            bool converged = false;
            for (int k = 0; k < max_md; ++k) {
                const Simplex old = s;
                d = do_simplex(c1, c2, s, d, lane);
                const V3 a = s.A.m;
                if ((old.hasA && veq(a, old.A.m)) || (old.hasB && veq(a, old.B.m)) ||
                    (old.hasC && veq(a, old.C.m)) || (old.hasD && veq(a, old.D.m))) {
                    s = old;
                    converged = true;
                    break;
                }
            }
            if (!converged) return r;
            r.flag = 1;
            if (s.hasC) {
                const V3 A = s.A.m, B = s.B.m, C = s.C.m;
                const V3 A0 = vneg(A), AB = vsub(B, A), AC = vsub(C, A);
                const V3 ABC = cross(AB, AC);
                double t;
                if (dot_blas(cross(ABC, AC), A0) >= 0) {
                    weighted_origin_to_line(A, C, t, r.dist);
                    r.p1 = lerp(t, s.A.s1, s.C.s1);
                    r.p2 = lerp(t, s.A.s2, s.C.s2);
                } else if (dot_blas(cross(AB, ABC), A0) >= 0) {
                    weighted_origin_to_line(A, B, t, r.dist);
                    r.p1 = lerp(t, s.A.s1, s.B.s1);
                    r.p2 = lerp(t, s.A.s2, s.B.s2);
                } else {
                    double b[3];
                    weighted_origin_to_plane(A, B, C, b, r.dist);
                    r.p1 = vadd(vadd(vscale(b[0], vadd(A, s.A.s2)), vscale(b[1], vadd(B, s.B.s2))),
                                vscale(b[2], vadd(C, s.C.s2)));
                    r.p2 = vadd(vadd(vscale(b[0], vsub(s.A.s1, A)), vscale(b[1], vsub(s.B.s1, B))),
                                vscale(b[2], vsub(s.C.s1, C)));
                }
            } else if (s.hasB) {
                double t;
                weighted_origin_to_line(s.A.m, s.B.m, t, r.dist);
                r.p1 = lerp(t, s.A.s1, s.B.s1);
                r.p2 = lerp(t, s.A.s2, s.B.s2);
            } else {
                r.dist = norm_blas(s.A.m);
                r.p1 = s.A.s1;
                r.p2 = s.A.s2;
            }
            return r;
        }
    }
    return r;
}

// ---------------------------------------------------------------------------
// numpy's pairwise add.reduce for < 128 contiguous doubles (loops_utils.h.src)
__device__ double np_sum(const double *v, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += v[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = v[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += v[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += v[i];
    return res;
}

// bezier.py:1320-1337: parameter of the closest point on the control polygon.
__device__ double param_estimate(const Cloud &c, V3 closest, int lane) {
    const int N = c.n, deg = c.n - 1;
    const unsigned hit = __ballot_sync(FULL, lane < N && veq(c.p, closest));
    if (hit) return (double)(__ffs(hit) - 1) / (double)deg;
    const V3 dv = vsub(closest, c.p);
    const double e = sqrt((0.0 + dv.x * dv.x) + dv.y * dv.y + dv.z * dv.z);
    double q[32], W[32];
    // every lane gathers all distances, computes its own weight, then all weights
    double ei[32];
    for (int j = 0; j < N; ++j) ei[j] = __shfl_sync(FULL, e, j);
    const int i = lane < N ? lane : 0;
    int cnt = 0;
    for (int j = 0; j < i; ++j) q[cnt++] = ei[i] / ei[j];
    const double s1 = np_sum(q, cnt);
    cnt = 0;
    for (int j = i + 1; j < N; ++j) q[cnt++] = ei[i] / ei[j];
    const double s2 = np_sum(q, cnt);
    const double w = 1.0 / ((1 + s1) + s2);
    for (int j = 0; j < N; ++j) W[j] = (__shfl_sync(FULL, w, j) * (double)j) / (double)N;
    return np_sum(W, N);
}

// _norm (bezier.py:1550-1558)
__device__ __forceinline__ double norm_numba(V3 v) { return sqrt(((0.0 + v.x * v.x) + v.y * v.y) + v.z * v.z); }

// _upperbound (bezier.py:1499-1516): first minimum of the 4 end-point distances
__device__ void upperbound(const Cloud &a, const Cloud &b, double &ub, double &u1, double &u2) {
    const V3 a0 = shfl3(a.p, 0), a1 = shfl3(a.p, a.n - 1), b0 = shfl3(b.p, 0), b1 = shfl3(b.p, b.n - 1);
    const double d0 = norm_numba(vsub(a0, b0)), d1 = norm_numba(vsub(a0, b1));
    const double d2 = norm_numba(vsub(a1, b0)), d3 = norm_numba(vsub(a1, b1));
    ub = d0; u1 = 0.0; u2 = 0.0;
    if (d1 < ub) { ub = d1; u1 = 0.0; u2 = 1.0; }
    if (d2 < ub) { ub = d2; u1 = 1.0; u2 = 0.0; }
    if (d3 < ub) { ub = d3; u1 = 1.0; u2 = 1.0; }
}

// Bezier.split -> deCasteljauSplit (bezier.py:533-572, 985-1027) on a lane-spread
// coordinate: returns this lane's control point of the left and right halves.
__device__ void split_coord(double c, int n1, double t, int lane, double &left, double &right) {
    const int n = n1 - 1;
    left = c; right = c;
    for (int lvl = 0; lvl <= n; ++lvl) {
        const double c0 = __shfl_sync(FULL, c, 0);
        if (lane == lvl) left = c0;
        if (lane == n - lvl) right = c;
        const double nxt = __shfl_down_sync(FULL, c, 1);
        if (lane < n - lvl) c = (1 - t) * c + t * nxt;
    }
}
__device__ void split_cloud(const Cloud &c, double t, int lane, Cloud &l, Cloud &r) {
    if (isnan(t)) t = 0;                                   // bezier.py:555-557
    l.n = r.n = c.n;
    split_coord(c.p.x, c.n, t, lane, l.p.x, r.p.x);
    split_coord(c.p.y, c.n, t, lane, l.p.y, r.p.y);
    split_coord(c.p.z, c.n, t, lane, l.p.z, r.p.z);
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ Cloud load_cloud_rows(const double *cpts, int dim, int n1, int lane) {
    // [dim][n1] row-major curve -> lane-spread 3-D points (2-D curves padded with z = 0)
    Cloud c;
    c.n = n1;
    const bool in = lane < n1;
    c.p.x = in ? cpts[lane] : 0.0;
    c.p.y = (in && dim > 1) ? cpts[n1 + lane] : 0.0;
    c.p.z = (in && dim > 2) ? cpts[2 * n1 + lane] : 0.0;
    return c;
}
__device__ __forceinline__ Cloud load_cloud_points(const double *pts, int n, int lane) {
    Cloud c;   // [n][3] point list
    c.n = n;
    const bool in = lane < n;
    c.p.x = in ? pts[3 * lane] : 0.0;
    c.p.y = in ? pts[3 * lane + 1] : 0.0;
    c.p.z = in ? pts[3 * lane + 2] : 0.0;
    return c;
}

// ------------------------------ kernels ------------------------------------
__global__ void gjk_kernel(const double *poly1, const double *poly2, const int *n1, const int *n2,
                           int n1max, int n2max, int count, int *flag, double *p1, double *p2,
                           double *dist) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= count) return;
    const Cloud c1 = load_cloud_points(poly1 + (size_t)item * n1max * 3, n1 ? n1[item] : n1max, lane);
    const Cloud c2 = load_cloud_points(poly2 + (size_t)item * n2max * 3, n2 ? n2[item] : n2max, lane);
    const GjkResult r = gjk_new(c1, c2, lane);
    if (lane == 0) {
        flag[item] = r.flag;
        const double nanv = nan("");
        const bool ok = r.flag > 0;
        p1[3 * item] = ok ? r.p1.x : nanv; p1[3 * item + 1] = ok ? r.p1.y : nanv; p1[3 * item + 2] = ok ? r.p1.z : nanv;
        p2[3 * item] = ok ? r.p2.x : nanv; p2[3 * item + 1] = ok ? r.p2.y : nanv; p2[3 * item + 2] = ok ? r.p2.z : nanv;
        dist[item] = ok ? r.dist : nanv;
    }
}

__global__ void split_kernel(const double *cpts, const double *tloc, int count, int dim, int n1,
                             double *left, double *right) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= count) return;
    double t = tloc[item];
    if (isnan(t)) t = 0;
    for (int d = 0; d < dim; ++d) {
        const size_t off = ((size_t)item * dim + d) * n1;
        const double c = lane < n1 ? cpts[off + lane] : 0.0;
        double l, r;
        split_coord(c, n1, t, lane, l, r);
        if (lane < n1) { left[off + lane] = l; right[off + lane] = r; }
    }
}

// Bezier.min / Bezier.max (bezier.py:631-667, 727-763), intended algorithm: split at
// the extreme control point (local parameter idx/deg) until an end point is extreme
// or the bound moves by < tol.  Explicit DFS stack of rows in global scratch.
__global__ void extrema_kernel(const double *rows, int count, int n1, double tol, int maximum,
                               int max_depth, double *scratch, double *out, int *status) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= count) return;
    const double sgn = maximum ? -1.0 : 1.0;
    const int deg = n1 - 1;
    // frame: [n1 doubles row][glob][stage][best]  -> n1 + 3 doubles
    const int fsz = n1 + 3;
    double *st = scratch + (size_t)item * max_depth * fsz;
    int sp = 0, stat = 0;
    if (lane < n1) st[lane] = rows[(size_t)item * n1 + lane];
    if (lane == 0) { st[n1] = maximum ? INFINITY : -INFINITY; st[n1 + 1] = 0.0; st[n1 + 2] = 0.0; }
    __syncwarp();
    double ret = 0.0;
    while (sp >= 0) {
        double *f = st + (size_t)sp * fsz;
        const int stage = (int)f[n1 + 1];
        const double c = lane < n1 ? f[lane] : 0.0;
        if (stage == 0) {
            // argmin of sgn*c, first occurrence (np.argmin / np.argmax)
            double v = lane < n1 ? sgn * c : INFINITY;
            int idx = lane;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(FULL, v, off);
                const int oi = __shfl_xor_sync(FULL, idx, off);
                if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
            }
            const double ext = __shfl_sync(FULL, c, idx);
            const double glob = f[n1];
            const bool leaf = fabs(glob - ext) < tol || idx == 0 || idx == deg || sp + 1 >= max_depth;
            if (!(fabs(glob - ext) < tol) && idx != 0 && idx != deg && sp + 1 >= max_depth) stat = 1;
            if (leaf) {
                ret = ext;
                --sp;
                __syncwarp();
                continue;
            }
            double l, r;
            split_coord(c, n1, (double)idx / (double)deg, lane, l, r);
            __syncwarp();
            // keep the right half in this frame (visited second), push the left half
            double *g = f + fsz;
            if (lane < n1) { f[lane] = r; g[lane] = l; }
            if (lane == 0) { f[n1] = ext; f[n1 + 1] = 1.0; g[n1] = ext; g[n1 + 1] = 0.0; g[n1 + 2] = 0.0; }
            ++sp;
            __syncwarp();
        } else if (stage == 1) {               // left child returned in `ret`
            double *g = f + fsz;
            if (lane == 0) { f[n1 + 2] = ret; f[n1 + 1] = 2.0; g[n1] = f[n1]; g[n1 + 1] = 0.0; g[n1 + 2] = 0.0; }
            if (lane < n1) g[lane] = f[lane];   // right half becomes the child
            ++sp;
            __syncwarp();
        } else {                                // right child returned
            const double a = f[n1 + 2];
            ret = maximum ? (a > ret ? a : ret) : (a < ret ? a : ret);
            --sp;
            __syncwarp();
        }
    }
    if (lane == 0) { out[item] = ret; status[item] = stat; }
}

// --------------------------- minDist family --------------------------------
// DFS frame of _minDist (bezier.py:1283-1408): the two parent curves, their
// parameter windows, the split parameters and the running best.  Children are
// re-derived from the parent by the (deterministic) de Casteljau split when they
// are visited, in the reference's order (c3,c5) (c3,c6) (c4,c5) (c4,c6).
struct MDFrame {
    double t1l, t1h, t2l, t2h, t1, t2;
    double ra, rt1, rt2;        // local retval (alpha, t1, t2)
    int stage;
};

__device__ __forceinline__ void store_cloud(double *dst, const Cloud &c, int lane) {
    if (lane < c.n) { dst[3 * lane] = c.p.x; dst[3 * lane + 1] = c.p.y; dst[3 * lane + 2] = c.p.z; }
}

__global__ void mindist_kernel(const double *c1, const double *c2, int count, int dim1, int dim2,
                               int n1a, int n1b, double eps, int max_depth, long long max_nodes,
                               double *scratch, double *out, int *status) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= count) return;
    const int csz = 3 * (n1a + n1b);
    const size_t per_item = (size_t)max_depth * (csz + sizeof(MDFrame) / sizeof(double) + 1);
    double *cs = scratch + (size_t)item * per_item;                       // curves per level
    MDFrame *fr = reinterpret_cast<MDFrame *>(cs + (size_t)max_depth * csz);
    int sp = 0, stat = 0;
    {
        const Cloud a = load_cloud_rows(c1 + (size_t)item * dim1 * n1a, dim1, n1a, lane);
        const Cloud b = load_cloud_rows(c2 + (size_t)item * dim2 * n1b, dim2, n1b, lane);
        store_cloud(cs, a, lane);
        store_cloud(cs + 3 * n1a, b, lane);
        if (lane == 0) { fr[0].t1l = 0; fr[0].t1h = 1; fr[0].t2l = 0; fr[0].t2h = 1; fr[0].stage = 0; fr[0].ra = INFINITY; }
    }
    __syncwarp();
    double ra = 0, rt1 = 0, rt2 = 0;            // value returned by the call being unwound
    long long nodes = 0;
    while (sp >= 0) {
        MDFrame &f = fr[sp];
        double *cur = cs + (size_t)sp * csz;
        const int stage = f.stage;
        if (stage == 0) {
            if (++nodes > max_nodes) { stat |= 4; break; }     // node budget exhausted
            const Cloud a = load_cloud_points(cur, n1a, lane);
            const Cloud b = load_cloud_points(cur + 3 * n1a, n1b, lane);
            const double alpha_in = f.ra;
            const GjkResult g = gjk_new(a, b, lane);
            double t1, t2, lb;
            if (g.flag > 0) {
                lb = g.dist;
                t1 = param_estimate(a, g.p1, lane);
                t2 = param_estimate(b, g.p2, lane);
            } else {
                t1 = 0.5; t2 = 0.5; lb = eps;
            }
            double ub, u1, u2;
            upperbound(a, b, ub, u1, u2);
            double al = alpha_in, nT1 = -1, nT2 = -1;
            if (ub <= al) {
                al = ub;
                nT1 = (1 - u1) * f.t1l + u1 * f.t1h;
                nT2 = (1 - u2) * f.t2l + u2 * f.t2h;
            }
            if (lb >= al * (1 - eps)) {
                ra = al; rt1 = nT1; rt2 = nT2;
                --sp;
                __syncwarp();
                continue;
            }
            // the reference raises RecursionError as soon as one path gets too deep
            // (SURVEY Q6): abort the whole search and flag it
            if (sp + 1 >= max_depth) { stat |= 1; break; }
            if (lane == 0) { f.t1 = t1; f.t2 = t2; f.ra = al; f.rt1 = nT1; f.rt2 = nT2; f.stage = 1; }
            __syncwarp();
        } else {
            // returning from child number stage-1 ... (stage 1 = about to visit child 0)
            if (stage > 1) {
                if (ra < f.ra) { if (lane == 0) { f.ra = ra; f.rt1 = rt1; f.rt2 = rt2; } }
                __syncwarp();
            }
            if (stage > 4) {
                ra = f.ra; rt1 = f.rt1; rt2 = f.rt2;
                --sp;
                __syncwarp();
                continue;
            }
            const int child = stage - 1;                     // 0..3
            const Cloud a = load_cloud_points(cur, n1a, lane);
            const Cloud b = load_cloud_points(cur + 3 * n1a, n1b, lane);
            Cloud l1, r1, l2, r2;
            split_cloud(a, f.t1, lane, l1, r1);
            split_cloud(b, f.t2, lane, l2, r2);
            const double t1len = f.t1h - f.t1l, t2len = f.t2h - f.t2l;
            const double m1 = f.t1l + f.t1 * t1len, m2 = f.t2l + f.t2 * t2len;
            double *nxt = cur + csz;
            store_cloud(nxt, (child < 2) ? l1 : r1, lane);
            store_cloud(nxt + 3 * n1a, (child & 1) ? r2 : l2, lane);
            if (lane == 0) {
                MDFrame &g = fr[sp + 1];
                g.t1l = (child < 2) ? f.t1l : m1; g.t1h = (child < 2) ? m1 : f.t1h;
                g.t2l = (child & 1) ? m2 : f.t2l; g.t2h = (child & 1) ? f.t2h : m2;
                g.stage = 0;
                g.ra = f.ra;                                  // alpha = retval[0]
                f.stage = stage + 1;
            }
            ++sp;
            __syncwarp();
        }
    }
    if (lane == 0) {
        const double nanv = nan("");
        out[3 * item] = stat ? nanv : ra; out[3 * item + 1] = stat ? nanv : rt1; out[3 * item + 2] = stat ? nanv : rt2;
        status[item] = stat;
    }
}

// _minDist2Poly (bezier.py:1411-1496, 1535-1547)
struct MPFrame { double t1l, t1h, t1, ra, rt1, rx, ry, rz; int stage; int has_pt; };

__global__ void mindist2poly_kernel(const double *c1, const double *polys, const int *npoly, int count,
                                    int dim1, int n1a, int mmax, double eps, int max_depth,
                                    long long max_nodes, double *scratch, double *out, int *status) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= count) return;
    const int csz = 3 * n1a;
    const size_t per_item = (size_t)max_depth * (csz + sizeof(MPFrame) / sizeof(double) + 1);
    double *cs = scratch + (size_t)item * per_item;
    MPFrame *fr = reinterpret_cast<MPFrame *>(cs + (size_t)max_depth * csz);
    const Cloud poly = load_cloud_points(polys + (size_t)item * mmax * 3, npoly ? npoly[item] : mmax, lane);
    int sp = 0, stat = 0;
    {
        const Cloud a = load_cloud_rows(c1 + (size_t)item * dim1 * n1a, dim1, n1a, lane);
        store_cloud(cs, a, lane);
        if (lane == 0) { fr[0].t1l = 0; fr[0].t1h = 1; fr[0].stage = 0; fr[0].ra = INFINITY; }
    }
    __syncwarp();
    double ra = 0, rt1 = 0, rx = 0, ry = 0, rz = 0;
    int rhas = 0;
    long long nodes = 0;
    while (sp >= 0) {
        MPFrame &f = fr[sp];
        double *cur = cs + (size_t)sp * csz;
        const int stage = f.stage;
        if (stage == 0) {
            if (++nodes > max_nodes) { stat |= 4; break; }
            const Cloud a = load_cloud_points(cur, n1a, lane);
            const GjkResult g = gjk_new(a, poly, lane);
            double al = f.ra, t1, lb, nT1 = -1;
            V3 cp = mk(-1, -1, -1);
            int has = 0;
            if (g.flag > 0) {
                lb = g.dist;
                cp = g.p2; has = 1;
                t1 = param_estimate(a, g.p1, lane);
                // _upperboundPoly: end points of the curve against the closest polytope point
                const V3 a0 = shfl3(a.p, 0), a1 = shfl3(a.p, a.n - 1);
                const double d0 = norm_numba(vsub(a0, g.p2)), d1 = norm_numba(vsub(a1, g.p2));
                double ub = d0, u1 = 0.0;
                if (d1 < ub) { ub = d1; u1 = 1.0; }
                if (ub <= al) { al = ub; nT1 = (1 - u1) * f.t1l + u1 * f.t1h; }
            } else {
                t1 = 0.5; lb = eps * eps * eps;
            }
            if (lb >= al * (1 - eps)) {
                ra = al; rt1 = nT1; rx = cp.x; ry = cp.y; rz = cp.z; rhas = has;
                --sp;
                __syncwarp();
                continue;
            }
            if (sp + 1 >= max_depth) { stat |= 1; break; }
            if (lane == 0) { f.t1 = t1; f.ra = al; f.rt1 = nT1; f.rx = cp.x; f.ry = cp.y; f.rz = cp.z; f.has_pt = has; f.stage = 1; }
            __syncwarp();
        } else {
            if (stage > 1) {
                if (ra < f.ra) { if (lane == 0) { f.ra = ra; f.rt1 = rt1; f.rx = rx; f.ry = ry; f.rz = rz; f.has_pt = rhas; } }
                __syncwarp();
            }
            if (stage > 2) {
                ra = f.ra; rt1 = f.rt1; rx = f.rx; ry = f.ry; rz = f.rz; rhas = f.has_pt;
                --sp;
                __syncwarp();
                continue;
            }
            const int child = stage - 1;
            const Cloud a = load_cloud_points(cur, n1a, lane);
            Cloud l1, r1;
            split_cloud(a, f.t1, lane, l1, r1);
            const double m1 = f.t1l + f.t1 * (f.t1h - f.t1l);
            store_cloud(cur + csz, child == 0 ? l1 : r1, lane);
            if (lane == 0) {
                MPFrame &g = fr[sp + 1];
                g.t1l = child == 0 ? f.t1l : m1; g.t1h = child == 0 ? m1 : f.t1h;
                g.stage = 0; g.ra = f.ra;
                f.stage = stage + 1;
            }
            ++sp;
            __syncwarp();
        }
    }
    if (lane == 0) {
        double *o = out + (size_t)item * 5;
        const double nanv = nan("");
        o[0] = stat ? nanv : ra; o[1] = stat ? nanv : rt1;
        o[2] = rhas ? rx : -1; o[3] = rhas ? ry : -1; o[4] = rhas ? rz : -1;
        status[item] = stat | (rhas ? 0 : 2);
    }
}

// _collCheckBez2Bez (bezier.py:1561-1615).  Frame: window-free; alpha threaded.
struct CCFrame { double alpha; int stage; int cnt; };

__global__ void collcheck_kernel(const double *c1, const double *c2, int count, int dim1, int dim2,
                                 int n1a, int n1b, double eps, double *scratch, double *out) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= count) return;
    const int max_depth = 102;
    const int csz = 3 * (n1a + n1b);
    const size_t per_item = (size_t)max_depth * (csz + sizeof(CCFrame) / sizeof(double) + 1);
    double *cs = scratch + (size_t)item * per_item;
    CCFrame *fr = reinterpret_cast<CCFrame *>(cs + (size_t)max_depth * csz);
    int sp = 0;
    {
        const Cloud a = load_cloud_rows(c1 + (size_t)item * dim1 * n1a, dim1, n1a, lane);
        const Cloud b = load_cloud_rows(c2 + (size_t)item * dim2 * n1b, dim2, n1b, lane);
        store_cloud(cs, a, lane);
        store_cloud(cs + 3 * n1a, b, lane);
        if (lane == 0) { fr[0].alpha = INFINITY; fr[0].stage = 0; fr[0].cnt = 0; }
    }
    __syncwarp();
    double ret = 0;
    while (sp >= 0) {
        CCFrame &f = fr[sp];
        double *cur = cs + (size_t)sp * csz;
        const int stage = f.stage;
        if (stage == 0) {
            const int cnt = f.cnt + 1;
            if (cnt > 100) { ret = -1; --sp; __syncwarp(); continue; }
            const Cloud a = load_cloud_points(cur, n1a, lane);
            const Cloud b = load_cloud_points(cur + 3 * n1a, n1b, lane);
            double ub, u1, u2;
            upperbound(a, b, ub, u1, u2);
            const GjkResult g = gjk_new(a, b, lane);
            if (g.flag > 0) { ret = 1; --sp; __syncwarp(); continue; }
            double al = f.alpha;
            if (ub <= al) al = ub;
            if (0 >= al * (1 - eps)) { ret = al; --sp; __syncwarp(); continue; }
            if (lane == 0) { f.alpha = al; f.cnt = cnt; f.stage = 1; }
            __syncwarp();
        } else {
            if (stage > 1) {
                if (lane == 0) f.alpha = (ret < f.alpha) ? ret : f.alpha;      // min(alpha, child)
                __syncwarp();
            }
            if (stage > 4) { ret = f.alpha; --sp; __syncwarp(); continue; }
            const int child = stage - 1;
            const Cloud a = load_cloud_points(cur, n1a, lane);
            const Cloud b = load_cloud_points(cur + 3 * n1a, n1b, lane);
            Cloud l1, r1, l2, r2;
            split_cloud(a, 0.5, lane, l1, r1);
            split_cloud(b, 0.5, lane, l2, r2);
            double *nxt = cur + csz;
            store_cloud(nxt, (child < 2) ? l1 : r1, lane);
            store_cloud(nxt + 3 * n1a, (child & 1) ? r2 : l2, lane);
            if (lane == 0) {
                CCFrame &g = fr[sp + 1];
                g.alpha = f.alpha; g.stage = 0; g.cnt = f.cnt;
                f.stage = stage + 1;
            }
            ++sp;
            __syncwarp();
        }
    }
    if (lane == 0) out[item] = ret;
}

// _collCheckBez2Poly (bezier.py:1618-1651): 1 iff both halves clear the polytope at
// some depth; bounded by a node budget (the reference's colliding case is an
// unpruned binary recursion to depth 100).
struct CPFrame { int stage; int cnt; int first; };

__global__ void collcheck2poly_kernel(const double *c1, const double *polys, const int *npoly,
                                      int count, int dim1, int n1a, int mmax, long long max_nodes,
                                      double *scratch, double *out, int *status) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (item >= count) return;
    const int max_depth = 102;
    const int csz = 3 * n1a;
    const size_t per_item = (size_t)max_depth * (csz + sizeof(CPFrame) / sizeof(double) + 2);
    double *cs = scratch + (size_t)item * per_item;
    CPFrame *fr = reinterpret_cast<CPFrame *>(cs + (size_t)max_depth * csz);
    const Cloud poly = load_cloud_points(polys + (size_t)item * mmax * 3, npoly ? npoly[item] : mmax, lane);
    int sp = 0, stat = 0;
    long long nodes = 0;
    {
        const Cloud a = load_cloud_rows(c1 + (size_t)item * dim1 * n1a, dim1, n1a, lane);
        store_cloud(cs, a, lane);
        if (lane == 0) { fr[0].stage = 0; fr[0].cnt = 0; }
    }
    __syncwarp();
    int ret = 0;
    while (sp >= 0) {
        CPFrame &f = fr[sp];
        double *cur = cs + (size_t)sp * csz;
        const int stage = f.stage;
        if (stage == 0) {
            const int cnt = f.cnt + 1;
            if (cnt > 100) { ret = -1; --sp; __syncwarp(); continue; }
            if (++nodes > max_nodes) { stat = 1; ret = 0; sp = -1; break; }
            const Cloud a = load_cloud_points(cur, n1a, lane);
            const GjkResult g = gjk_new(a, poly, lane);
            if (g.flag > 0) { ret = 1; --sp; __syncwarp(); continue; }
            if (lane == 0) { f.cnt = cnt; f.stage = 1; }
            __syncwarp();
        } else {
            if (stage == 2 && ret != 1) { ret = 0; --sp; __syncwarp(); continue; }   // `and` short-circuit
            if (stage == 3) { ret = (ret == 1) ? 1 : 0; --sp; __syncwarp(); continue; }
            const int child = stage - 1;
            const Cloud a = load_cloud_points(cur, n1a, lane);
            Cloud l1, r1;
            split_cloud(a, 0.5, lane, l1, r1);
            store_cloud(cur + csz, child == 0 ? l1 : r1, lane);
            if (lane == 0) {
                CPFrame &g = fr[sp + 1];
                g.stage = 0; g.cnt = f.cnt;
                f.stage = stage + 1;
            }
            ++sp;
            __syncwarp();
        }
    }
    if (lane == 0) { out[item] = (double)ret; status[item] = stat; }
}

// Bezier.__call__ / .curve -> deCasteljauCurve (bezier.py:187-203, 240-262, 944-982):
// full de Casteljau triangle per sample, one thread per (row, tau); no FMA.
__global__ void eval_kernel(const double *c, const double *tau, long long rows, int n1, int ntau,
                            double t0, double tf, double *out) {
    const long long total = rows * ntau;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / ntau;
        const int k = (int)(idx - r * ntau);
        const double t = (tau[k] - t0) / (tf - t0);
        double w[BEZ_MAX_GEOM_PTS];
        for (int i = 0; i < n1; ++i) w[i] = c[r * n1 + i];
        for (int lvl = 1; lvl < n1; ++lvl)
            for (int i = 0; i < n1 - lvl; ++i) w[i] = (1 - t) * w[i] + t * w[i + 1];
        out[idx] = w[0];
    }
}

inline unsigned warp_blocks(int count) { return (unsigned)(((long long)count * 32 + kGeomThreads - 1) / kGeomThreads); }

}  // namespace

// ------------------------------- C ABI --------------------------------------
extern "C" int bez_gjk(const double *d_poly1, const double *d_poly2, const int *d_n1, const int *d_n2,
                       int n1max, int n2max, int count, int *d_flag, double *d_p1, double *d_p2,
                       double *d_dist, void *stream) {
    BEZ_REQUIRE(d_poly1 && d_poly2 && d_flag && d_p1 && d_p2 && d_dist, "NULL argument");
    BEZ_REQUIRE(n1max >= 1 && n1max <= BEZ_MAX_GEOM_PTS && n2max >= 1 && n2max <= BEZ_MAX_GEOM_PTS,
                "polygons must have 1..32 points");
    if (count <= 0) return BEZ_OK;
    gjk_kernel<<<warp_blocks(count), kGeomThreads, 0, (cudaStream_t)stream>>>(
        d_poly1, d_poly2, d_n1, d_n2, n1max, n2max, count, d_flag, d_p1, d_p2, d_dist);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_split(const double *d_cpts, const double *d_tlocal, int count, int dim, int n,
                         double *d_left, double *d_right, void *stream) {
    BEZ_REQUIRE(d_cpts && d_tlocal && d_left && d_right, "NULL argument");
    BEZ_REQUIRE(dim >= 1 && n >= 0 && n + 1 <= BEZ_MAX_GEOM_PTS, "degree must be <= 31");
    if (count <= 0) return BEZ_OK;
    split_kernel<<<warp_blocks(count), kGeomThreads, 0, (cudaStream_t)stream>>>(d_cpts, d_tlocal, count, dim,
                                                                                  n + 1, d_left, d_right);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_curve_eval(const double *d_cpts, const double *d_tau, int64_t rows, int n, int ntau,
                              double t0, double tf, double *d_out, void *stream) {
    BEZ_REQUIRE(d_cpts && d_tau && d_out, "NULL argument");
    BEZ_REQUIRE(rows >= 0 && ntau >= 0 && n >= 0 && n + 1 <= BEZ_MAX_GEOM_PTS, "degree must be <= 31");
    if (rows == 0 || ntau == 0) return BEZ_OK;
    long long blocks = (rows * ntau + 127) / 128;
    if (blocks > 148 * 32) blocks = 148 * 32;
    eval_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(d_cpts, d_tau, rows, n + 1, ntau, t0, tf, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" size_t bez_extrema_scratch_doubles(int count, int n, int max_depth) {
    return (size_t)count * max_depth * (n + 4);
}
extern "C" int bez_extrema(const double *d_rows, int count, int n, double tol, int maximum, int max_depth,
                           double *d_scratch, double *d_out, int *d_status, void *stream) {
    BEZ_REQUIRE(d_rows && d_scratch && d_out && d_status, "NULL argument");
    BEZ_REQUIRE(n >= 0 && n + 1 <= BEZ_MAX_GEOM_PTS && max_depth >= 2, "degree must be <= 31, depth >= 2");
    if (count <= 0) return BEZ_OK;
    extrema_kernel<<<warp_blocks(count), kGeomThreads, 0, (cudaStream_t)stream>>>(
        d_rows, count, n + 1, tol, maximum, max_depth, d_scratch, d_out, d_status);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" size_t bez_mindist_scratch_doubles(int count, int n1, int n2, int max_depth) {
    return (size_t)count * max_depth * (3 * (n1 + 1 + n2 + 1) + sizeof(MDFrame) / sizeof(double) + 1);
}
extern "C" int bez_mindist(const double *d_c1, const double *d_c2, int count, int dim1, int dim2, int n1,
                           int n2, double eps, int max_depth, long long max_nodes, double *d_scratch,
                           double *d_out, int *d_status, void *stream) {
    BEZ_REQUIRE(d_c1 && d_c2 && d_scratch && d_out && d_status, "NULL argument");
    BEZ_REQUIRE(dim1 >= 2 && dim1 <= 3 && dim2 >= 2 && dim2 <= 3, "curves must be 2-D or 3-D");
    BEZ_REQUIRE(n1 >= 1 && n2 >= 1 && n1 + 1 <= BEZ_MAX_GEOM_PTS && n2 + 1 <= BEZ_MAX_GEOM_PTS && max_depth >= 2,
                "degrees must be 1..31");
    if (count <= 0) return BEZ_OK;
    mindist_kernel<<<warp_blocks(count), kGeomThreads, 0, (cudaStream_t)stream>>>(
        d_c1, d_c2, count, dim1, dim2, n1 + 1, n2 + 1, eps, max_depth, max_nodes, d_scratch, d_out, d_status);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" size_t bez_mindist2poly_scratch_doubles(int count, int n1, int max_depth) {
    return (size_t)count * max_depth * (3 * (n1 + 1) + sizeof(MPFrame) / sizeof(double) + 1);
}
extern "C" int bez_mindist2poly(const double *d_c1, const double *d_polys, const int *d_npoly, int count,
                                int dim1, int n1, int mmax, double eps, int max_depth, long long max_nodes,
                                double *d_scratch, double *d_out, int *d_status, void *stream) {
    BEZ_REQUIRE(d_c1 && d_polys && d_scratch && d_out && d_status, "NULL argument");
    BEZ_REQUIRE(dim1 >= 2 && dim1 <= 3, "curves must be 2-D or 3-D");
    BEZ_REQUIRE(n1 >= 1 && n1 + 1 <= BEZ_MAX_GEOM_PTS && mmax >= 1 && mmax <= BEZ_MAX_GEOM_PTS && max_depth >= 2,
                "degree / polytope size out of range");
    if (count <= 0) return BEZ_OK;
    mindist2poly_kernel<<<warp_blocks(count), kGeomThreads, 0, (cudaStream_t)stream>>>(
        d_c1, d_polys, d_npoly, count, dim1, n1 + 1, mmax, eps, max_depth, max_nodes, d_scratch, d_out, d_status);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" size_t bez_collcheck_scratch_doubles(int count, int n1, int n2) {
    return (size_t)count * 102 * (3 * (n1 + 1 + n2 + 1) + sizeof(CCFrame) / sizeof(double) + 1);
}
extern "C" int bez_collcheck(const double *d_c1, const double *d_c2, int count, int dim1, int dim2, int n1,
                             int n2, double eps, double *d_scratch, double *d_out, void *stream) {
    BEZ_REQUIRE(d_c1 && d_c2 && d_scratch && d_out, "NULL argument");
    BEZ_REQUIRE(dim1 >= 2 && dim1 <= 3 && dim2 >= 2 && dim2 <= 3, "curves must be 2-D or 3-D");
    BEZ_REQUIRE(n1 >= 1 && n2 >= 1 && n1 + 1 <= BEZ_MAX_GEOM_PTS && n2 + 1 <= BEZ_MAX_GEOM_PTS, "degrees must be 1..31");
    if (count <= 0) return BEZ_OK;
    collcheck_kernel<<<warp_blocks(count), kGeomThreads, 0, (cudaStream_t)stream>>>(
        d_c1, d_c2, count, dim1, dim2, n1 + 1, n2 + 1, eps, d_scratch, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" size_t bez_collcheck2poly_scratch_doubles(int count, int n1) {
    return (size_t)count * 102 * (3 * (n1 + 1) + sizeof(CPFrame) / sizeof(double) + 2);
}
extern "C" int bez_collcheck2poly(const double *d_c1, const double *d_polys, const int *d_npoly, int count,
                                  int dim1, int n1, int mmax, long long max_nodes, double *d_scratch,
                                  double *d_out, int *d_status, void *stream) {
    BEZ_REQUIRE(d_c1 && d_polys && d_scratch && d_out && d_status, "NULL argument");
    BEZ_REQUIRE(dim1 >= 2 && dim1 <= 3, "curves must be 2-D or 3-D");
    BEZ_REQUIRE(n1 >= 1 && n1 + 1 <= BEZ_MAX_GEOM_PTS && mmax >= 1 && mmax <= BEZ_MAX_GEOM_PTS,
                "degree / polytope size out of range");
    if (count <= 0) return BEZ_OK;
    collcheck2poly_kernel<<<warp_blocks(count), kGeomThreads, 0, (cudaStream_t)stream>>>(
        d_c1, d_polys, d_npoly, count, dim1, n1 + 1, mmax, max_nodes, d_scratch, d_out, d_status);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}
