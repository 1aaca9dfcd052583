// Single-curve Bezier algebra behind the drop-in bezier.Bezier methods
// (A1-A3 of SURVEY.md section 8 outside the fused constraint kernels): elev, diff,
// mul, normSquare for batches of independent curves.  These are the same formulas
// as the fused kernels, exposed per operation so user code that calls the methods
// one by one (Examples/*.py) gets identical semantics.  One thread per output
// control point; tables are device arrays built on the host with scipy binom.
#include "common.cuh"

namespace {

// out[r][i] = sum_j c[r][j] * T[j][i]                     (Bezier.elev, bezier.py:469-495)
__global__ void elev_kernel(const double *c, const double *T, long long rows, int n1, int L, double *out) {
    const long long total = rows * L;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / L;
        const int i = (int)(idx - r * L);
        double s = 0.0;
        for (int j = 0; j < n1; ++j) s = fma(c[r * n1 + j], T[(size_t)j * L + i], s);
        out[idx] = s;
    }
}

// Bezier.diff (bezier.py:497-519): np.dot(cpts, Dm) then elev(1); E1 = elevMatrix(n-1,1) [n][n+1]
__global__ void diff_kernel(const double *c, const double *E1, const double *T, long long rows,
                            long long rows_per_T, int n1, double *out) {
    const int n = n1 - 1;
    const long long total = rows * n1;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / n1;
        const int k = (int)(idx - r * n1);
        const double val = (double)n / T[r / rows_per_T];
        const double *p = c + r * n1;
        double q = 0.0;
        if (k < n) q = (p[k] * (-val) + p[k + 1] * val) * E1[(size_t)k * n1 + k];
        if (k > 0) q = (p[k - 1] * (-val) + p[k] * val) * E1[(size_t)(k - 1) * n1 + k] + q;
        out[idx] = q;
    }
}

// Bezier.mul / multiplyBezCurves (bezier.py:376-432, 1211-1246), row-wise:
// out[r][k] = sum_{i+j=k} W[i][j] a[r][i] b[r][j],   W [m1][n1]
__global__ void mul_kernel(const double *a, const double *b, const double *W, long long rows, int m1,
                           int n1, double *out) {
    const int L = m1 + n1 - 1;
    const long long total = rows * L;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / L;
        const int k = (int)(idx - r * L);
        const int ilo = k - (n1 - 1) > 0 ? k - (n1 - 1) : 0, ihi = k < m1 - 1 ? k : m1 - 1;
        double s = 0.0;
        for (int i = ilo; i <= ihi; ++i)
            s = fma(a[r * m1 + i] * b[r * n1 + (k - i)], W[(size_t)i * n1 + (k - i)], s);
        out[idx] = s;
    }
}

// Bezier.normSquare (bezier.py:869-889, 1724-1756): (dim/2) * sum_d c_d^2 per curve
__global__ void normsq_kernel(const double *c, const double *W, long long curves, int dim, int n1,
                              double *out) {
    const int L = 2 * n1 - 1;
    const long long total = curves * L;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / L;
        const int k = (int)(idx - r * L);
        const int ilo = k - (n1 - 1) > 0 ? k - (n1 - 1) : 0, ihi = k < n1 - 1 ? k : n1 - 1;
        const double *p = c + r * dim * n1;
        double s = 0.0;
        for (int i = ilo; i <= ihi; ++i) {
            double g = 0.0;
            for (int d = 0; d < dim; ++d) g = fma(p[d * n1 + i], p[d * n1 + (k - i)], g);
            s = fma(W[(size_t)i * n1 + (k - i)], g, s);
        }
        out[idx] = (s * (double)dim) / 2.0;
    }
}

// ---- cost callables (A14) -----------------------------------------------------
__device__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;       // valid in thread 0
}

// _euclideanObjective (optimization.py:462-489): total control-polygon length.
// cpts [B][N][S]; one block per evaluation point.
__global__ void euclid_kernel(const double *cpts, int N, int S, int numVeh, int dim, int n1, double *out) {
    __shared__ double red[32];
    const double *base = cpts + (size_t)blockIdx.x * N * S;
    double acc = 0.0;
    const int segs = numVeh * (n1 - 1);
    for (int s = threadIdx.x; s < segs; s += blockDim.x) {
        const int v = s / (n1 - 1), i = s - v * (n1 - 1);
        double q = 0.0;
        for (int d = 0; d < dim; ++d) {
            const double t = base[(size_t)v * S + d * n1 + i + 1] - base[(size_t)v * S + d * n1 + i];
            q = fma(t, t, q);
        }
        acc += sqrt(q);
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) out[blockIdx.x] = tot;
}

// _minAccelObjective (optimization.py:503-519): sum over vehicles of the control
// points of elev(normSquare(diff(diff(pos))), E).  T = elevMatrix(2n, E) [2n+1][L].
__global__ void accel_kernel(const double *cpts, const double *tf, const double *W, const double *E1,
                             const double *T, int N, int S, int numVeh, int dim, int n1, int L, double *out) {
    __shared__ double red[32];
    const int n = n1 - 1;
    const double *base = cpts + (size_t)blockIdx.x * N * S;
    const double val = (double)n / tf[blockIdx.x];
    double acc = 0.0;
    for (int v = threadIdx.x; v < numVeh; v += blockDim.x) {
        double a[3][32];
        for (int d = 0; d < dim; ++d) {
            double p[32], q[32];
            for (int k = 0; k < n1; ++k) p[k] = base[(size_t)v * S + d * n1 + k];
            for (int rep = 0; rep < 2; ++rep) {                 // diff twice
                for (int k = 0; k < n1; ++k) {
                    double r = 0.0;
                    if (k < n) r = (p[k] * (-val) + p[k + 1] * val) * E1[(size_t)k * n1 + k];
                    if (k > 0) r = (p[k - 1] * (-val) + p[k] * val) * E1[(size_t)(k - 1) * n1 + k] + r;
                    q[k] = r;
                }
                for (int k = 0; k < n1; ++k) p[k] = q[k];
            }
            for (int k = 0; k < n1; ++k) a[d][k] = p[k];
        }
        for (int k = 0; k <= 2 * n; ++k) {
            const int ilo = k - n > 0 ? k - n : 0, ihi = k < n ? k : n;
            double s = 0.0;
            for (int i = ilo; i <= ihi; ++i) {
                double g = 0.0;
                for (int d = 0; d < dim; ++d) g = fma(a[d][i], a[d][k - i], g);
                s = fma(W[(size_t)i * n1 + (k - i)], g, s);
            }
            s = (s * (double)dim) / 2.0;
            double rs = 0.0;
            for (int i = 0; i < L; ++i) rs += T[(size_t)k * L + i];
            acc = fma(s, rs, acc);
        }
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) out[blockIdx.x] = tot;
}

// Gradient of the cost callables as SciPy's '2-point' rule defines it (the objective analogue
// of jacobian.cu): g_k = (f(x + dx_k e_k) - f(x)) / dx_k, evaluated without cancellation.
// Variable k = (vehicle v, dimension d, free column col) moves control point c = col + offset of
// row d by dx_k and nothing else (fixed-tf models: the cost callables that depend on control
// points are never used with a tf variable).  One thread per variable; cpts [N][S] of the base x.
//   euclid: only the two polygon segments at c change; |a + dx e_d| - |a| = dx (2 a_d + dx) / (|a + dx e_d| + |a|)
//   accel : f is a quadratic form of the control points, acc(P + dx e) = acc(P) + dx acc(e):
//           quotient = sum_k rs_k (dim/2) sum_{i+j=k} W_ij (2 a_i u_j + dx u_i u_j),  u = acc(e_c)
__global__ void euclid_grad_kernel(const double *cpts, int S, int dim, int n1, int ncols, int offset,
                                   const double *dx, int nvar, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nvar) return;
    const int col = k % ncols, vd = k / ncols, d = vd % dim, v = vd / dim;
    const int c = col + offset;
    const double *p = cpts + (size_t)v * S;
    const double h = dx[k];
    double acc = 0.0;
    for (int side = 0; side < 2; ++side) {
        // side 0: segment (c-1 -> c), a = P_c - P_{c-1}, moves by +h e_d
        // side 1: segment (c -> c+1), a = P_{c+1} - P_c, moves by -h e_d
        const int lo = side == 0 ? c - 1 : c;
        if (lo < 0 || lo + 1 >= n1) continue;
        const double sh = side == 0 ? h : -h;
        double q0 = 0.0, q1 = 0.0, ad = 0.0;
        for (int j = 0; j < dim; ++j) {
            const double t = p[j * n1 + lo + 1] - p[j * n1 + lo];
            const double t1 = (j == d) ? t + sh : t;
            if (j == d) ad = t;
            q0 = fma(t, t, q0);
            q1 = fma(t1, t1, q1);
        }
        const double den = sqrt(q1) + sqrt(q0);
        acc += den > 0.0 ? sh * (2.0 * ad + sh) / den : fabs(sh);
    }
    out[k] = acc / h;
}

__global__ void accel_grad_kernel(const double *cpts, double tf, const double *W, const double *E1,
                                  const double *T, int S, int dim, int n1, int L, int ncols, int offset,
                                  const double *dx, int nvar, double *out) {
    extern __shared__ double rs[];                      // row sums of elevMatrix(2n, E)
    const int n = n1 - 1;
    for (int r = threadIdx.x; r <= 2 * n; r += blockDim.x) {
        double t = 0.0;
        for (int i = 0; i < L; ++i) t += T[(size_t)r * L + i];
        rs[r] = t;
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nvar) return;
    const int col = k % ncols, vd = k / ncols, d = vd % dim, v = vd / dim;
    const int c = col + offset;
    const double val = (double)n / tf;
    double a[32], u[32], q[32];
    for (int i = 0; i < n1; ++i) { a[i] = cpts[(size_t)v * S + d * n1 + i]; u[i] = (i == c) ? 1.0 : 0.0; }
    for (int which = 0; which < 2; ++which) {
        double *p = which == 0 ? a : u;
        for (int rep = 0; rep < 2; ++rep) {              // Bezier.diff twice (Q3: degree kept)
            for (int i = 0; i < n1; ++i) {
                double r = 0.0;
                if (i < n) r = (p[i] * (-val) + p[i + 1] * val) * E1[(size_t)i * n1 + i];
                if (i > 0) r = (p[i - 1] * (-val) + p[i] * val) * E1[(size_t)(i - 1) * n1 + i] + r;
                q[i] = r;
            }
            for (int i = 0; i < n1; ++i) p[i] = q[i];
        }
    }
    const double h = dx[k];
    double acc = 0.0;
    for (int kk = 0; kk <= 2 * n; ++kk) {
        const int ilo = kk - n > 0 ? kk - n : 0, ihi = kk < n ? kk : n;
        double s = 0.0;
        for (int i = ilo; i <= ihi; ++i)
            s = fma(W[(size_t)i * n1 + (kk - i)], fma(h * u[i], u[kk - i], 2.0 * a[i] * u[kk - i]), s);
        acc = fma((s * (double)dim) / 2.0, rs[kk], acc);
    }
    out[k] = acc;
}

inline unsigned blocks_for(long long total) {
    long long b = (total + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int bez_curve_elev(const double *d_cpts, const double *d_T, int64_t rows, int n, int R,
                              double *d_out, void *stream) {
    BEZ_REQUIRE(d_cpts && d_T && d_out, "NULL argument");
    BEZ_REQUIRE(rows >= 0 && n >= 0 && R >= 0, "negative size");
    if (rows == 0) return BEZ_OK;
    elev_kernel<<<blocks_for(rows * (n + R + 1)), 256, 0, (cudaStream_t)stream>>>(d_cpts, d_T, rows, n + 1,
                                                                                   n + R + 1, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_curve_diff(const double *d_cpts, const double *d_E1, const double *d_T, int64_t rows,
                              int64_t rows_per_T, int n, double *d_out, void *stream) {
    BEZ_REQUIRE(d_cpts && d_E1 && d_T && d_out, "NULL argument");
    BEZ_REQUIRE(rows >= 0 && n >= 1 && rows_per_T >= 1, "bad size");
    if (rows == 0) return BEZ_OK;
    diff_kernel<<<blocks_for(rows * (n + 1)), 256, 0, (cudaStream_t)stream>>>(d_cpts, d_E1, d_T, rows,
                                                                              rows_per_T, n + 1, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_curve_mul(const double *d_a, const double *d_b, const double *d_W, int64_t rows, int m,
                             int n, double *d_out, void *stream) {
    BEZ_REQUIRE(d_a && d_b && d_W && d_out, "NULL argument");
    BEZ_REQUIRE(rows >= 0 && m >= 0 && n >= 0, "negative size");
    if (rows == 0) return BEZ_OK;
    mul_kernel<<<blocks_for(rows * (m + n + 1)), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, d_W, rows, m + 1,
                                                                                 n + 1, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_curve_normsq(const double *d_cpts, const double *d_W, int64_t curves, int dim, int n,
                                double *d_out, void *stream) {
    BEZ_REQUIRE(d_cpts && d_W && d_out, "NULL argument");
    BEZ_REQUIRE(curves >= 0 && dim >= 1 && n >= 0, "bad size");
    if (curves == 0) return BEZ_OK;
    normsq_kernel<<<blocks_for(curves * (2 * n + 1)), 256, 0, (cudaStream_t)stream>>>(d_cpts, d_W, curves, dim,
                                                                                      n + 1, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_objective_euclidean(const bez_plan *plan, const double *d_cpts, int B, int N, int numVeh,
                                       double *d_out, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_out, "NULL argument");
    BEZ_REQUIRE(B >= 0 && numVeh >= 1 && numVeh <= N, "bad sizes");
    if (B == 0) return BEZ_OK;
    const int n1 = plan->n + 1, S = (plan->dim * n1 + 1) / 2 * 2;
    euclid_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(d_cpts, N, S, numVeh, plan->dim, n1, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_objective_accel(const bez_plan *plan, const double *d_cpts, const double *d_tf, int B,
                                   int N, int numVeh, double *d_out, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_tf && d_out, "NULL argument");
    BEZ_REQUIRE(B >= 0 && numVeh >= 1 && numVeh <= N, "bad sizes");
    if (B == 0) return BEZ_OK;
    const int n1 = plan->n + 1, S = (plan->dim * n1 + 1) / 2 * 2;
    accel_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(d_cpts, d_tf, plan->d_W, plan->d_E1, plan->d_T, N, S,
                                                     numVeh, plan->dim, n1, plan->L, d_out);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_objective_grad(const bez_plan *plan, const double *d_cpts, int kind, double tf, int numVeh,
                                  int ncols, int offset, const double *d_dx, double *d_out, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_dx && d_out, "NULL argument");
    BEZ_REQUIRE(numVeh >= 1 && ncols >= 0 && offset >= 0 && ncols + 2 * offset == plan->n + 1,
                "free columns + fixed end columns must add up to degree + 1");
    BEZ_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (euclidean) or 1 (accel)");
    const int n1 = plan->n + 1, S = (plan->dim * n1 + 1) / 2 * 2;
    const int nvar = numVeh * plan->dim * ncols;
    if (nvar == 0) return BEZ_OK;
    BEZ_ON_DEVICE(plan->device);
    const unsigned blocks = (unsigned)((nvar + 127) / 128);
    if (kind == 0) {
        euclid_grad_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(d_cpts, S, plan->dim, n1, ncols, offset, d_dx,
                                                                      nvar, d_out);
    } else {
        BEZ_REQUIRE(tf != 0.0, "tf must be non-zero");
        accel_grad_kernel<<<blocks, 128, sizeof(double) * (2 * plan->n + 1), (cudaStream_t)stream>>>(
            d_cpts, tf, plan->d_W, plan->d_E1, plan->d_T, S, plan->dim, n1, plan->L, ncols, offset, d_dx, nvar, d_out);
    }
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}
