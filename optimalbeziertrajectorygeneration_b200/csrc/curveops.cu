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
