// Shared by constraints.cu (DFMA stage 2) and constraints_mma.cu (DMMA stage 2):
// kernel arguments and stage 1 of the fused "square -> fold -> elevate" kernels.
#pragma once
#include "sq_elev_core.cuh"

namespace bezcore {

enum Mode { PAIR = 0, SPEED = 1 };

struct SqElevArgs {
    const double *cpts;     // [B][N][S]  S = dim*(n+1) rounded up to even
    const double *tf;       // [B] (SPEED)
    const double *PQ;       // [2n+1][LhPad]
    double *out;            // [B][nitems][L]
    double *itemmin;        // [B][nitems] or null
    long long item_begin;   // first pair / vehicle handled
    long long nitems;       // pairs / vehicles per evaluation point
    int B, N, L, Lh, LhPad;
    double alpha, beta;     // out = alpha * value + beta   (alpha = +-1)
    // Fused all-gather of the per-item minima over NVLink peer memory (tensor path only):
    // every minimum is also stored to npeers other GPUs' gathered matrices (pointers already
    // offset to this rank's block), so no collective follows the kernel.
    double *peer_min[BEZ_MAX_PEERS];
    int npeers;
};

// Stage 1 for one item (lane = item): the 2n+1 Bernstein coefficients (before the
// dim/2 scale) of |a|^2, a = c_i - c_j (PAIR) or the derivative curve (SPEED).
// li = item index inside the tile; lanes past the end recompute the last item so
// every staged row is finite.
template <int N_, int DIM, int MODE>
__device__ __forceinline__ void stage1_coeffs(const SqElevArgs &A, const ProdWeights<N_> &PW,
                                              const DiffWeights<N_> &DW, int b, long long t0, int li,
                                              double (&s)[2 * N_ + 1]) {
    constexpr int NC = N_ + 1;
    constexpr int S = (DIM * NC + 1) / 2 * 2;                 // doubles per vehicle row (16 B aligned)
    const double *base = A.cpts + (size_t)b * ((size_t)S * A.N);
    double a[DIM][NC];
    if (MODE == PAIR) {
        int vi, vj;
        bez_pair_decode(A.item_begin + t0 + li, A.N, vi, vj);
        const double2 *pi = reinterpret_cast<const double2 *>(base + (size_t)vi * S);
        const double2 *pj = reinterpret_cast<const double2 *>(base + (size_t)vj * S);
        double *af = &a[0][0];
#pragma unroll
        for (int q = 0; q < S / 2; ++q) {                      // Bezier.sub
            const double2 u = __ldg(pi + q), w = __ldg(pj + q);
            if (2 * q < DIM * NC) af[2 * q] = u.x - w.x;
            if (2 * q + 1 < DIM * NC) af[2 * q + 1] = u.y - w.y;
        }
    } else {
        const int v = (int)(A.item_begin + t0 + li);
        const double val = (double)N_ / __ldg(A.tf + b);       // diffMatrix: n/tf
        const double2 *pv = reinterpret_cast<const double2 *>(base + (size_t)v * S);
        double ptf[S];
#pragma unroll
        for (int q = 0; q < S / 2; ++q) {
            const double2 u = __ldg(pv + q);
            ptf[2 * q] = u.x;
            ptf[2 * q + 1] = u.y;
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double dd[NC];
#pragma unroll
            for (int k = 0; k < N_; ++k)                       // np.dot(cpts, Dm)
                dd[k] = ptf[d * NC + k] * (-val) + ptf[d * NC + k + 1] * val;
            dd[N_] = 0.0;
#pragma unroll
            for (int k = 0; k < NC; ++k) {                     // .elev(1) back to degree n
                double q = dd[k] * DW.lo[k];
                if (k > 0) q = dd[k - 1] * DW.hi[k] + q;
                a[d][k] = q;
            }
        }
    }
#pragma unroll
    for (int k = 0; k <= 2 * N_; ++k) s[k] = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
        for (int j = i; j < NC; ++j) {
            double g = a[0][i] * a[0][j];
#pragma unroll
            for (int d = 1; d < DIM; ++d) g = fma(a[d][i], a[d][j], g);
            s[i + j] = fma(PW.w[widx<N_>(i, j)], g, s[i + j]);
        }
}

// constraints_mma.cu: fp64 tensor path for 33..64 column pairs (65 <= L <= 128), degree <= 15
bool bez_sq_elev_mma_supported(const bez_plan *plan);
int bez_sq_elev_mma(const bez_plan *plan, const SqElevArgs &A, int mode, cudaStream_t st);

}  // namespace bezcore
