// Shared device code of the fused "square -> fold -> elevate" kernels
// (constraints.cu: values; jacobian.cu: finite-difference Jacobian columns).
#pragma once
#include "common.cuh"

namespace bezcore {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

template <int N_>
struct ProdWeights {            // unique weights W[i][j], i <= j; off-diagonal doubled
    double w[(N_ + 1) * (N_ + 2) / 2];
};
template <int N_>
struct DiffWeights {            // elevMatrix(n-1,1) diagonals used by Bezier.diff
    double lo[N_ + 1], hi[N_ + 1];
};

template <int N_>
__host__ __device__ constexpr int widx(int i, int j) {   // i <= j
    return i * (N_ + 1) - i * (i - 1) / 2 + (j - i);
}

// min without fmin()'s NaN bookkeeping (1 DSETP + 2 SEL instead of ~6 instructions)
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
// NaN-propagating variant for reductions that mix real values with +inf placeholders (idle
// lanes, dead columns): a NaN row must yield a NaN minimum, like numpy's min (a row is NaN
// in all its values or in none, because every output sums all 2n+1 coefficients).
__device__ __forceinline__ double dmin_nan(double a, double b) { return (a < b || a != a) ? a : b; }

// Min over the L outputs of each item: separate pass for the shapes the fused
// epilogue does not cover (L > 128 or degree < 4).
static __global__ void item_min_kernel(const double *__restrict__ vals, long long nrows, int L,
                                double *__restrict__ mins) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const double *r = vals + (size_t)warp * L;
    double m = INFINITY;
    for (int i = lane; i < L; i += 32) m = dmin_nan(m, r[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = dmin_nan(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) mins[warp] = m;
}

// Shared-memory geometry of one staged item row: SLOTS double2 slots
// (slot j < n : (e_j, o_j); slot n : (e_n, 0); the rest is padding so that the
// prefetch ring of depth kRing indexes statically), row stride RS slots with RS
// odd so that the per-lane 128-bit stores of stage 1 are bank-conflict free.
constexpr int kRing = 4;
template <int N_> struct RowGeom {
    static constexpr int SLOTS = (N_ + 1 + kRing - 1) / kRing * kRing;
    static constexpr int RS = SLOTS | 1;
};

// Stage 2 of sq_elev_kernel for one group of 32*CPL column pairs: lane l owns
// columns col_c = (g*CPL + c)*32 + l and their mirrors M - col_c.  Items are
// consumed two at a time (8 independent DFMA chains per lane for CPL = 2); their
// staged rows stream through a kRing-deep register ring refilled kRing slots
// ahead (LDS latency ~ 14 DFMA issue slots).  FULL = all 32 items of the tile
// are live: the loop body is then branch free, so the scheduler can overlap the
// stores of one unit with the DFMAs of the next.
template <int N_, int CPL, bool WITH_MIN, bool FULL, bool SCATTER = false>
__device__ __forceinline__ void sweep_columns(double2 *rows, const double *tab, double *outb,
                                              int g, int lane, int cnt, int L, int Lh, int LhPad,
                                              double beta, const long long *rowoff = nullptr) {
    constexpr int NC = N_ + 1;
    constexpr int SLOTS = RowGeom<N_>::SLOTS;
    constexpr int RS = RowGeom<N_>::RS;
    const int M = L - 1;
    bool live[CPL];
    double P[CPL][NC], Q[CPL][N_ > 0 ? N_ : 1];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        const int col = (g * CPL + c) * 32 + lane;
        live[c] = col < Lh;
        const int tc = col < LhPad ? col : 0;
#pragma unroll
        for (int j = 0; j < NC; ++j) P[c][j] = tab[j * LhPad + tc];
#pragma unroll
        for (int j = 0; j < N_; ++j) Q[c][j] = tab[(NC + j) * LhPad + tc];
    }
    // per-lane output cursors: column c sits at fwd + 32*c, its mirror at mir - 32*c
    // (compile-time displacements); both advance by L per item.  The centre column
    // of an even-degree result is its own mirror: both stores then write the same
    // value (its odd part is exactly 0), so no special case is needed.
    double *fwd = outb + (g * CPL * 32 + lane);
    double *mir = outb + (M - g * CPL * 32 - lane);

    double2 ra[kRing], rb[kRing];
#pragma unroll
    for (int r = 0; r < kRing; ++r) {
        ra[r] = rows[r];                 // row 0
        rb[r] = rows[RS + r];            // row 1
    }
    for (int p0 = 0; p0 < cnt; p0 += 2) {
        // rows of the next unit for the wrap-around prefetch (clamped in-bounds)
        const int pn = (p0 + 2 < 32) ? p0 + 2 : p0;
        const double2 *rowA = rows + (size_t)p0 * RS;
        const double2 *rowB = rowA + RS;
        const double2 *nxtA = rows + (size_t)pn * RS;
        const double2 *nxtB = nxtA + RS;
        double se[2][CPL], so[2][CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            se[0][c] = beta; se[1][c] = beta;
            so[0][c] = 0.0; so[1][c] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            const int r = j % kRing;
            if (j < N_) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    se[0][c] = fma(ra[r].x, P[c][j], se[0][c]);
                    so[0][c] = fma(ra[r].y, Q[c][j], so[0][c]);
                    se[1][c] = fma(rb[r].x, P[c][j], se[1][c]);
                    so[1][c] = fma(rb[r].y, Q[c][j], so[1][c]);
                }
            } else if (j == N_) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    se[0][c] = fma(ra[r].x, P[c][N_], se[0][c]);
                    se[1][c] = fma(rb[r].x, P[c][N_], se[1][c]);
                }
            }
            const int jn = j + kRing;
            if (jn < SLOTS) {
                if (jn <= N_) { ra[r] = rowA[jn]; rb[r] = rowB[jn]; }
            } else if (jn - SLOTS <= N_) {
                ra[r] = nxtA[jn - SLOTS]; rb[r] = nxtB[jn - SLOTS];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool item_live = FULL || (p0 + u < cnt);
            if (SCATTER) {          // items land at arbitrary rows (Jacobian layouts)
                const long long ro = rowoff[(p0 + u < cnt) ? p0 + u : cnt - 1];
                fwd = outb + ro + (g * CPL * 32 + lane);
                mir = outb + ro + (M - g * CPL * 32 - lane);
            }
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (live[c] && item_live) {
                    __stcs(fwd + 32 * c, se[u][c] + so[u][c]);
                    __stcs(mir - 32 * c, se[u][c] - so[u][c]);
                }
            }
            fwd += L;
            mir += L;
            if (WITH_MIN) {
                // min(se+so, se-so) == se - |so| bit for bit (one DADD)
                double mn = live[0] ? se[u][0] - fabs(so[u][0]) : INFINITY;
#pragma unroll
                for (int c = 1; c < CPL; ++c)
                    mn = live[c] ? dmin_nan(mn, se[u][c] - fabs(so[u][c])) : mn;
                // fold the warp in half, then park the 16 partial minima in the
                // (already consumed) staged row of this item
                mn = dmin_nan(mn, __shfl_xor_sync(0xffffffffu, mn, 16));
                if (lane < 16) reinterpret_cast<double *>(rows + (size_t)(p0 + u) * RS)[lane] = mn;
            }
        }
    }
}


}  // namespace bezcore
