// Shared by constraints.cu (DFMA stage 2) and constraints_mma.cu (DMMA stage 2):
// kernel arguments and stage 1 of the fused "square -> fold -> elevate" kernels.
#pragma once
#include "sq_elev_mma.cuh"

namespace bezcore {

enum Mode { PAIR = 0, SPEED = 1 };

struct SqElevArgs {
    const double *cpts;     // [B][N][S]  S = dim*(n+1) rounded up to even
    const double *tf;       // [B] (SPEED)
    const double *PQ;       // [2n+1][LhPad]
    double *out;            // [B][nitems][L]; may be null when the rows are not wanted (tensor path)
    long long item_begin;   // first pair / vehicle handled
    long long nitems;       // pairs / vehicles per evaluation point
    int B, N, L, Lh, LhPad;
    double alpha, beta;     // out = alpha * value + beta   (alpha = +-1)
    // Per-item minima and what is derived from them (bezmma::MinSinks): the local [B][pitch]
    // matrix, the fused all-gather over NVLink peer memory (every minimum is also stored to
    // npeers other GPUs' gathered matrices, pointers already offset to this rank's block, so no
    // collective follows the kernel), the packed active bitmask and the compacted list.
    bezmma::MinSinks sinks;
    int flags;              // kFlag* (experiments / ablations), 0 in production
};
constexpr int kFlagPrefetchL1 = 1;     // prefetch.global.L1 of the next tile's vehicle rows
constexpr int kFlagStridedTiles = 2;   // round-1 tile order (warp-strided, full decode per tile)
// (bits 4 and 8 were the round-2 ablations "no TMA store" / "no fence": profiles/r02_ablation_pair_kernel.txt;
//  reused since:)
constexpr int kFlagRowsByLdg = 4;      // A/B: partner rows through per-lane global loads (no TMA row fetch)
constexpr int kFlagEarlyFence = 8;     // A/B: proxy fence + bulk store right behind an m-tile's epilogue
constexpr int kFlagFullGrid = 16;       // A/B: do not leave a CTA slot / an SM free for the small kernels around the pair kernel
constexpr int kFlagFullGridWithPeers = kFlagFullGrid;
constexpr int kFlagFusedList = 64;     // A/B: append the active list from the epilogue even in large launches
constexpr int kFlagNoWarpSpecialisation = 128;  // A/B: 65 <= L <= 128 pair rows on the 8-warp kernel instead of sq_elev_ws.cuh
// (bit 32 was the third-generation "team" kernel: profiles/r02_ablation_pair_kernel.txt)


// Position of a lane's item in the lexicographic pair list, advanced incrementally from tile
// to tile (a warp owns a contiguous run of tiles): no 64-bit division, square root and fix-up
// loop per tile -- that chain was ~400 cycles of pure latency in front of every stage 1.
struct PairCursor {
    int b, i, j;            // evaluation point, pair (i, j)
    long long left;         // items of evaluation b from this one on (this one included)
};
__device__ __forceinline__ PairCursor pair_cursor_at(const SqElevArgs &A, long long f) {
    PairCursor c;
    c.b = (int)(f / A.nitems);
    const long long it = f - (long long)c.b * A.nitems;
    c.left = A.nitems - it;
    bez_pair_decode(A.item_begin + it, A.N, c.i, c.j);
    return c;
}
// cursor of flattened item f + 32 given the cursor of f (f + 32 < total)
__device__ __forceinline__ PairCursor pair_cursor_next(const SqElevArgs &A, const PairCursor &c, long long f) {
    if (c.left <= 32) return pair_cursor_at(A, f + 32);      // crosses into the next evaluation point
    PairCursor d = c;
    d.left -= 32;
    int q = c.j + 32;                                          // still row i if q <= N - 1
    // past the end of row i (j = i+1 .. N-1) by e = q - N: row i+1 starts at j = i+2, so j = i+2+e
    while (q > A.N - 1) { q -= (A.N - 2 - d.i); ++d.i; }
    d.j = q;
    return d;
}

// Stage 1 for one item (lane = item): the 2n+1 Bernstein coefficients (before the
// dim/2 scale) of |a|^2, a = c_i - c_j (PAIR) or the derivative curve (SPEED).
// li = item index inside the tile; lanes past the end recompute the last item so
// every staged row is finite.
// PAIR: (vi, vj) = the pair; SPEED: vi = the vehicle.
// JROW_SMEM (PAIR only): row vj has already been brought into shared memory (jrow, 16-byte
// aligned; the TMA row fetch of sq_elev_mma_kernel.cuh) and is read with LDS.128 instead of
// 17 global loads whose 32 lanes touch 32 different cache lines each.
template <int N_, int DIM, int MODE, bool JROW_SMEM = false>
__device__ __forceinline__ void stage1_coeffs(const SqElevArgs &A, const ProdWeights<N_> &PW,
                                              const DiffWeights<N_> &DW, int b, int vi, int vj,
                                              double (&s)[2 * N_ + 1], const double *jrow = nullptr) {
    constexpr int NC = N_ + 1;
    constexpr int S = (DIM * NC + 1) / 2 * 2;                 // doubles per vehicle row (16 B aligned)
    const double *base = A.cpts + (size_t)b * ((size_t)S * A.N);
    double a[DIM][NC];
    if (MODE == PAIR) {
        const double2 *pi = reinterpret_cast<const double2 *>(base + (size_t)vi * S);
        const double2 *pj = JROW_SMEM ? reinterpret_cast<const double2 *>(jrow)
                                      : reinterpret_cast<const double2 *>(base + (size_t)vj * S);
        double *af = &a[0][0];
#pragma unroll
        for (int q = 0; q < S / 2; ++q) {                      // Bezier.sub
            const double2 u = __ldg(pi + q), w = JROW_SMEM ? pj[q] : __ldg(pj + q);
            if (2 * q < DIM * NC) af[2 * q] = u.x - w.x;
            if (2 * q + 1 < DIM * NC) af[2 * q + 1] = u.y - w.y;
        }
    } else {
        const int v = vi;
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

// constraints_mma.cu: fp64 tensor path for L <= 128, degree <= 15, dim 2 / 3
bool bez_sq_elev_mma_supported(const bez_plan *plan);
bool bez_sq_elev_mma_wide_supported(const bez_plan *plan);
int bez_sq_elev_mma(const bez_plan *plan, const SqElevArgs &A, int mode, cudaStream_t st);
int bez_sq_elev_mma_flags();

}  // namespace bezcore
