// Finite-difference Jacobian columns of the separation and speed constraints
// (A7 of SURVEY.md section 8) for sm_100a.
//
// SciPy's SLSQP forms  J[:,k] = (f(x + h_k e_k) - f(x)) / dx_k,  dx_k = (x_k+h_k) - x_k
// with nvar+1 calls of the constraint callable (scipy/optimize/_slsqp_py.py:349-367,
// _numdiff.py:585-596, 683-712).  Both constraints are quadratic forms of the control
// points, which are affine in x (reshapeVector), so the quotient has the closed form
//     J[:,k] = elev( scale * B(2a + dx*delta, delta) )
// where a is the curve being squared (c_i - c_j, or the derivative curve), delta =
// d a / d x_k and B is the Bernstein product.  Evaluating that directly is free of the
// cancellation that limits the literal difference to ~1e-7 relative accuracy, costs the
// same fused "square -> fold -> elevate" pipeline as one constraint row, and only
// touches the rows that depend on x_k:
//   J_PAIR_VAR   item = (variable kk of vehicle v, partner u != v)     N-1 rows per variable
//   J_PAIR_DIR   item = pair p, for a variable that moves every curve (tf of time-optimal
//                Dubins problems): delta = dir_i - dir_j, dir = d y / d tf
//   J_SPEED_VAR  item = variable kk -> the speed row of its vehicle
//   J_SPEED_DIR  item = vehicle v, tf variable (moves control points and the n/tf factor)
// Output either in the sweep layout [item][L] or scattered into a dense J^T [nvar][ld]
// (rows of J^T are contiguous, so stores stay coalesced; the host hands SciPy J^T.T).
#include <stdlib.h>

#include "sq_elev_core.cuh"
#include "sq_elev_mma.cuh"
#include "sq_elev_ws.cuh"

namespace {
using namespace bezcore;

enum JMode { J_PAIR_VAR = 0, J_PAIR_DIR = 1, J_SPEED_VAR = 2, J_SPEED_DIR = 3 };

template <int N_>
struct FullWeights { double w[(N_ + 1) * (N_ + 1)]; };

struct JacArgs {
    int early_store;      // A/B: proxy fence + bulk store right behind each m-tile (BEZGPU_MMA_FLAGS bit 8)
    const double *cpts;   // [N][S] base point
    const double *dir;    // [N][S] d y / d x_kdir (DIR modes)
    const double *dx;     // [nvar]
    const double *PQ;     // folded elevation table
    double *out;
    long long nitems;
    long long ld;         // dense: row length of J^T
    int N, numVeh, L, Lh, LhPad;
    int ncols, offset;    // free columns per row of x, first free control point
    int kdir;             // variable index served by the DIR modes
    int dense;
    double tf;            // SPEED modes
    double scale;         // alpha * dim / 2
};

// Stage 1 of the Jacobian kernels for one item: s = B(2a + dx*delta, delta) (unscaled) and
// the offset of the item's output row.
// J_PAIR_VAR item = (variable kk, partner curve u): owner vehicle v, dimension d, control point c
struct PairVarItem { long long kk; int v, d, c, u; };
__device__ __forceinline__ PairVarItem pair_var_item(const JacArgs &A, int dim, long long item) {
    PairVarItem it;
    const int dim_ncols = dim * A.ncols;
    it.kk = item / (A.N - 1);
    const int uu = (int)(item - it.kk * (A.N - 1));
    it.v = (int)(it.kk / dim_ncols);
    const int rem = (int)(it.kk - (long long)it.v * dim_ncols);
    it.d = rem / A.ncols;
    it.c = A.offset + (rem - it.d * A.ncols);
    it.u = uu + (uu >= it.v ? 1 : 0);
    return it;
}

// urow (J_PAIR_VAR only, optional): the partner curve's whole row, already in shared memory.
// ONEHOT + scratch (J_PAIR_VAR only): 2n+1 doubles of shared memory private to the lane (may alias
// the urow region of the warp: every lane has read its row before anyone writes).  With it the
// Bernstein product uses that delta = +-e_c is one-hot: s[i + c] = w[i][c] * (u_i * sg) -- the same
// bits as the dense double loop (all its other terms add +-0), 22 instead of 242 fp64 instructions.
template <int N_, int DIM, int JMODE, bool ONEHOT = false>
__device__ __forceinline__ void jac_stage1(const JacArgs &A, const FullWeights<N_> &FW,
                                           const DiffWeights<N_> &DW, long long item,
                                           double (&s)[2 * N_ + 1], long long &ro, const double *urow = nullptr,
                                           double *scratch = nullptr) {
    constexpr int NC = N_ + 1;
    constexpr int S = (DIM * NC + 1) / 2 * 2;
    constexpr bool DIRMODE = (JMODE == J_PAIR_DIR || JMODE == J_SPEED_DIR);
    constexpr int ND = DIRMODE ? DIM : 1;
    const int dim_ncols = DIM * A.ncols;
    double a[ND][NC], dl[ND][NC];
    double dxk;
    int c_hot = 0;              // J_PAIR_VAR: delta = sg_hot * e_{c_hot}
    double sg_hot = 0.0;
    if (JMODE == J_PAIR_VAR) {
        const PairVarItem it = pair_var_item(A, DIM, item);
        const long long kk = it.kk;
        const int v = it.v, d = it.d, c = it.c, u = it.u;
        const int vi = v < u ? v : u, vj = v < u ? u : v;
        const double sg = v < u ? 1.0 : -1.0;
        c_hot = c;
        sg_hot = sg;
        const double *pv = A.cpts + (size_t)v * S + d * NC;
        if (urow) {                                 // partner row from shared memory (TMA row fetch)
            const double *pu = urow + d * NC;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const double cv = __ldg(pv + k), cu = pu[k];
                a[0][k] = v < u ? cv - cu : cu - cv;
                dl[0][k] = (k == c) ? sg : 0.0;
            }
        } else {
            const double *pu = A.cpts + (size_t)u * S + d * NC;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const double cv = __ldg(pv + k), cu = __ldg(pu + k);
                a[0][k] = v < u ? cv - cu : cu - cv;
                dl[0][k] = (k == c) ? sg : 0.0;
            }
        }
        dxk = __ldg(A.dx + kk);
        const long long p = bez_pair_row_offset(vi, A.N) + (vj - vi - 1);
        ro = A.dense ? kk * A.ld + p * A.L : item * (long long)A.L;
    } else if (JMODE == J_PAIR_DIR) {
        int vi, vj;
        bez_pair_decode(item, A.N, vi, vj);
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                a[d][k] = __ldg(A.cpts + (size_t)vi * S + d * NC + k) -
                          __ldg(A.cpts + (size_t)vj * S + d * NC + k);
                dl[d][k] = __ldg(A.dir + (size_t)vi * S + d * NC + k) -
                           __ldg(A.dir + (size_t)vj * S + d * NC + k);
            }
        dxk = __ldg(A.dx + A.kdir);
        ro = A.dense ? (long long)A.kdir * A.ld + item * A.L : item * (long long)A.L;
    } else if (JMODE == J_SPEED_VAR) {
        const long long kk = item;
        const int v = (int)(kk / dim_ncols);
        const int rem = (int)(kk - (long long)v * dim_ncols);
        const int d = rem / A.ncols;
        const int c = A.offset + (rem - d * A.ncols);
        const double val = (double)N_ / A.tf;
        const double *pv = A.cpts + (size_t)v * S + d * NC;
        double pt[NC], dd[NC], de[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) pt[k] = __ldg(pv + k);
#pragma unroll
        for (int k = 0; k < N_; ++k) {
            dd[k] = pt[k] * (-val) + pt[k + 1] * val;
            de[k] = ((k == c) ? -val : 0.0) + ((k + 1 == c) ? val : 0.0);
        }
        dd[N_] = 0.0; de[N_] = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            double q = dd[k] * DW.lo[k], r = de[k] * DW.lo[k];
            if (k > 0) { q = dd[k - 1] * DW.hi[k] + q; r = de[k - 1] * DW.hi[k] + r; }
            a[0][k] = q;
            dl[0][k] = r;
        }
        dxk = __ldg(A.dx + kk);
        ro = A.dense ? kk * A.ld + (long long)v * A.L : item * (long long)A.L;
    } else {   // J_SPEED_DIR: tf moves the control points (dir) and the n/tf factor
        const int v = (int)item;
        dxk = __ldg(A.dx + A.kdir);
        const double val = (double)N_ / A.tf;
        const double valp = (double)N_ / (A.tf + dxk);
        const double gam = -(double)N_ / (A.tf * (A.tf + dxk));     // (valp - val) / dx
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double pt[NC], pd[NC], dd[NC], de[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                pt[k] = __ldg(A.cpts + (size_t)v * S + d * NC + k);
                pd[k] = __ldg(A.dir + (size_t)v * S + d * NC + k);
            }
#pragma unroll
            for (int k = 0; k < N_; ++k) {
                dd[k] = pt[k] * (-val) + pt[k + 1] * val;
                de[k] = valp * (pd[k + 1] - pd[k]) + gam * (pt[k + 1] - pt[k]);
            }
            dd[N_] = 0.0; de[N_] = 0.0;
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                double q = dd[k] * DW.lo[k], r = de[k] * DW.lo[k];
                if (k > 0) { q = dd[k - 1] * DW.hi[k] + q; r = de[k - 1] * DW.hi[k] + r; }
                a[d][k] = q;
                dl[d][k] = r;
            }
        }
        ro = A.dense ? (long long)A.kdir * A.ld + (long long)v * A.L : item * (long long)A.L;
    }

    if (ONEHOT && JMODE == J_PAIR_VAR) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k <= 2 * N_; ++k) scratch[k] = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const double ui = fma(dxk, dl[0][i], 2.0 * a[0][i]);
            scratch[i + c_hot] = fma(FW.w[i * NC + c_hot], ui * sg_hot, 0.0);
        }
#pragma unroll
        for (int k = 0; k <= 2 * N_; ++k) s[k] = scratch[k];
        return;
    }
    // s = B(2a + dx*delta, delta)   (Bernstein product, full weight matrix)
#pragma unroll
    for (int k = 0; k <= 2 * N_; ++k) s[k] = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        double uvec[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) uvec[i] = fma(dxk, dl[d][i], 2.0 * a[d][i]);
#pragma unroll
        for (int i = 0; i < NC; ++i)
#pragma unroll
            for (int j = 0; j < NC; ++j)
                s[i + j] = fma(FW.w[i * NC + j], uvec[i] * dl[d][j], s[i + j]);
    }
}

template <int N_, int DIM, int JMODE>
__global__ void __launch_bounds__(kThreads, 3)
jac_sq_elev_kernel(const JacArgs A, const FullWeights<N_> FW, const DiffWeights<N_> DW) {
    constexpr int NC = N_ + 1;
    constexpr int NT = 2 * N_ + 1;
    constexpr int RS = RowGeom<N_>::RS;
    constexpr int CPL = 2;
    extern __shared__ __align__(16) double smem[];
    // layout: [kWarps][32][RS] double2 rows | [NT][LhPad] table | [kWarps][32] row offsets
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2 *rows = reinterpret_cast<double2 *>(smem) + (size_t)warp * 32 * RS;
    double *tab = smem + (size_t)kWarps * 32 * RS * 2;
    long long *rowoff = reinterpret_cast<long long *>(tab + (size_t)NT * A.LhPad) + warp * 32;

    for (int i = threadIdx.x; i < NT * A.LhPad; i += kThreads) tab[i] = __ldg(A.PQ + i);
    __syncthreads();

    const long long nwt = (A.nitems + 31) >> 5;
    const long long gwarp = (long long)blockIdx.x * kWarps + warp;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const int ngroups = (A.LhPad / 32 + CPL - 1) / CPL;

    for (long long wt = gwarp; wt < nwt; wt += nwarps) {
        const long long t0 = wt << 5;
        const int cnt = (int)((A.nitems - t0) < 32 ? (A.nitems - t0) : 32);
        {
            const long long item = t0 + (lane < cnt ? lane : cnt - 1);
            double s[2 * N_ + 1];
            long long ro;
            jac_stage1<N_, DIM, JMODE>(A, FW, DW, item, s, ro);
            rowoff[lane] = ro;
            double2 *row = rows + (size_t)lane * RS;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                const double lo = s[j] * A.scale, hi = s[2 * N_ - j] * A.scale;
                row[j] = make_double2(lo + hi, lo - hi);
            }
            row[N_] = make_double2(s[N_] * A.scale, 0.0);
        }
        __syncwarp();
        for (int g = 0; g < ngroups; ++g) {
            if (cnt == 32)
                sweep_columns<N_, CPL, false, true, true>(rows, tab, A.out, g, lane, 32, A.L, A.Lh, A.LhPad, 0.0, rowoff);
            else
                sweep_columns<N_, CPL, false, false, true>(rows, tab, A.out, g, lane, cnt, A.L, A.Lh, A.LhPad, 0.0, rowoff);
        }
        __syncwarp();
    }
}

// Sweep-layout variant on the fp64 tensor path (sq_elev_mma.cuh): the rows of consecutive
// items are contiguous, so the DMMA stage 2 + TMA bulk-store epilogue of the constraint
// kernels applies unchanged.  Used for the per-variable modes (one dimension live in stage 1).
template <int N_, int DIM, int JMODE>
__global__ void __launch_bounds__(kThreads, 2)
jac_sq_elev_mma_kernel(const JacArgs A, const FullWeights<N_> FW, const DiffWeights<N_> DW) {
    using namespace bezmma;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t per_warp = (size_t)kRowsDoubles + 16 * (size_t)A.L;
    double *rows = smem + warp * per_warp;
    double *obuf = rows + kRowsDoubles;
    for (int i = lane; i < kRowsDoubles; i += 32) rows[i] = 0.0;     // padding slots must be 0
    const unsigned obuf_s = (unsigned)__cvta_generic_to_shared(obuf);
    BFrags<N_, 4> Bf;
    load_bfrags<N_, 4>(Bf, A.PQ, A.L, A.LhPad, lane);
    MinSinks no_sinks;
    memset(&no_sinks, 0, sizeof(no_sinks));
    __syncwarp();
    const long long nwt = (A.nitems + 31) >> 5;
    const long long gwarp = (long long)blockIdx.x * kWarps + warp;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(A.out) & 15u) == 0;
    for (long long wt = gwarp; wt < nwt; wt += nwarps) {
        const long long t0 = wt << 5;
        const int cnt = (int)((A.nitems - t0) < 32 ? (A.nitems - t0) : 32);
        {
            double s[2 * N_ + 1];
            long long ro;
            jac_stage1<N_, DIM, JMODE, JMODE == J_PAIR_VAR>(A, FW, DW, t0 + (lane < cnt ? lane : cnt - 1), s, ro, nullptr,
                                                            rows + lane * kRowStride);
            double *row = rows + lane * kRowStride;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                const double lo = s[j] * A.scale, hi = s[2 * N_ - j] * A.scale;
                row[slot_e(j)] = lo + hi;
                row[slot_o(j)] = lo - hi;
            }
            row[slot_e(N_)] = s[N_] * A.scale;
            if (JMODE == J_PAIR_VAR) {       // the scratch use of the row overwrote the padding slots of the k-steps
#pragma unroll
                for (int j = N_ + 1; j < 4 * Geom<N_>::KE; ++j) row[slot_e(j)] = 0.0;
#pragma unroll
                for (int j = N_; j < 4 * Geom<N_>::KO; ++j) row[slot_o(j)] = 0.0;
            }
        }
        __syncwarp();
        mma_tile<N_, 4, 0, true>(rows, obuf, obuf_s, Bf, A.out + (size_t)t0 * A.L, no_sinks, t0, cnt, A.L, 0.0, lane,
                                 base_aligned, A.early_store != 0);
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
}

// The same sweep kernel, warp-specialised like the pair kernel (sq_elev_ws.cuh): per scheduler one
// producer warp (stage 1 only) feeds two consumer warps (DMMA stage 2 + TMA bulk stores) through a
// ring of three row slots; no TMA row fetch here, the items of a tile read scattered rows.
template <int N_, int DIM, int JMODE>
__global__ void __launch_bounds__(bezws::kWsThreads, 1)
jac_sq_elev_ws_kernel(const JacArgs A, const FullWeights<N_> FW, const DiffWeights<N_> DW) {
    using namespace bezmma;
    using namespace bezws;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int group = warp & 3;
    constexpr bool TMAROWS = (JMODE == J_PAIR_VAR);
    constexpr int S_ = (DIM * (N_ + 1) + 1) / 2 * 2;
    constexpr int kSlotD = TMAROWS ? slot_doubles<N_, DIM>() : kRowsDoubles;
    double *slots = smem + (size_t)group * kSlots * kSlotD;
    double *stag0 = smem + (size_t)4 * kSlots * kSlotD;
    const size_t stag_per = 16 * (size_t)A.L;
    const unsigned bars_s = (unsigned)__cvta_generic_to_shared(stag0 + kConsumers * stag_per) + 8u * kBarsPerGroup * group;
    auto full_b = [&](int s) { return bars_s + 8u * s; };
    auto empty_b = [&](int s) { return bars_s + 8u * (kSlots + s); };
    auto rows_b = [&](int s) { return bars_s + 8u * (2 * kSlots + s); };
    if (warp < kProducers) {
        for (int i = lane; i < kSlots * kSlotD; i += 32) slots[i] = 0.0;           // padding slots must be 0
        if (lane == 0) {
            for (int s = 0; s < kSlots; ++s) { mbar_init(full_b(s), 32); mbar_init(empty_b(s), 32); mbar_init(rows_b(s), 1); }
            mbar_fence_init();
        }
    }
    __syncthreads();
    const long long nwt = (A.nitems + 31) >> 5;
    const long long ngroups = (long long)gridDim.x * 4, gidx = (long long)blockIdx.x * 4 + group;
    const long long t_begin = gidx * nwt / ngroups;
    const int n = (int)((gidx + 1) * nwt / ngroups - t_begin);
    if (warp < kProducers) {
        reg_dec<120>();
        for (int k = 0; k < n; ++k) {
            const int s = k % kSlots, use = k / kSlots;
            const long long t0 = (t_begin + k) << 5;
            const int cnt = (int)((A.nitems - t0) < 32 ? (A.nitems - t0) : 32);
            double sc[2 * N_ + 1];
            long long ro;
            const long long item = t0 + (lane < cnt ? lane : cnt - 1);
            double *slot = slots + (size_t)s * kSlotD;
            if (TMAROWS) {
                // the partner rows of a tile are consecutive rows of the control-point array (one run
                // per variable, broken once at the owner vehicle): one TMA bulk copy per run
                const PairVarItem it = pair_var_item(A, DIM, item);
                // the owner's row and the step length go to L1 while the slot is awaited and the rows fly
                asm volatile("prefetch.global.L1 [%0];" :: "l"(A.cpts + (size_t)it.v * S_ + it.d * (N_ + 1)));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(A.cpts + (size_t)it.v * S_ + it.d * (N_ + 1) + N_));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(A.dx + it.kk));
                if (use > 0) mbar_wait(empty_b(s), (unsigned)(use - 1) & 1u);
                const int pu = __shfl_up_sync(0xffffffffu, it.u, 1);
                const bool start = lane == 0 || it.u != pu + 1;
                const unsigned runs = __ballot_sync(0xffffffffu, start);
                const unsigned slot_s = (unsigned)__cvta_generic_to_shared(slot);
                fence_async_smem();
                if (lane == 0) mbar_arrive_expect_tx(rows_b(s), 32u * S_ * 8u);
                __syncwarp();
                if (start) {
                    const unsigned higher = lane == 31 ? 0u : (runs & (0xffffffffu << (lane + 1)));
                    const int end = higher ? __ffs(higher) - 1 : 32;
                    // a run never leaves the array: rows u .. u + len - 1 <= N - 1 (dead lanes repeat the last item)
                    bulk_load(slot_s + (unsigned)lane * (S_ * 8u), A.cpts + (size_t)it.u * S_,
                              (unsigned)(end - lane) * (S_ * 8u), rows_b(s));
                }
                mbar_wait(rows_b(s), (unsigned)use & 1u);
                jac_stage1<N_, DIM, JMODE, true>(A, FW, DW, item, sc, ro, slot + lane * S_, slot + lane * kRowStride);
                __syncwarp();
            } else {
                jac_stage1<N_, DIM, JMODE>(A, FW, DW, item, sc, ro);
                if (use > 0) mbar_wait(empty_b(s), (unsigned)(use - 1) & 1u);
            }
            double *row = slot + lane * kRowStride;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                const double lo = sc[j] * A.scale, hi = sc[2 * N_ - j] * A.scale;
                row[slot_e(j)] = lo + hi;
                row[slot_o(j)] = lo - hi;
            }
            row[slot_e(N_)] = sc[N_] * A.scale;
            if (TMAROWS) {
#pragma unroll
                for (int j = N_ + 1; j < 4 * Geom<N_>::KE; ++j) row[slot_e(j)] = 0.0;
#pragma unroll
                for (int j = N_; j < 4 * Geom<N_>::KO; ++j) row[slot_o(j)] = 0.0;
            }
            mbar_arrive(full_b(s));
        }
    } else {
        reg_inc<192>();
        const int cidx = warp - kProducers, ci = cidx >> 2;
        double *obuf = stag0 + (size_t)cidx * stag_per;
        const unsigned obuf_s = (unsigned)__cvta_generic_to_shared(obuf);
        const bool base_aligned = (reinterpret_cast<uintptr_t>(A.out) & 15u) == 0;
        BFrags<N_, 4> Bf;
        load_bfrags<N_, 4>(Bf, A.PQ, A.L, A.LhPad, lane);
        MinSinks no_sinks;
        memset(&no_sinks, 0, sizeof(no_sinks));
        for (int k = ci; k < n; k += 2) {
            const int s = k % kSlots, use = k / kSlots;
            const long long t0 = (t_begin + k) << 5;
            const int cnt = (int)((A.nitems - t0) < 32 ? (A.nitems - t0) : 32);
            mbar_wait(full_b(s), (unsigned)use & 1u);
            auto release = [&]() { mbar_arrive(empty_b(s)); };
            mma_tile<N_, 4, 0, true, decltype(release)>(slots + (size_t)s * kSlotD, obuf, obuf_s, Bf,
                                                        A.out + (size_t)t0 * A.L, no_sinks, t0, cnt, A.L, 0.0, lane,
                                                        base_aligned, false, release);
            __syncwarp();
        }
        if (lane == 0) bulk_wait_all();
    }
}

template <int N_, int DIM, int JMODE> size_t jac_ws_shmem(int L) {
    const size_t slot = (JMODE == J_PAIR_VAR) ? bezws::slot_doubles<N_, DIM>() : bezmma::kRowsDoubles;
    return ((size_t)4 * bezws::kSlots * slot + (size_t)bezws::kConsumers * 16 * L + 4 * bezws::kBarsPerGroup) * sizeof(double);
}

template <int N_, int DIM, int JMODE>
int launch_jac_ws(const bez_plan *plan, const JacArgs &A, cudaStream_t st) {
    FullWeights<N_> FW;
    DiffWeights<N_> DW;
    for (int i = 0; i < (N_ + 1) * (N_ + 1); ++i) FW.w[i] = plan->h_W[i];
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t shmem = jac_ws_shmem<N_, DIM, JMODE>(A.L);
    auto kern = jac_sq_elev_ws_kernel<N_, DIM, JMODE>;
    int sms = 148, per_sm = 1;
    if (int rc = bez_kernel_config((const void *)kern, bezws::kWsThreads, shmem, &sms, &per_sm)) return rc;
    const long long nwt = (A.nitems + 31) / 32;
    long long grid = sms;
    if (grid > (nwt + 7) / 8) grid = (nwt + 7) / 8;
    if (grid < 1) return BEZ_OK;
    kern<<<(unsigned)grid, bezws::kWsThreads, shmem, st>>>(A, FW, DW);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

template <int N_, int DIM, int JMODE>
int launch_jac_mma(const bez_plan *plan, const JacArgs &A, cudaStream_t st) {
    FullWeights<N_> FW;
    DiffWeights<N_> DW;
    for (int i = 0; i < (N_ + 1) * (N_ + 1); ++i) FW.w[i] = plan->h_W[i];
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t ws_shmem = jac_ws_shmem<N_, DIM, JMODE>(A.L);
    // Per-variable separation sweep (the 27 GB block of the C4 Jacobian): warp-specialised with a TMA fetch of the
    // partner rows (tools/prof_jac.py: 4.72-4.76 ms in bursts of four launches, 4.39 ms = 0.96 of the HBM peak for an
    // isolated launch under ncu; the 8-warp kernel 4.90 ms).  Without the row fetch (the other modes) one producer
    // cannot feed two consumers -- its stage 1 waits for scattered global loads tile after tile (7.1 ms) -- so they
    // stay on the 8-warp kernel.  BEZGPU_MMA_FLAGS bit 128 keeps everything there (A/B).
    if (JMODE == J_PAIR_VAR && ws_shmem <= 232448 && !(bez_sq_elev_mma_flags() & kFlagNoWarpSpecialisation))
        return launch_jac_ws<N_, DIM, JMODE>(plan, A, st);
    const size_t shmem = (size_t)kWarps * (bezmma::kRowsDoubles + 16 * (size_t)A.L) * sizeof(double);
    auto kern = jac_sq_elev_mma_kernel<N_, DIM, JMODE>;
    int sms = 148, per_sm = 1;
    if (int rc = bez_kernel_config((const void *)kern, kThreads, shmem, &sms, &per_sm)) return rc;
    const long long nwt = (A.nitems + 31) / 32;
    long long grid = (long long)sms * per_sm;
    const long long need = (nwt + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    if (grid < 1) return BEZ_OK;
    kern<<<(unsigned)grid, kThreads, shmem, st>>>(A, FW, DW);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

template <int N_, int DIM, int JMODE>
int launch_jac(const bez_plan *plan, const JacArgs &A, cudaStream_t st) {
    FullWeights<N_> FW;
    DiffWeights<N_> DW;
    for (int i = 0; i < (N_ + 1) * (N_ + 1); ++i) FW.w[i] = plan->h_W[i];
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t shmem = ((size_t)kWarps * 32 * RowGeom<N_>::RS * 2 + (size_t)(2 * N_ + 1) * A.LhPad) * sizeof(double) +
                         (size_t)kWarps * 32 * sizeof(long long);
    if (shmem > 227 * 1024) {
        bez_set_error("degree %d with elevation %d needs %zu bytes of shared memory (> 227 KB)",
                      plan->n, plan->elev, shmem);
        return BEZ_EUNSUPPORTED;
    }
    auto kern = jac_sq_elev_kernel<N_, DIM, JMODE>;
    int sms = 148, per_sm = 1;
    if (int rc = bez_kernel_config((const void *)kern, kThreads, shmem, &sms, &per_sm)) return rc;
    const long long nwt = (A.nitems + 31) / 32;
    long long grid = (long long)sms * per_sm;
    const long long need = (nwt + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    if (grid < 1) return BEZ_OK;
    kern<<<(unsigned)grid, kThreads, shmem, st>>>(A, FW, DW);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

template <int N_, int JMODE>
int jdispatch_dim(const bez_plan *plan, const JacArgs &A, cudaStream_t st) {
    if constexpr (JMODE == J_PAIR_VAR || JMODE == J_SPEED_VAR) {
        const char *e = getenv("BEZGPU_FORCE_DFMA");
        if (!A.dense && A.Lh > 32 && A.Lh <= 64 && !(e && e[0] == '1')) {
            switch (plan->dim) {
                case 1: return launch_jac_mma<N_, 1, JMODE>(plan, A, st);
                case 2: return launch_jac_mma<N_, 2, JMODE>(plan, A, st);
                case 3: return launch_jac_mma<N_, 3, JMODE>(plan, A, st);
            }
        }
    }
    switch (plan->dim) {
        case 1: return launch_jac<N_, 1, JMODE>(plan, A, st);
        case 2: return launch_jac<N_, 2, JMODE>(plan, A, st);
        case 3: return launch_jac<N_, 3, JMODE>(plan, A, st);
    }
    return BEZ_EUNSUPPORTED;
}

template <int JMODE>
int jdispatch(const bez_plan *plan, const JacArgs &A, cudaStream_t st) {
    switch (plan->n) {
#define CASE(n_) case n_: return jdispatch_dim<n_, JMODE>(plan, A, st);
#ifdef BEZ_ONLY_N   /* development builds: one degree only (fast compile) */
        CASE(BEZ_ONLY_N)
#else
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        CASE(9) CASE(10) CASE(11) CASE(12)
#endif
#undef CASE
    }
    bez_set_error("degree %d has no Jacobian kernel instantiation (1..12 supported)", plan->n);
    return BEZ_EUNSUPPORTED;
}

int fill_common(const bez_plan *plan, JacArgs &A, const double *d_cpts, const double *d_dir,
                const double *d_dx, int N, int numVeh, int ncols, int offset, int kdir, int dense,
                double *d_out, int64_t ld) {
    A.cpts = d_cpts; A.dir = d_dir; A.dx = d_dx; A.PQ = plan->d_PQ; A.out = d_out;
    A.ld = ld; A.N = N; A.numVeh = numVeh; A.L = plan->L; A.Lh = plan->Lh; A.LhPad = plan->LhPad;
    A.ncols = ncols; A.offset = offset; A.kdir = kdir; A.dense = dense; A.tf = 1.0; A.scale = 1.0;
    A.early_store = (bez_sq_elev_mma_flags() & kFlagEarlyFence) ? 1 : 0;
    return BEZ_OK;
}

// literal 2-point quotient for constraints without a closed form (angular rate):
//   JT[k][r] = (F[k+1][r] - F[0][r]) / dx[k]      (_numdiff.py:709-711)
__global__ void fd_quotient_kernel(const double *__restrict__ F, const double *__restrict__ dx,
                                   int nvar, long long m, double *__restrict__ JT) {
    const long long total = (long long)nvar * m;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long k = idx / m, r = idx - k * m;
        JT[idx] = __ddiv_rn(__dsub_rn(F[(k + 1) * m + r], F[r]), __ldg(dx + k));
    }
}

}  // namespace

static __global__ void fd_quotient_batched_kernel(const double *__restrict__ F, const double *__restrict__ dx,
                                           long long count, int nvar, long long m, double *__restrict__ JT) {
    const long long per = (long long)nvar * m, total = count * per;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / per, rem = idx - p * per;
        const long long k = rem / m, r = rem - k * m;
        const double *Fp = F + p * (nvar + 1) * m;
        JT[idx] = __ddiv_rn(__dsub_rn(Fp[(k + 1) * m + r], Fp[r]), __ldg(dx + p * nvar + k));
    }
}

extern "C" int bez_fd_quotient_batched(const double *d_F, const double *d_dx, int64_t count, int nvar,
                                       int64_t m, double *d_JT, void *stream) {
    BEZ_REQUIRE(d_F && d_dx && d_JT, "NULL argument");
    BEZ_REQUIRE(count >= 0 && nvar >= 0 && m >= 0, "negative size");
    if (count == 0 || nvar == 0 || m == 0) return BEZ_OK;
    long long blocks = ((long long)count * nvar * m + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    fd_quotient_batched_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_F, d_dx, count, nvar, m, d_JT);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_fd_quotient(const double *d_F, const double *d_dx, int nvar, int64_t m,
                               double *d_JT, void *stream) {
    BEZ_REQUIRE(d_F && d_dx && d_JT, "NULL argument");
    BEZ_REQUIRE(nvar >= 0 && m >= 0, "negative size");
    if (nvar == 0 || m == 0) return BEZ_OK;
    long long blocks = ((long long)nvar * m + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    fd_quotient_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_F, d_dx, nvar, m, d_JT);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_jac_sepsq_elev(const bez_plan *plan, const double *d_cpts, int N, int numVeh,
                                  int ncols, int offset, const double *d_dx, const double *d_dir,
                                  int kdir, int dense, double *d_out, int64_t ld, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_dx && d_out, "NULL argument");
    BEZ_REQUIRE(N >= 2 && numVeh >= 1 && numVeh <= N && ncols >= 0 && offset >= 0, "bad sizes");
    BEZ_REQUIRE(kdir < 0 || d_dir, "direction rows are NULL");
    const long long P = (long long)N * (N - 1) / 2;
    const long long nvarN = (long long)numVeh * plan->dim * ncols;
    BEZ_REQUIRE(!dense || ld >= P * plan->L, "ld smaller than the constraint block");
    BEZ_ON_DEVICE(plan->device);
    cudaStream_t st = (cudaStream_t)stream;
    JacArgs A;
    fill_common(plan, A, d_cpts, d_dir, d_dx, N, numVeh, ncols, offset, kdir, dense, d_out, ld);
    A.scale = 0.5 * plan->dim;
    if (dense) {
        const long long nvar = nvarN + (kdir >= 0 ? 1 : 0);
        BEZ_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)nvar * (size_t)ld, st));
    }
    A.nitems = nvarN * (N - 1);
    if (A.nitems > 0) {
        int rc = jdispatch<J_PAIR_VAR>(plan, A, st);
        if (rc != BEZ_OK) return rc;
    }
    if (kdir >= 0) {
        if (!dense) A.out = d_out + (size_t)A.nitems * plan->L;
        A.nitems = P;
        return jdispatch<J_PAIR_DIR>(plan, A, st);
    }
    return BEZ_OK;
}

extern "C" int bez_jac_speed_sq_elev(const bez_plan *plan, const double *d_cpts, int N, int numVeh,
                                     int ncols, int offset, double tf, double alpha,
                                     const double *d_dx, const double *d_dir, int kdir, int dense,
                                     double *d_out, int64_t ld, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_dx && d_out, "NULL argument");
    BEZ_REQUIRE(N >= 1 && numVeh >= 1 && numVeh <= N && ncols >= 0 && offset >= 0, "bad sizes");
    BEZ_REQUIRE(kdir < 0 || d_dir, "direction rows are NULL");
    const long long nvarN = (long long)numVeh * plan->dim * ncols;
    BEZ_REQUIRE(!dense || ld >= (long long)numVeh * plan->L, "ld smaller than the constraint block");
    BEZ_ON_DEVICE(plan->device);
    cudaStream_t st = (cudaStream_t)stream;
    JacArgs A;
    fill_common(plan, A, d_cpts, d_dir, d_dx, N, numVeh, ncols, offset, kdir, dense, d_out, ld);
    A.scale = alpha * 0.5 * plan->dim;
    A.tf = tf;
    if (dense) {
        const long long nvar = nvarN + (kdir >= 0 ? 1 : 0);
        BEZ_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)nvar * (size_t)ld, st));
    }
    A.nitems = nvarN;
    if (A.nitems > 0) {
        int rc = jdispatch<J_SPEED_VAR>(plan, A, st);
        if (rc != BEZ_OK) return rc;
    }
    if (kdir >= 0) {
        if (!dense) A.out = d_out + (size_t)A.nitems * plan->L;
        A.nitems = numVeh;
        return jdispatch<J_SPEED_DIR>(plan, A, st);
    }
    return BEZ_OK;
}
