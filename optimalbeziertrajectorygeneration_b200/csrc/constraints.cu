// Fused constraint kernels (A0-A5 of SURVEY.md section 8) for sm_100a.
//
//   assemble_cpts_kernel : x -> SoA control points            (reshapeVector)
//   sq_elev_kernel<PAIR> : sub -> normSquare -> elev -> -maxSep^2, all pairs
//   sq_elev_kernel<SPEED>: diff -> normSquare -> elev -> alpha*v+beta, per vehicle
//
// Design of sq_elev_kernel (the C4 hot kernel, 968 B written per pair):
//   stage 1 (thread per item): load the two curves' control points (SoA,
//     coalesced over the vehicle index), form the difference / derivative row
//     a[d][0..n], the Gram sums G[i][j] = sum_d a[d][i] a[d][j] for i <= j and
//     the 2n+1 Bernstein coefficients s_k = (dim/2) * sum_{i+j=k} W[i][j] G[i][j]
//     (weights arrive as by-value kernel parameters -> constant-bank operands).
//     The coefficients are folded into even/odd parts e_j = s_j + s_{2n-j},
//     o_j = s_j - s_{2n-j} and parked in shared memory (one 16-byte aligned
//     row [e0,o0,e1,o1,...,e_n,pad] per item).
//   stage 2 (thread per output column, "column-stationary"): lane i keeps the
//     2n+1 folded elevation weights P[.][i], Q[.][i] of its column in registers
//     and sweeps over the items of the tile, reading each item's (e,o) row with
//     broadcast LDS.128.  One pass yields both b_i and its mirror b_{M-i}
//     (the elevation matrix is centro-symmetric), i.e. (2n+1) DFMA per two
//     outputs instead of 2(2n+1), and each warp emits two 256-byte coalesced
//     streaming stores per item.  Nothing but the final L values per item ever
//     goes to HBM.
#include <stdlib.h>

#include "sq_elev_stage1.cuh"

namespace {
using namespace bezcore;

// One warp = one tile of 32 items.  No block-wide barriers in the main loop: the
// warps of a block only share the (read-only) staged elevation table.
//   CPL = column pairs per lane in one sweep (1 or 2): lane l owns columns
//   {l, l+32}[:CPL] of the current 32*CPL-column group and their mirrors.
template <int N_, int DIM, int MODE, int CPL, bool WITH_MIN>
__global__ void __launch_bounds__(kThreads, 3)
sq_elev_kernel(const SqElevArgs A, const ProdWeights<N_> PW, const DiffWeights<N_> DW) {
    constexpr int NC = N_ + 1;
    constexpr int NT = 2 * N_ + 1;           // table rows
    constexpr int SLOTS = RowGeom<N_>::SLOTS;
    constexpr int RS = RowGeom<N_>::RS;
    static_assert(!WITH_MIN || 2 * RS >= 16, "staged row too short to park 16 partial minima");
    extern __shared__ __align__(16) double smem[];
    // layout: [kWarps][32][RS] double2 item rows | [NT][LhPad] table
    double2 *rows = reinterpret_cast<double2 *>(smem) + (size_t)(threadIdx.x >> 5) * 32 * RS;
    double *tab = smem + (size_t)kWarps * 32 * RS * 2;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int i = tid; i < NT * A.LhPad; i += kThreads) tab[i] = __ldg(A.PQ + i);
    __syncthreads();

    // tiles run over the flattened item list [B][nitems] (output rows are contiguous across
    // evaluation points), so small per-evaluation item counts still fill the warps
    const long long total = A.nitems * (long long)A.B;
    const long long nwt = (total + 31) >> 5;
    const long long gwarp = (long long)blockIdx.x * kWarps + (tid >> 5);
    const long long nwarps = (long long)gridDim.x * kWarps;
    const int M = A.L - 1;
    const int ngroups = (A.LhPad / 32 + CPL - 1) / CPL;     // sweeps over the columns
    const double scale = A.alpha * (0.5 * (double)DIM);       // Q1: dim/2 (sign of alpha folded in)

    for (long long wt = gwarp; wt < nwt; wt += nwarps) {
        const long long g0 = wt << 5;                          // first flattened item of the tile
        const int cnt = (int)((total - g0) < 32 ? (total - g0) : 32);

        // ------------------------- stage 1: lane = item -------------------------
        {
            double s[2 * N_ + 1];
            // lanes past the end recompute the last item so every staged row is finite
            const long long gi = g0 + (lane < cnt ? lane : cnt - 1);
            const int b = (int)(gi / A.nitems);
            const long long it = A.item_begin + gi - (long long)b * A.nitems;
            int vi = (int)it, vj = 0;
            if (MODE == PAIR) bez_pair_decode(it, A.N, vi, vj);
            stage1_coeffs<N_, DIM, MODE>(A, PW, DW, b, vi, vj, s);
            double2 *row = rows + (size_t)lane * RS;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                const double lo = s[j] * scale, hi = s[2 * N_ - j] * scale;
                row[j] = make_double2(lo + hi, lo - hi);
            }
            row[N_] = make_double2(s[N_] * scale, 0.0);
        }
        __syncwarp();

        // ------------------- stage 2: lane = output column(s) -------------------
        double *outb = A.out + (size_t)g0 * A.L;
        for (int g = 0; g < ngroups; ++g) {
            if (cnt == 32)
                sweep_columns<N_, CPL, WITH_MIN, true>(rows, tab, outb, g, lane, 32, A.L, A.Lh, A.LhPad, A.beta);
            else
                sweep_columns<N_, CPL, WITH_MIN, false>(rows, tab, outb, g, lane, cnt, A.L, A.Lh, A.LhPad, A.beta);
        }
        if (WITH_MIN) {
            __syncwarp();
            // lane p reduces the 16 partial minima parked in row p
            const double *mb = reinterpret_cast<const double *>(rows + (size_t)lane * RS);
            double m0 = mb[0], m1 = mb[1], m2 = mb[2], m3 = mb[3];
#pragma unroll
            for (int q = 4; q < 16; q += 4) {
                m0 = dmin_nan(m0, mb[q]); m1 = dmin_nan(m1, mb[q + 1]);
                m2 = dmin_nan(m2, mb[q + 2]); m3 = dmin_nan(m3, mb[q + 3]);
            }
            if (lane < cnt) A.sinks.itemmin[g0 + lane] = dmin_nan(dmin_nan(m0, m1), dmin_nan(m2, m3));
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// reshapeVector (optimization.py:242-285) + obstacle rows (optimization.py:86-94)
struct AssembleArgs {
    const double *x;
    int B, nvar, numVeh, nObs, n, dim, fixed_ends, dubins, timeopt;
    int evals_per_obst_set;   // > 0: evaluation b uses obstacle set b / evals_per_obst_set
    double tf_fixed;
    const double *init, *fin, *ispeed, *fspeed, *icos, *isin, *fcos, *fsin, *obst;
    double *cpts, *tf;
};

// IdxT = unsigned when the element count fits 32 bits (every BASELINE config): the index decomposition
// is four divisions per element, and 64-bit ones made this kernel take 200 us on the single SM the
// persistent pair kernel leaves to the small kernels of the next step (3 us on an empty GPU).
template <typename IdxT>
__global__ void assemble_cpts_kernel(const AssembleArgs A) {
    const int NC = A.n + 1;
    const int N = A.numVeh + A.nObs;
    const int S = (A.dim * NC + 1) / 2 * 2;
    const IdxT total = (IdxT)A.B * (IdxT)N * (IdxT)S;
    const int offset = (A.fixed_ends ? 1 : 0) + (A.dubins ? 1 : 0);
    const int ncols = NC - 2 * offset;
    for (IdxT idx = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (IdxT)gridDim.x * blockDim.x) {
        const int e = (int)(idx % (IdxT)S);
        const IdxT r = idx / (IdxT)S;
        const int v = (int)(r % (IdxT)N);
        const int b = (int)(r / (IdxT)N);
        if (e >= A.dim * NC) { A.cpts[idx] = 0.0; continue; }          // alignment pad
        const int d = e / NC, k = e - d * NC;
        const double *x = A.x + (size_t)b * A.nvar;
        const double tf = A.timeopt ? x[A.nvar - 1] : A.tf_fixed;
        double val;
        if (v >= A.numVeh) {
            const size_t set = A.evals_per_obst_set > 0 ? (size_t)(b / A.evals_per_obst_set) : 0;
            val = A.obst[(set * A.nObs + (v - A.numVeh)) * A.dim + d];    // constant curve (Q8)
        } else if (A.fixed_ends && k == 0) {
            val = A.init[(size_t)v * A.dim + d];
        } else if (A.fixed_ends && k == A.n) {
            val = A.fin[(size_t)v * A.dim + d];
        } else if (A.dubins && k == 1) {
            // initPoints + (initSpeeds*tf/deg) * cos|sin(initAngs): separate roundings as numpy
            const double mag = __ddiv_rn(__dmul_rn(A.ispeed[v], tf), (double)A.n);
            const double cs = (d == 0) ? A.icos[v] : A.isin[v];
            val = __dadd_rn(A.init[(size_t)v * A.dim + d], __dmul_rn(mag, cs));
        } else if (A.dubins && k == A.n - 1) {
            const double mag = __ddiv_rn(__dmul_rn(A.fspeed[v], tf), (double)A.n);
            const double cs = (d == 0) ? A.fcos[v] : A.fsin[v];
            val = __dsub_rn(A.fin[(size_t)v * A.dim + d], __dmul_rn(mag, cs));
        } else {
            val = x[(size_t)(v * A.dim + d) * ncols + (k - offset)];
        }
        A.cpts[idx] = val;
        if (v == 0 && e == 0) A.tf[b] = tf;
    }
}

// ---------------------------------------------------------------------------
template <int N_, int DIM, int MODE, int CPL, bool WITH_MIN>
int launch_sq_elev2(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    ProdWeights<N_> PW;
    DiffWeights<N_> DW;
    for (int i = 0; i <= N_; ++i)
        for (int j = i; j <= N_; ++j) {
            double w = plan->h_W[i * (N_ + 1) + j];
            PW.w[widx<N_>(i, j)] = (i == j) ? w : 2.0 * w;    // W is symmetric; G[i][j]==G[j][i]
        }
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }

    const size_t shmem = ((size_t)kWarps * 32 * RowGeom<N_>::RS * 2 + (size_t)(2 * N_ + 1) * A.LhPad) * sizeof(double);
    if (shmem > 227 * 1024) {
        bez_set_error("degree %d with elevation %d needs %zu bytes of shared memory (> 227 KB)",
                      plan->n, plan->elev, shmem);
        return BEZ_EUNSUPPORTED;
    }
    auto kern = sq_elev_kernel<N_, DIM, MODE, CPL, WITH_MIN>;
    int sms = 148, per_sm = 1;
    if (int rc = bez_kernel_config((const void *)kern, kThreads, shmem, &sms, &per_sm)) return rc;
    const long long nwt = (A.nitems * (long long)A.B + 31) / 32;
    long long grid = (long long)sms * per_sm;
    const long long need = (nwt + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    if (grid < 1) return BEZ_OK;
    kern<<<(unsigned)grid, kThreads, shmem, st>>>(A, PW, DW);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

template <int N_, int DIM, int MODE>
int launch_sq_elev(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    const bool two = A.LhPad > 32;      // more than one 32-column group: 2 column pairs per lane
    // the fused per-item minimum parks partial results in consumed rows, which
    // needs a single sweep over the columns (L <= 128) and rows of >= 16 doubles
    const bool fused_min = A.sinks.itemmin && A.LhPad <= 64 && 2 * RowGeom<N_>::RS >= 16;
    if (fused_min) {
        if constexpr (2 * RowGeom<N_>::RS >= 16)
            return two ? launch_sq_elev2<N_, DIM, MODE, 2, true>(plan, A, st)
                       : launch_sq_elev2<N_, DIM, MODE, 1, true>(plan, A, st);
    }
    int rc = two ? launch_sq_elev2<N_, DIM, MODE, 2, false>(plan, A, st)
                 : launch_sq_elev2<N_, DIM, MODE, 1, false>(plan, A, st);
    if (rc != BEZ_OK || !A.sinks.itemmin) return rc;
    const long long rows = (long long)A.B * A.nitems;         // rare shapes: separate reduction pass
    item_min_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(A.out, rows, A.L, A.sinks.itemmin);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

template <int N_, int MODE>
int dispatch_dim(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    switch (plan->dim) {
        case 1: return launch_sq_elev<N_, 1, MODE>(plan, A, st);
        case 2: return launch_sq_elev<N_, 2, MODE>(plan, A, st);
        case 3: return launch_sq_elev<N_, 3, MODE>(plan, A, st);
    }
    return BEZ_EUNSUPPORTED;
}

template <int MODE>
int dispatch_degree(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    switch (plan->n) {
#define CASE(n_) case n_: return dispatch_dim<n_, MODE>(plan, A, st);
#ifdef BEZ_ONLY_N   /* development builds: one degree only (fast compile) */
        CASE(BEZ_ONLY_N)
#else
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#endif
#undef CASE
    }
    bez_set_error("degree %d has no fused kernel instantiation (1..16 supported)", plan->n);
    return BEZ_EUNSUPPORTED;
}

}  // namespace

// Packed active bitmask / compacted list from a [B][pitch] matrix of per-item minima: the
// post-pass of the shapes whose kernels do not fuse it (the tensor-path kernels do, emit_minima).
__global__ void emit_from_minima_kernel(bezmma::MinSinks S, long long total) {
    const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll;
    const int lane = threadIdx.x & 31;
    if (g0 >= total) return;
    const long long f = g0 + lane;
    const bool valid = f < total;
    long long src = f;
    if (S.min_pitch > 0 && valid) {
        const long long b = f / S.nitems;
        src = b * S.min_pitch + (f - b * S.nitems);
    }
    const double v = valid ? S.itemmin[src] : 0.0;
    const bool act = valid && v < S.threshold;
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (S.mask && lane == 0) S.mask[g0 >> 5] = bal;
    if (S.list_count && bal) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(S.list_count, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        const long long pos = (long long)base + __popc(bal & ((1u << lane) - 1u));
        if (act && pos < S.list_cap) { S.list_idx[pos] = f; S.list_val[pos] = v; }
    }
}

// Compacted (f, min) list from the packed bitmask + the minimum matrix.  Large launches of the tensor-path kernels use this instead of
// appending from the epilogue: the append costs the hot kernel 2 % (an atomic round trip in front
// of a warp-wide shuffle for every tile that has an active item), this pass ~3 us.
__global__ void compact_from_mask_kernel(bezmma::MinSinks S, long long total) {
    // one thread per mask word, ONE atomic per block (256 words): block-wide exclusive scan of the
    // per-word popcounts (same-address atomics serialise in L2: one per word cost more than the
    // fused append it replaces)
    __shared__ unsigned warp_sum[8];
    __shared__ unsigned long long block_base;
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned bits = (w < ((total + 31) >> 5)) ? S.mask[w] : 0u;
    const unsigned cnt = __popc(bits);
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { const unsigned v = warp_sum[i]; warp_sum[i] = tot; tot += v; }
        block_base = tot ? atomicAdd(S.list_count, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    long long pos = (long long)block_base + warp_sum[warp] + (incl - cnt);
    while (bits) {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        const long long f = (w << 5) + bit;
        long long src = f;
        if (S.min_pitch > 0) {
            const long long b = f / S.nitems;
            src = b * S.min_pitch + (f - b * S.nitems);
        }
        if (pos < S.list_cap) { S.list_idx[pos] = f; S.list_val[pos] = S.itemmin[src]; }
        ++pos;
    }
}

static int fill_sinks(bezmma::MinSinks &S, const bez_reduce_opts *o, long long nitems, double *legacy_min) {
    memset(&S, 0, sizeof(S));
    S.nitems = nitems;
    S.itemmin = legacy_min;
    if (!o) return BEZ_OK;
    if (o->itemmin) S.itemmin = o->itemmin;
    BEZ_REQUIRE(o->min_pitch == 0 || o->min_pitch >= nitems, "min_pitch is smaller than the item count");
    S.min_pitch = (o->min_pitch == nitems) ? 0 : o->min_pitch;
    BEZ_REQUIRE(o->npeers >= 0 && o->npeers <= BEZ_MAX_PEERS && (o->npeers == 0 || o->peer_min), "bad peer list");
    S.npeers = o->npeers;
    for (int q = 0; q < o->npeers; ++q) S.peer_min[q] = reinterpret_cast<double *>((uintptr_t)o->peer_min[q]);
    S.mask = o->active_mask;
    S.threshold = o->threshold;
    BEZ_REQUIRE(!o->list_count || (o->list_idx && o->list_val && o->list_cap >= 0), "incomplete active list");
    S.list_count = (unsigned long long *)o->list_count;
    S.list_idx = (long long *)o->list_idx;
    S.list_val = o->list_val;
    S.list_cap = o->list_cap;
    return BEZ_OK;
}

// common tail of the pair / speed entry points
static int run_sq_elev(const bez_plan *plan, SqElevArgs &A, int mode, const bez_reduce_opts *opts,
                       double *legacy_min, cudaStream_t st) {
    if (int rc = fill_sinks(A.sinks, opts, A.nitems, legacy_min)) return rc;
    const bezmma::MinSinks &S = A.sinks;
    const bool derived = S.mask || S.list_count;
    BEZ_REQUIRE(A.out || S.itemmin || derived || S.npeers > 0, "nothing to compute: no rows and no minima requested");
    A.flags = bez_sq_elev_mma_flags();
    if ((bez_sq_elev_mma_supported(plan) && (A.out || mode == PAIR)) ||
        (bez_sq_elev_mma_wide_supported(plan) && A.out)) {
        const long long total = A.nitems * (long long)A.B;
        const bool defer_list = S.list_count && S.mask && S.itemmin && total >= (1 << 16) &&
                                !(A.flags & kFlagFusedList);
        unsigned long long *list_count = A.sinks.list_count;
        if (defer_list) A.sinks.list_count = nullptr;          // the kernel still packs the bitmask
        int rc = bez_sq_elev_mma(plan, A, mode, st);
        A.sinks.list_count = list_count;
        if (rc != BEZ_OK || !defer_list) return rc;
        const long long words = (total + 31) >> 5;
        compact_from_mask_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(A.sinks, total);
        BEZ_CUDA(cudaGetLastError());
        return BEZ_OK;
    }
    // column-stationary DFMA kernels: rows are always written; minima to the local matrix only
    if (!A.out || S.npeers > 0) {
        bez_set_error("degree %d / elevation %d / dim %d is outside the tensor-path kernels: rows cannot be "
                      "skipped and the fused peer stores are not available", plan->n, plan->elev, plan->dim);
        return BEZ_EUNSUPPORTED;
    }
    BEZ_REQUIRE(!derived || S.itemmin, "the active mask / list of this shape needs the itemmin matrix");
    BEZ_REQUIRE(S.min_pitch == 0, "this shape needs min_pitch == nitems");
    int rc = mode == PAIR ? dispatch_degree<PAIR>(plan, A, st) : dispatch_degree<SPEED>(plan, A, st);
    if (rc != BEZ_OK || !derived) return rc;
    const long long total = A.nitems * (long long)A.B;
    emit_from_minima_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(S, total);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

extern "C" int bez_pair_sepsq_elev_ex(const bez_plan *plan, const double *d_cpts, int B, int N,
                                      int64_t pair_begin, int64_t npairs, double maxSep2,
                                      double *d_out, const bez_reduce_opts *opts, void *stream) {
    BEZ_REQUIRE(plan && d_cpts, "NULL argument");
    BEZ_REQUIRE(B >= 0 && N >= 0, "negative size");
    const long long P = (long long)N * (N - 1) / 2;
    BEZ_REQUIRE(pair_begin >= 0 && npairs >= 0 && pair_begin + npairs <= (P > 0 ? P : 0),
                "pair range outside the N(N-1)/2 list");
    if (B == 0 || npairs == 0) return BEZ_OK;
    BEZ_ON_DEVICE(plan->device);
    SqElevArgs A;
    A.cpts = d_cpts; A.tf = nullptr; A.PQ = plan->d_PQ; A.out = d_out;
    A.item_begin = pair_begin; A.nitems = npairs; A.B = B; A.N = N;
    A.L = plan->L; A.Lh = plan->Lh; A.LhPad = plan->LhPad;
    A.alpha = 1.0; A.beta = -maxSep2;
    return run_sq_elev(plan, A, PAIR, opts, nullptr, (cudaStream_t)stream);
}

extern "C" int bez_pair_sepsq_elev(const bez_plan *plan, const double *d_cpts, int B, int N,
                                   int64_t pair_begin, int64_t npairs, double maxSep2,
                                   double *d_out, double *d_pairmin, void *stream) {
    BEZ_REQUIRE(d_out, "NULL argument");
    bez_reduce_opts o;
    memset(&o, 0, sizeof(o));
    o.itemmin = d_pairmin;
    return bez_pair_sepsq_elev_ex(plan, d_cpts, B, N, pair_begin, npairs, maxSep2, d_out, &o, stream);
}

extern "C" int bez_pair_sepsq_elev_p2p(const bez_plan *plan, const double *d_cpts, int B, int N,
                                       int64_t pair_begin, int64_t npairs, double maxSep2,
                                       double *d_out, double *d_pairmin,
                                       const uint64_t *h_peer_min, int npeers, void *stream) {
    BEZ_REQUIRE(d_out && d_pairmin, "NULL argument");
    bez_reduce_opts o;
    memset(&o, 0, sizeof(o));
    o.itemmin = d_pairmin;
    o.peer_min = h_peer_min;
    o.npeers = npeers;
    return bez_pair_sepsq_elev_ex(plan, d_cpts, B, N, pair_begin, npairs, maxSep2, d_out, &o, stream);
}

extern "C" int bez_speed_sq_elev_ex(const bez_plan *plan, const double *d_cpts, const double *d_tf,
                                    int B, int N, int veh_begin, int nveh, double alpha, double beta,
                                    double *d_out, const bez_reduce_opts *opts, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_tf && d_out, "NULL argument");
    BEZ_REQUIRE(B >= 0 && N >= 0 && veh_begin >= 0 && nveh >= 0 && veh_begin + nveh <= N,
                "vehicle range outside [0, N)");
    if (B == 0 || nveh == 0) return BEZ_OK;
    BEZ_ON_DEVICE(plan->device);
    SqElevArgs A;
    A.cpts = d_cpts; A.tf = d_tf; A.PQ = plan->d_PQ; A.out = d_out;
    A.item_begin = veh_begin; A.nitems = nveh; A.B = B; A.N = N;
    A.L = plan->L; A.Lh = plan->Lh; A.LhPad = plan->LhPad;
    A.alpha = alpha; A.beta = beta;
    return run_sq_elev(plan, A, SPEED, opts, nullptr, (cudaStream_t)stream);
}

extern "C" int bez_speed_sq_elev(const bez_plan *plan, const double *d_cpts, const double *d_tf,
                                 int B, int N, int veh_begin, int nveh,
                                 double alpha, double beta, double *d_out, void *stream) {
    return bez_speed_sq_elev_ex(plan, d_cpts, d_tf, B, N, veh_begin, nveh, alpha, beta, d_out, nullptr, stream);
}

extern "C" int bez_assemble_cpts(const bez_plan *plan, const double *d_x, int B, int nvar,
                                 int numVeh, int nObs, int fixed_ends, int dubins, int timeopt,
                                 double tf_fixed,
                                 const double *d_init, const double *d_final,
                                 const double *d_ispeed, const double *d_fspeed,
                                 const double *d_icos, const double *d_isin,
                                 const double *d_fcos, const double *d_fsin,
                                 const double *d_obst,
                                 double *d_cpts, double *d_tf, void *stream) {
    return bez_assemble_cpts_sets(plan, d_x, B, nvar, numVeh, nObs, fixed_ends, dubins, timeopt, tf_fixed,
                                  d_init, d_final, d_ispeed, d_fspeed, d_icos, d_isin, d_fcos, d_fsin,
                                  d_obst, 0, d_cpts, d_tf, stream);
}

extern "C" int bez_assemble_cpts_sets(const bez_plan *plan, const double *d_x, int B, int nvar,
                                      int numVeh, int nObs, int fixed_ends, int dubins, int timeopt,
                                      double tf_fixed,
                                      const double *d_init, const double *d_final,
                                      const double *d_ispeed, const double *d_fspeed,
                                      const double *d_icos, const double *d_isin,
                                      const double *d_fcos, const double *d_fsin,
                                      const double *d_obst, int evals_per_obst_set,
                                      double *d_cpts, double *d_tf, void *stream) {
    BEZ_REQUIRE(evals_per_obst_set >= 0, "evals_per_obst_set is negative");
    BEZ_REQUIRE(plan && d_cpts && d_tf, "NULL argument");
    BEZ_REQUIRE(d_x || nvar == 0, "x is NULL");
    BEZ_REQUIRE(B >= 0 && numVeh >= 1 && nObs >= 0, "bad sizes");
    BEZ_REQUIRE(!fixed_ends || (d_init && d_final), "fixed_ends needs init/final points");
    BEZ_REQUIRE(!dubins || (fixed_ends && plan->dim == 2 && d_ispeed && d_fspeed && d_icos &&
                            d_isin && d_fcos && d_fsin),
                "dubins needs dim == 2, fixed ends and speed/angle tables");
    BEZ_REQUIRE(nObs == 0 || d_obst, "obstacles are NULL");
    const int offset = (fixed_ends ? 1 : 0) + (dubins ? 1 : 0);
    const int ncols = plan->n + 1 - 2 * offset;
    BEZ_REQUIRE(ncols >= 0, "degree too small for the fixed columns");
    BEZ_REQUIRE(nvar == numVeh * plan->dim * ncols + (timeopt ? 1 : 0), "nvar does not match the model");
    if (B == 0) return BEZ_OK;
    BEZ_ON_DEVICE(plan->device);
    AssembleArgs A;
    A.x = d_x; A.B = B; A.nvar = nvar; A.numVeh = numVeh; A.nObs = nObs; A.n = plan->n;
    A.dim = plan->dim; A.fixed_ends = fixed_ends; A.dubins = dubins; A.timeopt = timeopt;
    A.tf_fixed = tf_fixed; A.init = d_init; A.fin = d_final; A.ispeed = d_ispeed;
    A.fspeed = d_fspeed; A.icos = d_icos; A.isin = d_isin; A.fcos = d_fcos; A.fsin = d_fsin;
    A.obst = d_obst; A.cpts = d_cpts; A.tf = d_tf; A.evals_per_obst_set = evals_per_obst_set;
    const int S = (plan->dim * (plan->n + 1) + 1) / 2 * 2;
    const long long total = (long long)B * (numVeh + nObs) * S;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (total + 148LL * 16 * 256 < 0xffffffffLL)
        assemble_cpts_kernel<unsigned><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A);
    else
        assemble_cpts_kernel<long long><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}
