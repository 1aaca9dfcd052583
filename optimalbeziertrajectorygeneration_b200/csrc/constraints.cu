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
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 256;   // items (pairs / vehicles) per tile == threads per block
constexpr int kChunk = 4;    // items processed together in stage 2 (ILP)

enum Mode { PAIR = 0, SPEED = 1 };

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

struct SqElevArgs {
    const double *cpts;     // [B][dim][n+1][N]
    const double *tf;       // [B] (SPEED)
    const double *PQ;       // [2n+1][LhPad]
    double *out;            // [B][nitems][L]
    double *itemmin;        // [B][nitems] or null
    long long item_begin;   // first pair / vehicle handled
    long long nitems;       // pairs / vehicles per evaluation point
    int B, N, L, Lh, LhPad;
    double alpha, beta;     // out = alpha * value + beta
};

template <int N_, int DIM, int MODE>
__global__ void __launch_bounds__(kThreads, 2)
sq_elev_kernel(const SqElevArgs A, const ProdWeights<N_> PW, const DiffWeights<N_> DW) {
    constexpr int NC = N_ + 1;
    constexpr int ROW = 2 * N_ + 2;          // doubles per shared row (16 B aligned)
    extern __shared__ __align__(16) double smem[];   // [kTile][ROW]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const long long tiles_per_eval = (A.nitems + kTile - 1) / kTile;
    const long long ntiles = tiles_per_eval * A.B;
    const int CG = A.LhPad >> 5;             // column groups of 32
    const int M = A.L - 1;
    const size_t bstride = (size_t)DIM * NC * A.N;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_eval);
        const long long t0 = (tile - (long long)b * tiles_per_eval) * kTile;  // first item of tile
        const int cnt = (int)((A.nitems - t0) < kTile ? (A.nitems - t0) : kTile);

        // ------------------------------ stage 1 ------------------------------
        if (tid < cnt) {
            const double *base = A.cpts + (size_t)b * bstride;
            double a[DIM][NC];
            if (MODE == PAIR) {
                int vi, vj;
                bez_pair_decode(A.item_begin + t0 + tid, A.N, vi, vj);
#pragma unroll
                for (int d = 0; d < DIM; ++d)
#pragma unroll
                    for (int k = 0; k < NC; ++k) {
                        const double *row = base + (size_t)(d * NC + k) * A.N;
                        a[d][k] = __ldg(row + vi) - __ldg(row + vj);   // Bezier.sub
                    }
            } else {
                const int v = (int)(A.item_begin + t0 + tid);
                const double val = (double)N_ / __ldg(A.tf + b);       // diffMatrix: n/tf
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    double pt[NC], dd[NC];
#pragma unroll
                    for (int k = 0; k < NC; ++k)
                        pt[k] = __ldg(base + (size_t)(d * NC + k) * A.N + v);
#pragma unroll
                    for (int k = 0; k < N_; ++k)                       // np.dot(cpts, Dm)
                        dd[k] = pt[k] * (-val) + pt[k + 1] * val;
                    dd[N_] = 0.0;
#pragma unroll
                    for (int k = 0; k < NC; ++k) {                     // .elev(1) back to degree n
                        double q = dd[k] * DW.lo[k];
                        if (k > 0) q = dd[k - 1] * DW.hi[k] + q;
                        a[d][k] = q;
                    }
                }
            }
            double s[2 * N_ + 1];
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
            const double scale = 0.5 * (double)DIM;                    // Q1: dim/2
            double *row = smem + (size_t)tid * ROW;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                const double lo = s[j] * scale, hi = s[2 * N_ - j] * scale;
                *reinterpret_cast<double2 *>(row + 2 * j) = make_double2(lo + hi, lo - hi);
            }
            *reinterpret_cast<double2 *>(row + 2 * N_) = make_double2(s[N_] * scale, 0.0);
        }
        __syncthreads();

        // ------------------------------ stage 2 ------------------------------
        {
            const int nslots = (CG <= kWarps) ? (kWarps / CG) : 1;
            const int slot = (CG <= kWarps) ? (warp / CG) : 0;
            if (slot < nslots) {
                for (int cg = (CG <= kWarps) ? (warp % CG) : warp; cg < CG; cg += kWarps) {
                    const int col = cg * 32 + lane;
                    const bool live = col < A.Lh;
                    const int mcol = M - col;
                    double P[NC], Q[N_ > 0 ? N_ : 1];
#pragma unroll
                    for (int j = 0; j < NC; ++j) P[j] = __ldg(A.PQ + (size_t)j * A.LhPad + col);
#pragma unroll
                    for (int j = 0; j < N_; ++j) Q[j] = __ldg(A.PQ + (size_t)(NC + j) * A.LhPad + col);

                    double *outb = A.out + ((size_t)b * A.nitems + t0) * A.L;
                    for (int p0 = slot * kChunk; p0 < cnt; p0 += nslots * kChunk) {
                        double se[kChunk], so[kChunk];
#pragma unroll
                        for (int u = 0; u < kChunk; ++u) {
                            const int p = (p0 + u < cnt) ? (p0 + u) : (cnt - 1);
                            const double *row = smem + (size_t)p * ROW;
                            double e = 0.0, o = 0.0;
#pragma unroll
                            for (int j = 0; j < N_; ++j) {
                                const double2 eo = *reinterpret_cast<const double2 *>(row + 2 * j);
                                e = fma(eo.x, P[j], e);
                                o = fma(eo.y, Q[j], o);
                            }
                            e = fma(row[2 * N_], P[N_], e);
                            se[u] = e;
                            so[u] = o;
                        }
#pragma unroll
                        for (int u = 0; u < kChunk; ++u) {
                            if (live && p0 + u < cnt) {
                                double *o = outb + (size_t)(p0 + u) * A.L;
                                __stcs(o + col, fma(A.alpha, se[u] + so[u], A.beta));
                                if (mcol != col) __stcs(o + mcol, fma(A.alpha, se[u] - so[u], A.beta));
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Min over the L outputs of each item (active-pair flag source).  Separate,
// bandwidth-trivial pass used only when the caller asks for it and the fused
// epilogue is not available.
__global__ void item_min_kernel(const double *__restrict__ vals, long long nrows, int L,
                                double *__restrict__ mins) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const double *r = vals + (size_t)warp * L;
    double m = INFINITY;
    for (int i = lane; i < L; i += 32) m = fmin(m, r[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) mins[warp] = m;
}

// ---------------------------------------------------------------------------
// reshapeVector (optimization.py:242-285) + obstacle rows (optimization.py:86-94)
struct AssembleArgs {
    const double *x;
    int B, nvar, numVeh, nObs, n, dim, fixed_ends, dubins, timeopt;
    double tf_fixed;
    const double *init, *fin, *ispeed, *fspeed, *icos, *isin, *fcos, *fsin, *obst;
    double *cpts, *tf;
};

__global__ void assemble_cpts_kernel(const AssembleArgs A) {
    const int NC = A.n + 1;
    const int N = A.numVeh + A.nObs;
    const long long total = (long long)A.B * A.dim * NC * N;
    const int offset = (A.fixed_ends ? 1 : 0) + (A.dubins ? 1 : 0);
    const int ncols = NC - 2 * offset;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(idx % N);
        long long r = idx / N;
        const int k = (int)(r % NC);
        r /= NC;
        const int d = (int)(r % A.dim);
        const int b = (int)(r / A.dim);
        const double *x = A.x + (size_t)b * A.nvar;
        const double tf = A.timeopt ? x[A.nvar - 1] : A.tf_fixed;
        double val;
        if (v >= A.numVeh) {
            val = A.obst[(size_t)(v - A.numVeh) * A.dim + d];            // constant curve (Q8)
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
        if (v == 0 && k == 0 && d == 0) A.tf[b] = tf;
    }
}

// ---------------------------------------------------------------------------
template <int N_, int DIM, int MODE>
int launch_sq_elev(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    ProdWeights<N_> PW;
    DiffWeights<N_> DW;
    for (int i = 0; i <= N_; ++i)
        for (int j = i; j <= N_; ++j) {
            double w = plan->h_W[i * (N_ + 1) + j];
            PW.w[widx<N_>(i, j)] = (i == j) ? w : 2.0 * w;    // W is symmetric; G[i][j]==G[j][i]
        }
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }

    const size_t shmem = (size_t)kTile * (2 * N_ + 2) * sizeof(double);
    static bool attr_done = false;    // per template instantiation
    if (!attr_done) {
        BEZ_CUDA(cudaFuncSetAttribute(sq_elev_kernel<N_, DIM, MODE>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
        attr_done = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = ((A.nitems + kTile - 1) / kTile) * A.B;
    long long grid = (long long)sms * 2;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) return BEZ_OK;
    sq_elev_kernel<N_, DIM, MODE><<<(unsigned)grid, kThreads, shmem, st>>>(A, PW, DW);
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
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    }
    bez_set_error("degree %d has no fused kernel instantiation (1..16 supported)", plan->n);
    return BEZ_EUNSUPPORTED;
}

int run_item_min(const SqElevArgs &A, cudaStream_t st) {
    if (!A.itemmin) return BEZ_OK;
    const long long rows = (long long)A.B * A.nitems;
    const int threads = 256;
    const long long blocks = (rows * 32 + threads - 1) / threads;
    item_min_kernel<<<(unsigned)blocks, threads, 0, st>>>(A.out, rows, A.L, A.itemmin);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

}  // namespace

extern "C" int bez_pair_sepsq_elev(const bez_plan *plan, const double *d_cpts, int B, int N,
                                   int64_t pair_begin, int64_t npairs, double maxSep2,
                                   double *d_out, double *d_pairmin, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_out, "NULL argument");
    BEZ_REQUIRE(B >= 0 && N >= 0, "negative size");
    const long long P = (long long)N * (N - 1) / 2;
    BEZ_REQUIRE(pair_begin >= 0 && npairs >= 0 && pair_begin + npairs <= (P > 0 ? P : 0),
                "pair range outside the N(N-1)/2 list");
    if (B == 0 || npairs == 0) return BEZ_OK;
    BEZ_CUDA(cudaSetDevice(plan->device));
    SqElevArgs A;
    A.cpts = d_cpts; A.tf = nullptr; A.PQ = plan->d_PQ; A.out = d_out; A.itemmin = d_pairmin;
    A.item_begin = pair_begin; A.nitems = npairs; A.B = B; A.N = N;
    A.L = plan->L; A.Lh = plan->Lh; A.LhPad = plan->LhPad;
    A.alpha = 1.0; A.beta = -maxSep2;
    int rc = dispatch_degree<PAIR>(plan, A, (cudaStream_t)stream);
    if (rc != BEZ_OK) return rc;
    return run_item_min(A, (cudaStream_t)stream);
}

extern "C" int bez_speed_sq_elev(const bez_plan *plan, const double *d_cpts, const double *d_tf,
                                 int B, int N, int veh_begin, int nveh,
                                 double alpha, double beta, double *d_out, void *stream) {
    BEZ_REQUIRE(plan && d_cpts && d_tf && d_out, "NULL argument");
    BEZ_REQUIRE(B >= 0 && N >= 0 && veh_begin >= 0 && nveh >= 0 && veh_begin + nveh <= N,
                "vehicle range outside [0, N)");
    if (B == 0 || nveh == 0) return BEZ_OK;
    BEZ_CUDA(cudaSetDevice(plan->device));
    SqElevArgs A;
    A.cpts = d_cpts; A.tf = d_tf; A.PQ = plan->d_PQ; A.out = d_out; A.itemmin = nullptr;
    A.item_begin = veh_begin; A.nitems = nveh; A.B = B; A.N = N;
    A.L = plan->L; A.Lh = plan->Lh; A.LhPad = plan->LhPad;
    A.alpha = alpha; A.beta = beta;
    return dispatch_degree<SPEED>(plan, A, (cudaStream_t)stream);
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
    BEZ_CUDA(cudaSetDevice(plan->device));
    AssembleArgs A;
    A.x = d_x; A.B = B; A.nvar = nvar; A.numVeh = numVeh; A.nObs = nObs; A.n = plan->n;
    A.dim = plan->dim; A.fixed_ends = fixed_ends; A.dubins = dubins; A.timeopt = timeopt;
    A.tf_fixed = tf_fixed; A.init = d_init; A.fin = d_final; A.ispeed = d_ispeed;
    A.fspeed = d_fspeed; A.icos = d_icos; A.isin = d_isin; A.fcos = d_fcos; A.fsin = d_fsin;
    A.obst = d_obst; A.cpts = d_cpts; A.tf = d_tf;
    const long long total = (long long)B * plan->dim * (plan->n + 1) * (numVeh + nObs);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    assemble_cpts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}
