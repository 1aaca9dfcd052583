// Tensor-path instantiations of the fused constraint kernels (see sq_elev_mma.cuh): kernel,
// launcher and the degree / dimension / shape dispatch.  Included by the small translation
// units constraints_mma_<mode>_<range>.cu, each of which instantiates one MODE for a range of
// degrees (BEZ_MMA_MODE, BEZ_MMA_NLO, BEZ_MMA_NHI) so that they compile in parallel.
#pragma once
#include <stdlib.h>

#include "sq_elev_stage1.cuh"
#include "sq_elev_mma_wide.cuh"
#include "sq_elev_ws.cuh"

namespace bezmma_inst {
using namespace bezcore;

// DMMA stage 2 + TMA bulk-store epilogue for every shape with L <= 128 and n <= 15:
// NP n-tile pairs cover 16 NP column-pair slots (NP = 4: 65 <= L <= 128, the headline shapes;
// NP = 2: 33 <= L <= 64; NP = 1: L <= 32).  One warp = 32 items, no block-wide barriers in the
// main loop; per warp 8.4 KB of staged rows + two [8][L] output staging buffers.
//
// TMAROWS (PAIR): the partner-vehicle rows of a tile (lane l: row j_l of evaluation b_l) are
// contiguous runs of the control-point array -- one run while the tile stays inside a row of
// the pair list -- so they are fetched with one `cp.async.bulk` (TMA) per run into the warp's
// row region, completion on a per-warp mbarrier.  The fetch of tile t + 1 is issued from inside
// the DMMA phase of tile t (mma_tile's rows_free hook: the region is shared with the staged
// (e, o) rows and is dead once the last A fragments are in registers) and lands behind the last
// m-tile.  Per-lane global loads of the same rows touch 32 different cache lines per LDG.128
// (1088 L1 wavefronts per tile, l1tex throughput 66 %) and left the warp 9 % of its time on the
// long scoreboard (profiles/r02_ncu_pair_kernel.txt); the i rows (one row for the whole tile
// in general) stay on the L1-resident broadcast loads.
template <int N_, int DIM> __host__ __device__ constexpr int region_doubles(bool tmarows) {
    constexpr int S_ = (DIM * (N_ + 1) + 1) / 2 * 2;
    return (tmarows && 32 * S_ > bezmma::kRowsDoubles) ? 32 * S_ : bezmma::kRowsDoubles;
}

template <int N_, int DIM, int MODE, int NP, int MINMODE, bool STORE, bool TMAROWS>
__global__ void __launch_bounds__(kThreads, 2)
sq_elev_mma_kernel(const SqElevArgs A, const ProdWeights<N_> PW, const DiffWeights<N_> DW) {
    using namespace bezmma;
    static_assert(!TMAROWS || MODE == PAIR, "the TMA row fetch is for the pair kernel");
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kRegion = region_doubles<N_, DIM>(TMAROWS);
    constexpr int S_ = (DIM * (N_ + 1) + 1) / 2 * 2;
    const size_t per_warp = (size_t)kRegion + (STORE ? 16 * (size_t)A.L : 0);
    double *rows = smem + warp * per_warp;
    double *obuf = rows + kRegion;
    const unsigned rows_s = (unsigned)__cvta_generic_to_shared(rows);
    const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(smem + kWarps * per_warp) + 8u * warp;
    if (TMAROWS) {
        if (lane == 0) { mbar_init(mbar_s, 1); mbar_fence_init(); }
    } else {
        for (int i = lane; i < kRowsDoubles; i += 32) rows[i] = 0.0;     // padding slots must be 0
    }
    const unsigned obuf_s = (unsigned)__cvta_generic_to_shared(obuf);
    BFrags<N_, NP> Bf;
    load_bfrags<N_, NP>(Bf, A.PQ, A.L, A.LhPad, lane);
    __syncwarp();

    // Tiles run over the flattened item list (evaluation-major, [B][nitems]): the output rows
    // of consecutive items are contiguous across evaluation points, so a tile may straddle
    // them (C5 has 16 pair rows / 1 speed row per evaluation).  Every warp owns a contiguous
    // run of tiles, so that the position in the pair list advances incrementally (PairCursor).
    const long long total = A.nitems * (long long)A.B;
    const long long nwt = (total + 31) >> 5;
    const long long gwarp = (long long)blockIdx.x * kWarps + warp;
    const long long nwarps = (long long)gridDim.x * kWarps;
    // contiguous runs pay off when an evaluation has many items (incremental decode); with few
    // items per evaluation (C5: 16 pair rows) every tile needs a full decode anyway and the
    // warp-strided order keeps all SMs writing one compact window of HBM (measured: 5 % faster)
    const bool strided = (A.flags & kFlagStridedTiles) != 0 || MODE != PAIR || A.nitems < 4096;
    long long wt = strided ? gwarp : gwarp * nwt / nwarps;
    const long long wt_end = strided ? nwt : (gwarp + 1) * nwt / nwarps;
    const long long wt_step = strided ? nwarps : 1;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(A.out) & 15u) == 0;
    const bool early_store = (A.flags & kFlagEarlyFence) != 0;

    PairCursor cur;
    bool have_next = false;
    unsigned phase = 0;
    // TMA fetch of the rows c.j of all lanes (warp converged): one bulk copy per run of lanes
    // whose rows are consecutive in memory; lane l's row lands at rows + l * S_
    auto fetch_rows = [&]() {
        if (!TMAROWS || !have_next) return;
        const int pb = __shfl_up_sync(0xffffffffu, cur.b, 1), pi = __shfl_up_sync(0xffffffffu, cur.i, 1);
        const int pj = __shfl_up_sync(0xffffffffu, cur.j, 1);
        const bool start = lane == 0 || cur.b != pb || cur.i != pi || cur.j != pj + 1;
        const unsigned runs = __ballot_sync(0xffffffffu, start);
        fence_async_smem();             // generic-proxy accesses of the region before the async-proxy writes
        if (lane == 0) mbar_arrive_expect_tx(mbar_s, 32u * S_ * 8u);
        __syncwarp();
        if (start) {
            const unsigned higher = lane == 31 ? 0u : (runs & (0xffffffffu << (lane + 1)));
            const int end = higher ? __ffs(higher) - 1 : 32;
            bulk_load(rows_s + (unsigned)lane * (S_ * 8u), A.cpts + ((size_t)cur.b * A.N + cur.j) * S_,
                      (unsigned)(end - lane) * (S_ * 8u), mbar_s);
        }
    };
    if (MODE == PAIR && wt < wt_end) {
        const long long f = (wt << 5) + lane;
        cur = pair_cursor_at(A, f < total ? f : total - 1);
        have_next = true;
        fetch_rows();
    }
    for (; wt < wt_end; wt += wt_step) {
        const long long g0 = wt << 5;                          // first flattened item of the tile
        const int cnt = (int)((total - g0) < 32 ? (total - g0) : 32);
        {
            double s[2 * N_ + 1];
            if (MODE == PAIR) {
                if (TMAROWS) { mbar_wait(mbar_s, phase); phase ^= 1u; }
                stage1_coeffs<N_, DIM, MODE, TMAROWS>(A, PW, DW, cur.b, cur.i, cur.j, s, rows + lane * S_);
                if (TMAROWS) __syncwarp();                     // every lane has its row: the region is rewritten below
                // this lane's item of the warp's next tile (lanes past the end of the list stay on
                // the last item, so every staged row is finite)
                const long long fn = g0 + (long long)wt_step * 32 + lane;
                have_next = wt + wt_step < wt_end;
                if (have_next) {
                    if (wt_step == 1 && fn < total) cur = pair_cursor_next(A, cur, g0 + lane);
                    else cur = pair_cursor_at(A, fn < total ? fn : total - 1);
                    if (A.flags & kFlagPrefetchL1) {
                        const char *rb = reinterpret_cast<const char *>(A.cpts + (size_t)cur.b * ((size_t)S_ * A.N));
                        const char *pj = rb + (size_t)cur.j * (S_ * 8);
                        const char *pi = rb + (size_t)cur.i * (S_ * 8);
#pragma unroll
                        for (int o = 0; o < S_ * 8; o += 128) {
                            asm volatile("prefetch.global.L1 [%0];" :: "l"(pj + o));
                            asm volatile("prefetch.global.L1 [%0];" :: "l"(pi + o));
                        }
                    }
                }
            } else {
                // lanes past the end recompute the last item so every staged row is finite
                const long long gi = g0 + (lane < cnt ? lane : cnt - 1);
                const int b = (int)(gi / A.nitems);
                stage1_coeffs<N_, DIM, MODE>(A, PW, DW, b, (int)(A.item_begin + gi - (long long)b * A.nitems), 0, s);
            }
            double *row = rows + lane * kRowStride;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                row[slot_e(j)] = s[j] + s[2 * N_ - j];
                row[slot_o(j)] = s[j] - s[2 * N_ - j];
            }
            row[slot_e(N_)] = s[N_];
            if (TMAROWS) {               // the fetched rows overwrite the padding slots of the k-steps
#pragma unroll
                for (int j = N_ + 1; j < 4 * Geom<N_>::KE; ++j) row[slot_e(j)] = 0.0;
#pragma unroll
                for (int j = N_; j < 4 * Geom<N_>::KO; ++j) row[slot_o(j)] = 0.0;
            }
        }
        __syncwarp();
        mma_tile<N_, NP, MINMODE, STORE, decltype(fetch_rows)>(rows, obuf, obuf_s, Bf, STORE ? A.out + (size_t)g0 * A.L : nullptr,
                                                               A.sinks, g0, cnt, A.L, A.beta, lane, base_aligned, early_store,
                                                               fetch_rows);
        __syncwarp();
    }
    if (STORE && lane == 0) bulk_wait_all();      // staging buffers must outlive the last bulk reads
}

template <int N_, int DIM, int MODE, int NP, int MINMODE, bool STORE, bool TMAROWS>
int launch_sq_elev_mma_rows(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    ProdWeights<N_> PW;
    DiffWeights<N_> DW;
    const double scale = A.alpha * (0.5 * (double)DIM);       // Q1: dim/2 and the sign of alpha, folded
    for (int i = 0; i <= N_; ++i)                             // into the product weights
        for (int j = i; j <= N_; ++j) {
            double w = plan->h_W[i * (N_ + 1) + j] * scale;
            PW.w[widx<N_>(i, j)] = (i == j) ? w : 2.0 * w;
        }
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t shmem = (size_t)kWarps * (region_doubles<N_, DIM>(TMAROWS) + (STORE ? 16 * (size_t)A.L : 0) + 1) * sizeof(double);
    auto kern = sq_elev_mma_kernel<N_, DIM, MODE, NP, MINMODE, STORE, TMAROWS>;
    int sms = 148, per_sm = 1;
    if (int rc = bez_kernel_config((const void *)kern, kThreads, shmem, &sms, &per_sm)) return rc;
    const long long nwt = (A.nitems * (long long)A.B + 31) / 32;
    long long grid = (long long)sms * per_sm;
    const long long need = (nwt + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    // Fused all-gather: the caller's completion barrier (a tiny kernel on a high-priority side
    // stream, sharding.PeerMinima) has to run *next to* the following persistent launch, so one
    // CTA slot of the machine is left free for it (0.3 % of the grid) -- otherwise it only gets
    // an SM when a whole persistent kernel has drained and ends up between two of them.
    if (A.sinks.npeers > 0 && grid == (long long)sms * per_sm && grid > 1 && !(A.flags & kFlagFullGridWithPeers)) grid -= 1;
    if (grid < 1) return BEZ_OK;
    kern<<<(unsigned)grid, kThreads, shmem, st>>>(A, PW, DW);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

// The pair kernel fetches its partner rows with TMA; the per-lane global loads stay available
// for A/B runs in development builds (BEZGPU_MMA_FLAGS bit 4).
template <int N_, int DIM, int MODE, int NP, int MINMODE, bool STORE>
int launch_sq_elev_mma(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    if constexpr (MODE == PAIR) {
#ifdef BEZ_ONLY_N
        if (A.flags & kFlagRowsByLdg) return launch_sq_elev_mma_rows<N_, DIM, MODE, NP, MINMODE, STORE, false>(plan, A, st);
#endif
        return launch_sq_elev_mma_rows<N_, DIM, MODE, NP, MINMODE, STORE, true>(plan, A, st);
    } else {
        return launch_sq_elev_mma_rows<N_, DIM, MODE, NP, MINMODE, STORE, false>(plan, A, st);
    }
}

// rows + minima / rows only / minima only (the combination "neither" is rejected by the caller)
template <int N_, int DIM, int MODE, int NP>
int mma_dispatch_variant(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    const bezmma::MinSinks &S = A.sinks;
    const bool wm = S.itemmin || S.mask || S.list_count || S.npeers > 0;
    // the headline shapes (pair rows, 65 <= L <= 128) run warp-specialised whenever the three row
    // slots per scheduler and the consumers' staging buffers fit the SM's shared memory
    if constexpr (NP == 4 && MODE == PAIR) {
        if (!(A.flags & kFlagNoWarpSpecialisation) && bezws::ws_fits<N_, DIM>(A.L, A.out != nullptr)) {
            if (A.out == nullptr) return bezws::launch_sq_elev_ws<N_, DIM, 1, false>(plan, A, st);
            return wm ? bezws::launch_sq_elev_ws<N_, DIM, 1, true>(plan, A, st)
                      : bezws::launch_sq_elev_ws<N_, DIM, 0, true>(plan, A, st);
        }
    }
    if (A.out == nullptr) {
        if (MODE == PAIR) return launch_sq_elev_mma<N_, DIM, MODE, NP, 1, false>(plan, A, st);
        bez_set_error("rows may only be skipped for the pair kernel");
        return BEZ_EUNSUPPORTED;
    }
    return wm ? launch_sq_elev_mma<N_, DIM, MODE, NP, 1, true>(plan, A, st)
              : launch_sq_elev_mma<N_, DIM, MODE, NP, 0, true>(plan, A, st);
}

template <int N_, int DIM, int MODE>
int mma_dispatch_shape(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    if (plan->Lh > 64) {                                  // L > 128: column-tiled variant (rows always written)
        if (A.out == nullptr) {
            bez_set_error("L = %d > 128: the rows cannot be skipped", plan->L);
            return BEZ_EUNSUPPORTED;
        }
        return bezwide::launch_sq_elev_mma_wide<N_, DIM, MODE, 1>(plan, A, st);
    }
    if (plan->Lh <= 16) return mma_dispatch_variant<N_, DIM, MODE, 1>(plan, A, st);
    if (plan->Lh <= 32) return mma_dispatch_variant<N_, DIM, MODE, 2>(plan, A, st);
    return mma_dispatch_variant<N_, DIM, MODE, 4>(plan, A, st);
}

template <int N_, int MODE>
int mma_dispatch_dim(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    switch (plan->dim) {
        case 2: return mma_dispatch_shape<N_, 2, MODE>(plan, A, st);
        case 3: return mma_dispatch_shape<N_, 3, MODE>(plan, A, st);
    }
    return BEZ_EUNSUPPORTED;
}

template <int MODE, int NLO, int NHI>
int mma_dispatch_degree(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    switch (plan->n) {
#define CASE(n_) case n_: if constexpr (n_ >= NLO && n_ <= NHI) return mma_dispatch_dim<n_, MODE>(plan, A, st); break;
#ifdef BEZ_ONLY_N
        CASE(BEZ_ONLY_N)
#else
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15)
#endif
#undef CASE
    }
    return BEZ_EUNSUPPORTED;
}

}  // namespace bezmma_inst

#ifdef BEZ_MMA_FN
namespace bezcore {
int BEZ_MMA_FN(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    return bezmma_inst::mma_dispatch_degree<BEZ_MMA_MODE, BEZ_MMA_NLO, BEZ_MMA_NHI>(plan, A, st);
}
}  // namespace bezcore
#endif
