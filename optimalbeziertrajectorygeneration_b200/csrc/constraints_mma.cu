// Tensor-path instantiations of the fused constraint kernels (see sq_elev_mma.cuh).
// A separate translation unit so the two kernel families compile in parallel.
#include <stdlib.h>

#include "sq_elev_mma.cuh"
#include "sq_elev_stage1.cuh"

namespace {
using namespace bezcore;

// ---------------------------------------------------------------------------
// Tensor-path variant (DMMA stage 2 + TMA bulk-store epilogue, sq_elev_mma.cuh) for
// the shapes of the headline workload: 33..64 column pairs (65 <= L <= 128), n <= 15.
// Same tiling (one warp = 32 items, no block-wide barriers in the main loop), same
// stage 1; per warp 8.4 KB of staged rows + two [8][L] output staging buffers.
template <int N_, int DIM, int MODE, int MINMODE>
__global__ void __launch_bounds__(kThreads, 2)
sq_elev_mma_kernel(const SqElevArgs A, const ProdWeights<N_> PW, const DiffWeights<N_> DW) {
    using namespace bezmma;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t per_warp = (size_t)kRowsDoubles + 16 * (size_t)A.L;
    double *rows = smem + warp * per_warp;
    double *obuf = rows + kRowsDoubles;
    for (int i = lane; i < kRowsDoubles; i += 32) rows[i] = 0.0;     // padding slots must be 0
    const unsigned obuf_s = (unsigned)__cvta_generic_to_shared(obuf);
    const LaneGeom G = lane_geom(lane, A.Lh);
    BFrags<N_> Bf;
    load_bfrags<N_>(Bf, A.PQ, A.Lh, A.LhPad, lane);
    __syncwarp();

    // Tiles run over the flattened item list (evaluation-major, [B][nitems]): the output rows
    // of consecutive items are contiguous across evaluation points, so a tile may straddle
    // them (C5 has 16 pair rows / 1 speed row per evaluation).
    const long long total = A.nitems * (long long)A.B;
    const long long nwt = (total + 31) >> 5;
    const long long gwarp = (long long)blockIdx.x * kWarps + warp;
    const long long nwarps = (long long)gridDim.x * kWarps;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(A.out) & 15u) == 0;

    for (long long wt = gwarp; wt < nwt; wt += nwarps) {
        const long long g0 = wt << 5;                          // first flattened item of the tile
        const int cnt = (int)((total - g0) < 32 ? (total - g0) : 32);
        {
            // lanes past the end recompute the last item so every staged row is finite
            const long long gi = g0 + (lane < cnt ? lane : cnt - 1);
            const int b = (int)(gi / A.nitems);
            double s[2 * N_ + 1];
            stage1_coeffs<N_, DIM, MODE>(A, PW, DW, b, gi - (long long)b * A.nitems, 0, s);
            double *row = rows + lane * kRowStride;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                row[slot_e(j)] = s[j] + s[2 * N_ - j];
                row[slot_o(j)] = s[j] - s[2 * N_ - j];
            }
            row[slot_e(N_)] = s[N_];
        }
        __syncwarp();
        mma_tile<N_, MINMODE>(rows, obuf, obuf_s, Bf, G, A.out + (size_t)g0 * A.L, MINMODE ? A.itemmin + g0 : nullptr,
                               cnt, A.L, A.Lh, A.beta, lane, base_aligned, A.peer_min, MINMODE ? A.npeers : 0, g0);
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();      // staging buffers must outlive the last bulk reads
}

template <int N_, int DIM, int MODE, int MINMODE>
int launch_sq_elev_mma(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    ProdWeights<N_> PW;
    DiffWeights<N_> DW;
    const double scale = A.alpha * (0.5 * (double)DIM);       // Q1: dim/2 and the sign of alpha, folded
    for (int i = 0; i <= N_; ++i)                             // into the product weights
        for (int j = i; j <= N_; ++j) {
            double w = plan->h_W[i * (N_ + 1) + j] * scale;
            PW.w[widx<N_>(i, j)] = (i == j) ? w : 2.0 * w;
        }
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t shmem = (size_t)kWarps * (bezmma::kRowsDoubles + 16 * (size_t)A.L) * sizeof(double);
    auto kern = sq_elev_mma_kernel<N_, DIM, MODE, MINMODE>;
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

template <int N_, int MODE>
int mma_dispatch_dim(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    const bool wm = A.itemmin != nullptr;
    switch (plan->dim) {
#define CASE(d_) case d_: return wm ? launch_sq_elev_mma<N_, d_, MODE, 1>(plan, A, st) \
                                     : launch_sq_elev_mma<N_, d_, MODE, 0>(plan, A, st);
        CASE(1) CASE(2) CASE(3)
#undef CASE
    }
    return BEZ_EUNSUPPORTED;
}

template <int MODE>
int mma_dispatch_degree(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    switch (plan->n) {
#define CASE(n_) case n_: return mma_dispatch_dim<n_, MODE>(plan, A, st);
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

}  // namespace

namespace bezcore {

// BEZGPU_FORCE_DFMA=1 keeps every shape on the column-stationary DFMA kernel (A/B runs and
// tests/test_gpu_constraints.py::test_tensor_path_matches_dfma_path); read on every call.
static bool bez_force_dfma() {
    const char *e = getenv("BEZGPU_FORCE_DFMA");
    return e && e[0] == '1';
}

bool bez_sq_elev_mma_supported(const bez_plan *plan) {
#ifdef BEZ_ONLY_N
    if (plan->n != BEZ_ONLY_N) return false;
#endif
    return plan->n <= 15 && plan->Lh > 32 && plan->Lh <= 64 && !bez_force_dfma();
}

int bez_sq_elev_mma(const bez_plan *plan, const SqElevArgs &A, int mode, cudaStream_t st) {
    return mode == PAIR ? mma_dispatch_degree<PAIR>(plan, A, st) : mma_dispatch_degree<SPEED>(plan, A, st);
}

}  // namespace bezcore
