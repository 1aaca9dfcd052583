// Warp-specialised variant of the fused pair kernel (65 <= L <= 128): one persistent CTA of
// 12 warps per SM -- per scheduler one PRODUCER warp (stage 1: TMA row fetch, difference curve,
// Gram sums, folded rows) feeding two CONSUMER warps (stage 2: DMMA, epilogue, minimum, TMA
// stores; sq_elev_mma.cuh unchanged) through a ring of three row slots and full / empty
// mbarriers.  Registers are redistributed with setmaxnreg (producers 120, consumers 192).
//
// Why: with two in-order warps per scheduler the fp64 pipe idles whenever both are outside a
// DMMA burst (stage 1's load / Gram phase, the STS / fence / store stretch of an epilogue):
// 66-68 % busy at 0.36 ms (profiles/r02_ncu_pair_kernel_tma.txt) where the HBM bound needs 76 %.
// A third instruction stream per scheduler fills those holes; the register file does not hold
// a third full warp (255 registers: the elevation table alone is 96), but a stage-1-only warp
// needs no table and a stage-2-only warp no difference curve.
#pragma once
#include "sq_elev_stage1.cuh"

namespace bezws {
using namespace bezcore;
using namespace bezmma;

constexpr int kProducers = 4, kConsumers = 8;
constexpr int kWsThreads = 32 * (kProducers + kConsumers);
constexpr int kSlots = 3;                      // row slots per scheduler group
constexpr int kBarsPerGroup = 3 * kSlots;      // full[], empty[], rows[]

template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(R)); }
__device__ __forceinline__ void mbar_arrive(unsigned mbar_s) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar_s) : "memory");
}

template <int N_, int DIM> __host__ __device__ constexpr int slot_doubles() {
    constexpr int S_ = (DIM * (N_ + 1) + 1) / 2 * 2;
    return 32 * S_ > kRowsDoubles ? 32 * S_ : kRowsDoubles;
}

template <int N_, int DIM, int MINMODE, bool STORE>
__global__ void __launch_bounds__(kWsThreads, 1)
sq_elev_ws_kernel(const SqElevArgs A, const ProdWeights<N_> PW, const DiffWeights<N_> DW) {
    constexpr int NP = 4;
    constexpr int S_ = (DIM * (N_ + 1) + 1) / 2 * 2;
    constexpr int kSlotD = slot_doubles<N_, DIM>();
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for ptxas
    const int group = warp & 3;                                              // = scheduler of this warp
    // [4 groups][kSlots][kSlotD] | [8 consumers][2][8 L] | barriers
    double *slots = smem + (size_t)group * kSlots * kSlotD;
    double *stag0 = smem + (size_t)4 * kSlots * kSlotD;
    const size_t stag_per = STORE ? 16 * (size_t)A.L : 0;
    const unsigned bars_s = (unsigned)__cvta_generic_to_shared(stag0 + kConsumers * stag_per) + 8u * kBarsPerGroup * group;
    auto full_b = [&](int s) { return bars_s + 8u * s; };
    auto empty_b = [&](int s) { return bars_s + 8u * (kSlots + s); };
    auto rows_b = [&](int s) { return bars_s + 8u * (2 * kSlots + s); };
    if (warp < kProducers && lane == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(full_b(s), 32); mbar_init(empty_b(s), 32); mbar_init(rows_b(s), 1); }
        mbar_fence_init();
    }
    __syncthreads();

    // every group owns a contiguous run of warp tiles; tile k of the run goes to consumer k & 1
    const long long total = A.nitems * (long long)A.B;
    const long long nwt = (total + 31) >> 5;
    const long long ngroups = (long long)gridDim.x * 4, gidx = (long long)blockIdx.x * 4 + group;
    const long long t_begin = gidx * nwt / ngroups;
    const int n = (int)((gidx + 1) * nwt / ngroups - t_begin);

    if (warp < kProducers) {
        // ---------------------------------------------------------------- producer: stage 1
        reg_dec<120>();
        PairCursor cur;
        if (n > 0) {
            const long long f = (t_begin << 5) + lane;
            cur = pair_cursor_at(A, f < total ? f : total - 1);
        }
        for (int k = 0; k < n; ++k) {
            const int s = k % kSlots, use = k / kSlots;
            double *slot = slots + (size_t)s * kSlotD;
            const unsigned slot_s = (unsigned)__cvta_generic_to_shared(slot);
            if (use > 0) mbar_wait(empty_b(s), (unsigned)(use - 1) & 1u);    // consumers are done with the slot
            {   // TMA fetch of the partner rows: one bulk copy per run of lanes with consecutive rows
                const int pb = __shfl_up_sync(0xffffffffu, cur.b, 1), pi = __shfl_up_sync(0xffffffffu, cur.i, 1);
                const int pj = __shfl_up_sync(0xffffffffu, cur.j, 1);
                const bool start = lane == 0 || cur.b != pb || cur.i != pi || cur.j != pj + 1;
                const unsigned runs = __ballot_sync(0xffffffffu, start);
                if (lane == 0) mbar_arrive_expect_tx(rows_b(s), 32u * S_ * 8u);
                __syncwarp();
                if (start) {
                    const unsigned higher = lane == 31 ? 0u : (runs & (0xffffffffu << (lane + 1)));
                    const int end = higher ? __ffs(higher) - 1 : 32;
                    bulk_load(slot_s + (unsigned)lane * (S_ * 8u), A.cpts + ((size_t)cur.b * A.N + cur.j) * S_,
                              (unsigned)(end - lane) * (S_ * 8u), rows_b(s));
                }
            }
            // the cursor of the next tile, while the rows are on their way
            const long long g0 = (t_begin + k) << 5;
            PairCursor nxt = cur;
            if (k + 1 < n) {
                const long long fn = g0 + 32 + lane;
                nxt = fn < total ? pair_cursor_next(A, cur, g0 + lane) : pair_cursor_at(A, total - 1);
            }
            mbar_wait(rows_b(s), (unsigned)use & 1u);
            double sc[2 * N_ + 1];
            stage1_coeffs<N_, DIM, PAIR, true>(A, PW, DW, cur.b, cur.i, cur.j, sc, slot + lane * S_);
            __syncwarp();                                                   // every lane has its row
            double *row = slot + lane * kRowStride;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                row[slot_e(j)] = sc[j] + sc[2 * N_ - j];
                row[slot_o(j)] = sc[j] - sc[2 * N_ - j];
            }
            row[slot_e(N_)] = sc[N_];
#pragma unroll
            for (int j = N_ + 1; j < 4 * Geom<N_>::KE; ++j) row[slot_e(j)] = 0.0;
#pragma unroll
            for (int j = N_; j < 4 * Geom<N_>::KO; ++j) row[slot_o(j)] = 0.0;
            mbar_arrive(full_b(s));                                         // 32 arrivals (release)
            cur = nxt;
        }
    } else {
        // ---------------------------------------------------------------- consumer: stage 2
        reg_inc<192>();
        const int cidx = warp - kProducers, ci = cidx >> 2;                  // consumer 0 / 1 of the group
        double *obuf = stag0 + (size_t)cidx * stag_per;
        const unsigned obuf_s = (unsigned)__cvta_generic_to_shared(obuf);
        const bool base_aligned = (reinterpret_cast<uintptr_t>(A.out) & 15u) == 0;
        BFrags<N_, NP> Bf;
        load_bfrags<N_, NP>(Bf, A.PQ, A.L, A.LhPad, lane);
        for (int k = ci; k < n; k += 2) {
            const int s = k % kSlots, use = k / kSlots;
            const long long g0 = (t_begin + k) << 5;
            const int cnt = (int)((total - g0) < 32 ? (total - g0) : 32);
            mbar_wait(full_b(s), (unsigned)use & 1u);
            auto release = [&]() { mbar_arrive(empty_b(s)); };
            mma_tile<N_, NP, MINMODE, STORE, decltype(release)>(slots + (size_t)s * kSlotD, obuf, obuf_s, Bf,
                                                                STORE ? A.out + (size_t)g0 * A.L : nullptr, A.sinks, g0, cnt,
                                                                A.L, A.beta, lane, base_aligned, false, release);
            __syncwarp();
        }
        if (STORE && lane == 0) bulk_wait_all();      // staging buffers must outlive the last bulk reads
    }
}

template <int N_, int DIM> size_t ws_shmem_bytes(int L, bool store) {
    return ((size_t)4 * kSlots * slot_doubles<N_, DIM>() + (store ? (size_t)kConsumers * 16 * L : 0) + 4 * kBarsPerGroup) *
           sizeof(double);
}
// 227 KB of dynamic shared memory per CTA on sm_100 (C4: 102 KB of row slots + 121 KB of staging)
template <int N_, int DIM> bool ws_fits(int L, bool store) { return ws_shmem_bytes<N_, DIM>(L, store) <= 232448; }

template <int N_, int DIM, int MINMODE, bool STORE>
int launch_sq_elev_ws(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    ProdWeights<N_> PW;
    DiffWeights<N_> DW;
    const double scale = A.alpha * (0.5 * (double)DIM);
    for (int i = 0; i <= N_; ++i)
        for (int j = i; j <= N_; ++j) {
            double w = plan->h_W[i * (N_ + 1) + j] * scale;
            PW.w[widx<N_>(i, j)] = (i == j) ? w : 2.0 * w;
        }
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t shmem = ws_shmem_bytes<N_, DIM>(A.L, STORE);
    auto kern = sq_elev_ws_kernel<N_, DIM, MINMODE, STORE>;
    int sms = 148, per_sm = 1;
    if (int rc = bez_kernel_config((const void *)kern, kWsThreads, shmem, &sms, &per_sm)) return rc;
    const long long nwt = (A.nitems * (long long)A.B + 31) / 32;
    long long grid = sms;
    if (grid > (nwt + 7) / 8) grid = (nwt + 7) / 8;
    // One SM stays free.  The CTAs of this kernel fill an SM completely (223 KB, all registers), so every small
    // kernel queued behind or next to it -- the next step's counter reset / assemble / speed rows on the other
    // launch stream, the list compaction, the completion barrier of the fused all-gather -- would otherwise wait
    // for a whole persistent kernel to drain (tools/timeline_e2e.py: an 88 us bubble per two steps end to end).
    // 0.7 % of the grid; measured +0.5-1 % device-resident and +2 % end to end.  BEZGPU_MMA_FLAGS bit 16: full grid.
    if (grid == sms && grid > 1 && !(A.flags & kFlagFullGrid)) grid -= 1;
    if (grid < 1) return BEZ_OK;
    kern<<<(unsigned)grid, kWsThreads, shmem, st>>>(A, PW, DW);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

}  // namespace bezws
