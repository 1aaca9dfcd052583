// Third generation of the fused "square -> fold -> elevate" kernel for the headline shapes
// (65 <= L <= 128): TEAMS of two warps share the elevation table.
//
// Why: in sq_elev_mma_kernel every warp keeps the whole folded table as B fragments (96
// registers), which caps the SM at 8 warps (255 registers each).  ptxas issues every DMMA with
// a 16-cycle issue stall (the SMSP's fp64 pipe takes one DMMA per 16 cycles), so a warp cannot
// overlap its own epilogue / stage 1 with its DMMAs; with two warps per scheduler the pipe idles
// whenever both are in a non-DMMA phase: measured 0.354 ms of pure issue time per 4 evaluations
// without any HBM traffic (profiles/r02_ablation_pair_kernel.txt), above the 0.316 ms HBM floor.
// Here a team of two warps works on 64 items: warp w computes stage 1 for items 32 w .. 32 w + 31
// and then, for ALL eight m-tiles, the DMMAs of its half of the column slots (n-tile pairs 2 w,
// 2 w + 1: 48 registers of B fragments).  Both warps write their columns of an m-tile into one
// shared [8][L] staging block in its final HBM layout; after a team barrier one lane issues the
// TMA bulk store.  One CTA = one team (64 threads, __syncthreads() = the team barrier), six CTAs
// = 12 warps per SM at <= 168 registers: three warps per scheduler instead of two.
#pragma once
#include "sq_elev_stage1.cuh"

namespace bezteam {
using namespace bezcore;
using namespace bezmma;

constexpr int kTeamThreads = 64;
constexpr int kTeamItems = 64;
constexpr int kTeamRows = kTeamItems * kRowStride;       // doubles of staged (e,o) rows per team

// cursor of flattened item f + 64 given the cursor of f (f + 64 < total)
__device__ __forceinline__ PairCursor pair_cursor_next64(const SqElevArgs &A, const PairCursor &c, long long f) {
    if (c.left <= 64) return pair_cursor_at(A, f + 64);
    PairCursor d = c;
    d.left -= 64;
    int q = c.j + 64;
    while (q > A.N - 1) { q -= (A.N - 2 - d.i); ++d.i; }
    d.j = q;
    return d;
}

template <int N_, int DIM, int MODE, int MINMODE, bool STORE>
__global__ void __launch_bounds__(kTeamThreads, 6)
sq_elev_team_kernel(const SqElevArgs A, const ProdWeights<N_> PW, const DiffWeights<N_> DW) {
    constexpr int KE = Geom<N_>::KE, KO = Geom<N_>::KO;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int L = A.L, M = L - 1;
    double *rows = smem;                                   // [64][kRowStride]
    double *minbuf = rows + kTeamRows;                     // [2][64] partial minima (column halves)
    double *obuf = minbuf + 2 * kTeamItems;                // 2 x [8][L] staging
    for (int i = tid; i < kTeamRows; i += kTeamThreads) rows[i] = 0.0;   // padding slots must be 0
    const unsigned obuf_s = (unsigned)__cvta_generic_to_shared(obuf);

    // B fragments of this warp's two n-tile pairs (column-pair slots 32 warp .. 32 warp + 31)
    double Bp[4][KE], Bq[4][KO > 0 ? KO : 1];
    {
        constexpr int NC = N_ + 1;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            int col = col_of(4 * warp + ni, g);
            if (col > M) col = 0;
#pragma unroll
            for (int ks = 0; ks < KE; ++ks) {
                const int j = 4 * ks + t;
                Bp[ni][ks] = (j <= N_) ? __ldg(A.PQ + (size_t)j * A.LhPad + col) : 0.0;
            }
#pragma unroll
            for (int ks = 0; ks < KO; ++ks) {
                const int j = 4 * ks + t;
                Bq[ni][ks] = (j < N_) ? __ldg(A.PQ + (size_t)(NC + j) * A.LhPad + col) : 0.0;
            }
        }
    }
    __syncthreads();

    // chunks of 64 items over the flattened item list; every team owns a contiguous run
    const long long total = A.nitems * (long long)A.B;
    const long long nch = (total + kTeamItems - 1) / kTeamItems;
    long long ch = (long long)blockIdx.x * nch / gridDim.x;
    const long long ch_end = (long long)(blockIdx.x + 1) * nch / gridDim.x;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(A.out) & 15u) == 0;

    PairCursor cur;
    if (MODE == PAIR && ch < ch_end) {
        const long long f = ch * kTeamItems + tid;
        cur = pair_cursor_at(A, f < total ? f : total - 1);
    }
    for (; ch < ch_end; ++ch) {
        const long long g0 = ch * kTeamItems;
        const int cnt = (int)((total - g0) < kTeamItems ? (total - g0) : kTeamItems);
        // ---- stage 1: thread = item (items past the end recompute the last one: finite rows)
        {
            double s[2 * N_ + 1];
            if (MODE == PAIR) {
                stage1_coeffs<N_, DIM, MODE>(A, PW, DW, cur.b, cur.i, cur.j, s);
                const long long fn = g0 + kTeamItems + tid;
                if (ch + 1 < ch_end) {
                    if (fn < total) cur = pair_cursor_next64(A, cur, g0 + tid);
                    else cur = pair_cursor_at(A, total - 1);
                }
            } else {
                const long long gi = g0 + (tid < cnt ? tid : cnt - 1);
                const int b = (int)(gi / A.nitems);
                stage1_coeffs<N_, DIM, MODE>(A, PW, DW, b, (int)(A.item_begin + gi - (long long)b * A.nitems), 0, s);
            }
            double *row = rows + tid * kRowStride;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                row[slot_e(j)] = s[j] + s[2 * N_ - j];
                row[slot_o(j)] = s[j] - s[2 * N_ - j];
            }
            row[slot_e(N_)] = s[N_];
        }
        __syncthreads();                                   // rows of all 64 items staged

        // ---- stage 2: all m-tiles, this warp's half of the column slots
        const int nmt = (cnt + 7) >> 3;
        const double *ar = rows + g * kRowStride + 4 * t;
        double aE[KE], aO[KO > 0 ? KO : 1];
#pragma unroll
        for (int ks = 0; ks < KE; ++ks) aE[ks] = ar[ks];
#pragma unroll
        for (int ks = 0; ks < KO; ++ks) aO[ks] = ar[16 + ks];
#pragma unroll 2
        for (int mi = 0; mi < nmt; ++mi) {
            const unsigned par = (unsigned)mi & 1u;
            double *ob = obuf + (size_t)par * 8 * L;
            double *of = ob + g * L + 4 * t + 32 * warp;       // forward cursor: column 32 warp + 4 t of row g
            double *om = ob + g * L + M - 4 * t - 32 * warp;   // mirror cursor
            double C[2][2][4];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
#pragma unroll
                for (int u = 0; u < 2; ++u) { C[p][u][0] = A.beta; C[p][u][1] = A.beta; C[p][u][2] = 0.0; C[p][u][3] = 0.0; }
#pragma unroll
                for (int ks = 0; ks < KE; ++ks) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) dmma884(C[p][u][0], C[p][u][1], aE[ks], Bp[2 * p + u][ks]);
                    if (ks < KO) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) dmma884(C[p][u][2], C[p][u][3], aO[ks], Bq[2 * p + u][ks]);
                    }
                }
            }
            if (mi + 1 < nmt) {                                 // A fragments of the next m-tile
                const double *an = ar + (size_t)8 * (mi + 1) * kRowStride;
#pragma unroll
                for (int ks = 0; ks < KE; ++ks) aE[ks] = an[ks];
#pragma unroll
                for (int ks = 0; ks < KO; ++ks) aO[ks] = an[16 + ks];
            }
            double mn = INFINITY;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                double cand[2][2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int cb = 16 * p + 2 * u;
                    if (STORE) {
                        of[cb] = C[p][u][0] + C[p][u][2]; om[-cb] = C[p][u][0] - C[p][u][2];
                        of[cb + 1] = C[p][u][1] + C[p][u][3]; om[-cb - 1] = C[p][u][1] - C[p][u][3];
                    }
                    if (MINMODE) {
                        cand[u][0] = C[p][u][0] - fabs(C[p][u][2]);
                        cand[u][1] = C[p][u][1] - fabs(C[p][u][3]);
                    }
                }
                if (MINMODE) {
                    const double m4 = dmin(dmin(cand[0][0], cand[0][1]), dmin(cand[1][0], cand[1][1]));
                    mn = p == 0 ? m4 : dmin(mn, m4);
                }
            }
            if (MINMODE) {                                      // row g of this m-tile over this warp's columns
                mn = dmin(mn, __shfl_xor_sync(0xffffffffu, mn, 1));
                mn = dmin(mn, __shfl_xor_sync(0xffffffffu, mn, 2));
                if (t == 0) minbuf[warp * kTeamItems + 8 * mi + g] = mn;
            }
            if (STORE) {
                fence_async_smem();                             // generic-proxy writes -> visible to the TMA read
                // the store this warp issued one m-tile ago (other buffer) must have been read before
                // the team writes into that buffer again in the next m-tile
                if (warp == (int)(par ^ 1u) && lane == 0) bulk_wait_read<0>();
            }
            __syncthreads();                                    // both column halves of the block written
            if (STORE && warp == (int)par) {
                const int nrows = (cnt - 8 * mi) < 8 ? (cnt - 8 * mi) : 8;
                double *dst = A.out + ((size_t)g0 + 8 * mi) * L;
                const unsigned bytes = (unsigned)(nrows * L) * 8u;
                if (base_aligned && (nrows == 8 || (bytes & 15u) == 0)) {
                    if (lane == 0) { bulk_store(dst, obuf_s + par * (unsigned)(64 * L), bytes); bulk_commit(); }
                } else {                                        // odd row count x odd L or unaligned base
                    for (int i = lane; i < nrows * L; i += 32) __stcs(dst + i, ob[i]);
                    if (lane == 0) bulk_commit();
                }
            }
        }
        // ---- minima of the 64 items: warp w emits items 32 w .. 32 w + 31 (lane = item)
        if (MINMODE) {
            // (the last team barrier above ordered every minbuf write before these reads; the next
            // writes come after the next chunk's stage-1 barrier)
            const int it = 32 * warp + lane;
            const double v = dmin(minbuf[it], minbuf[kTeamItems + it]);
            emit_minima_item_order(A.sinks, v, g0 + 32 * warp, cnt - 32 * warp, lane);
        }
        if (STORE && (nmt & 1)) {
            // an odd number of m-tiles (last chunk only) would break the buffer / issuer rotation
            if (lane == 0) bulk_wait_read<0>();
            __syncthreads();
        }
        // no barrier here: the next chunk's stage 1 only writes `rows`, which nobody reads after the
        // last m-tile barrier, and `minbuf` is next written behind the stage-1 barrier
    }
    if (STORE && lane == 0) bulk_wait_all();                    // staging buffers must outlive the last bulk reads
}

template <int N_, int DIM, int MODE, int MINMODE, bool STORE>
int launch_sq_elev_team(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    ProdWeights<N_> PW;
    DiffWeights<N_> DW;
    const double scale = A.alpha * (0.5 * (double)DIM);
    for (int i = 0; i <= N_; ++i)
        for (int j = i; j <= N_; ++j) {
            double w = plan->h_W[i * (N_ + 1) + j] * scale;
            PW.w[widx<N_>(i, j)] = (i == j) ? w : 2.0 * w;
        }
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t shmem = ((size_t)kTeamRows + 2 * kTeamItems + (STORE ? 16 * (size_t)A.L : 0)) * sizeof(double);
    auto kern = sq_elev_team_kernel<N_, DIM, MODE, MINMODE, STORE>;
    int sms = 148, per_sm = 1;
    if (int rc = bez_kernel_config((const void *)kern, kTeamThreads, shmem, &sms, &per_sm)) return rc;
    const long long nch = (A.nitems * (long long)A.B + kTeamItems - 1) / kTeamItems;
    long long grid = (long long)sms * per_sm;
    if (grid > nch) grid = nch;
    if (A.sinks.npeers > 0 && grid == (long long)sms * per_sm && grid > 1 && !(A.flags & kFlagFullGridWithPeers)) grid -= 1;
    if (grid < 1) return BEZ_OK;
    kern<<<(unsigned)grid, kTeamThreads, shmem, st>>>(A, PW, DW);
    BEZ_CUDA(cudaGetLastError());
    return BEZ_OK;
}

}  // namespace bezteam
