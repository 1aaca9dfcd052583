// Column tiling of the tensor-path kernel for L > 128 (more than 64 column pairs).
//
// Same stage 1 and the same DMMA GEMM as sq_elev_mma_kernel, but the folded elevation table no
// longer fits the register file as B fragments: the column pairs are processed in chunks of 64
// (8 n-tiles = 4 n-tile pairs, exactly the NP = 4 machinery), the B fragments of a chunk are
// reloaded per (tile, chunk) from the L1/L2-resident table (48 loads for 192 DMMAs), and a chunk's
// part of an m-tile -- 64 forward columns and 64 mirror columns of 8 rows -- is staged in shared
// memory and written with coalesced streaming stores (the two 512-byte pieces of a row are not a
// contiguous block of the output, so there is no single TMA bulk store; rows of odd L are only
// 8-byte aligned).  The per-item minimum accumulates across the chunks in registers.
#pragma once
#include "sq_elev_stage1.cuh"

namespace bezwide {
using namespace bezcore;
using namespace bezmma;

constexpr int kChunk = 64;                          // column pairs per chunk

template <int N_, int DIM, int MODE, int MINMODE>
__global__ void __launch_bounds__(kThreads, 2)
sq_elev_mma_wide_kernel(const SqElevArgs A, const ProdWeights<N_> PW, const DiffWeights<N_> DW) {
    constexpr int KE = Geom<N_>::KE, KO = Geom<N_>::KO, NC = N_ + 1;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int L = A.L, M = L - 1;
    constexpr int kStage = 8 * 2 * kChunk;          // [8 rows][64 forward | 64 mirror]
    double *rows = smem + warp * (size_t)(kRowsDoubles + kStage);
    double *ob = rows + kRowsDoubles;
    for (int i = lane; i < kRowsDoubles; i += 32) rows[i] = 0.0;
    __syncwarp();
    const int nchunks = (A.Lh + kChunk - 1) / kChunk;

    const long long total = A.nitems * (long long)A.B;
    const long long nwt = (total + 31) >> 5;
    const long long gwarp = (long long)blockIdx.x * kWarps + warp;
    const long long nwarps = (long long)gridDim.x * kWarps;
    for (long long wt = gwarp; wt < nwt; wt += nwarps) {
        const long long g0 = wt << 5;
        const int cnt = (int)((total - g0) < 32 ? (total - g0) : 32);
        {
            double s[2 * N_ + 1];
            const long long gi = g0 + (lane < cnt ? lane : cnt - 1);
            const int b = (int)(gi / A.nitems);
            const long long it = A.item_begin + gi - (long long)b * A.nitems;
            int vi = (int)it, vj = 0;
            if (MODE == PAIR) bez_pair_decode(it, A.N, vi, vj);
            stage1_coeffs<N_, DIM, MODE>(A, PW, DW, b, vi, vj, s);
            double *row = rows + lane * kRowStride;
#pragma unroll
            for (int j = 0; j < N_; ++j) {
                row[slot_e(j)] = s[j] + s[2 * N_ - j];
                row[slot_o(j)] = s[j] - s[2 * N_ - j];
            }
            row[slot_e(N_)] = s[N_];
        }
        __syncwarp();
        double mnv[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
        for (int ch = 0; ch < nchunks; ++ch) {
            // B fragments of this chunk; slots past the row (column > M) or past the table take the
            // weights of column pair 0 (valid duplicates for the minimum, never stored)
            double Bp[8][KE], Bq[8][KO > 0 ? KO : 1];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                int col = kChunk * ch + col_of(ni, g);
                if (col > M || col >= A.LhPad) col = 0;
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
            const int c0 = kChunk * ch;                          // first column pair of the chunk
#pragma unroll 1
            for (int mi = 0; mi < 4; ++mi) {
                if (8 * mi >= cnt) break;
                const double *ar = rows + (8 * mi + g) * kRowStride + 4 * t;
                double aE[KE], aO[KO > 0 ? KO : 1];
#pragma unroll
                for (int ks = 0; ks < KE; ++ks) aE[ks] = ar[ks];
#pragma unroll
                for (int ks = 0; ks < KO; ++ks) aO[ks] = ar[16 + ks];
                double mn = INFINITY;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    double c[2][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) { c[u][0] = A.beta; c[u][1] = A.beta; c[u][2] = 0.0; c[u][3] = 0.0; }
#pragma unroll
                    for (int ks = 0; ks < KE; ++ks) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) dmma884(c[u][0], c[u][1], aE[ks], Bp[2 * p + u][ks]);
                        if (ks < KO) {
#pragma unroll
                            for (int u = 0; u < 2; ++u) dmma884(c[u][2], c[u][3], aO[ks], Bq[2 * p + u][ks]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int sl = 16 * p + 2 * u + 4 * t;   // slot of this lane's first column in the chunk
                        double *fw = ob + g * (2 * kChunk) + sl;
                        double *mr = ob + g * (2 * kChunk) + kChunk + (kChunk - 1 - sl);
                        fw[0] = c[u][0] + c[u][2]; mr[0] = c[u][0] - c[u][2];
                        fw[1] = c[u][1] + c[u][3]; mr[-1] = c[u][1] - c[u][3];
                        if (MINMODE) {
                            const double m2 = dmin(c[u][0] - fabs(c[u][2]), c[u][1] - fabs(c[u][3]));
                            mn = dmin_nan(mn, m2);
                        }
                    }
                }
                if (MINMODE) mnv[mi] = dmin_nan(mnv[mi], mn);
                __syncwarp();
                // coalesced stores of the chunk: forward columns c0 .. c0+63 and their mirrors
                const int nrows = (cnt - 8 * mi) < 8 ? (cnt - 8 * mi) : 8;
                double *dst = A.out + ((size_t)g0 + 8 * mi) * L;
                for (int i = lane; i < nrows * 2 * kChunk; i += 32) {
                    const int r = i >> 7, q = i & 127;
                    const int slot = q < kChunk ? q : (kChunk - 1 - (q - kChunk));
                    const int col = q < kChunk ? c0 + slot : M - c0 - slot;
                    // slots c0 + slot < Lh and their mirrors cover every column exactly once (the middle
                    // column of an odd L twice, with the same value); slots >= Lh only feed the minimum
                    if (c0 + slot < A.Lh) __stcs(dst + (size_t)r * L + col, ob[r * (2 * kChunk) + q]);
                }
                __syncwarp();
            }
        }
        if (MINMODE) {
            // lane (g,t): partial minima of row g of the 4 m-tiles -> full minimum of item 8 t + g
            const bool b0 = t & 1, b1 = t & 2;
            const double r0 = __shfl_xor_sync(0xffffffffu, b0 ? mnv[0] : mnv[1], 1);
            const double r1 = __shfl_xor_sync(0xffffffffu, b0 ? mnv[2] : mnv[3], 1);
            const double a0 = dmin_nan(b0 ? mnv[1] : mnv[0], r0);
            const double a1 = dmin_nan(b0 ? mnv[3] : mnv[2], r1);
            const double r2 = __shfl_xor_sync(0xffffffffu, b1 ? a0 : a1, 2);
            const double v = dmin_nan(b1 ? a1 : a0, r2);
            emit_minima(A.sinks, v, g0, cnt, lane);
        }
        __syncwarp();
    }
}

template <int N_, int DIM, int MODE, int MINMODE>
int launch_sq_elev_mma_wide(const bez_plan *plan, const SqElevArgs &A, cudaStream_t st) {
    ProdWeights<N_> PW;
    DiffWeights<N_> DW;
    const double scale = A.alpha * (0.5 * (double)DIM);
    for (int i = 0; i <= N_; ++i)
        for (int j = i; j <= N_; ++j) {
            double w = plan->h_W[i * (N_ + 1) + j] * scale;
            PW.w[widx<N_>(i, j)] = (i == j) ? w : 2.0 * w;
        }
    for (int i = 0; i <= N_; ++i) { DW.lo[i] = plan->h_E1lo[i]; DW.hi[i] = plan->h_E1hi[i]; }
    const size_t shmem = (size_t)kWarps * (kRowsDoubles + 8 * 2 * kChunk) * sizeof(double);
    auto kern = sq_elev_mma_wide_kernel<N_, DIM, MODE, MINMODE>;
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

}  // namespace bezwide
