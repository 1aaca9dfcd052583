// Stage 2 of the fused "square -> fold -> elevate" kernels on the fp64 tensor
// path (DMMA.8x8x4, mma.sync.m8n8k4.f64) with a TMA bulk-store epilogue.
//
// Why: the column-stationary DFMA sweep of sq_elev_core.cuh streams every item's
// folded row through broadcast LDS.128 -- 25 L1 data-pipe wavefronts and 5.6 KB of
// LSU->RF writeback per item, which ncu showed to be the binding unit (70 % of the
// L1 data pipe at 60 % of HBM, profiles/r01_ncu_pair_kernel_dfma.txt).  One DMMA
// does the work of 8 warp-wide DFMAs with all operands in registers, so the same
// fp64 pipe time (DMMA and DFMA share one 64 FMA/clk/SM pipe, tools/pipe_bench.cu)
// needs ~6x fewer issue slots and no operand traffic; the 8 x L output block of an
// m-tile is staged in shared memory in its final HBM layout and leaves the SM as a
// single cp.async.bulk (TMA) store instead of ~30 LSU store wavefronts.
//
// GEMM view per warp tile of 32 items (folding of sq_elev_core.cuh kept):
//     se[item][c] = beta + sum_j e[item][j] P[j][c]        M = items   (4 m-tiles of 8)
//     so[item][c] =        sum_j o[item][j] Q[j][c]        N = column pairs (<= 8 n-tiles of 8)
//     out[item][c] = se + so,  out[item][L-1-c] = se - so  K = folded index j (k-steps of 4)
// Fragment ownership (PTX ISA, mma.m8n8k4.f64): lane = 4 g + t holds A[g][t],
// B[t][g], C[g][2t], C[g][2t+1].
#pragma once
#include "sq_elev_core.cuh"

namespace bezmma {
using bezcore::dmin;

// Staged item row: e_j at slot_e(j), o_j at slot_o(j), stride kRowStride doubles.
// With slot = 4 (j mod 4) + j / 4 the A-fragment load of a k-step (lane (g,t) reads
// row g, j = 4 ks + t) hits double-bank (g + 4 t + ks) mod 16: conflict free per
// half-warp; the odd row stride keeps the lane = item stores of stage 1 conflict
// free as well.  Slots that no j maps to stay zero (zeroed once per kernel).
constexpr int kRowStride = 33;
constexpr int kRowsDoubles = 32 * kRowStride;      // 8448 B per warp (multiple of 16 B)
__host__ __device__ constexpr int slot_e(int j) { return 4 * (j & 3) + (j >> 2); }
__host__ __device__ constexpr int slot_o(int j) { return 16 + slot_e(j); }

template <int N_> struct Geom {
    static constexpr int KE = (N_ + 1 + 3) / 4;    // k-steps over e_0..e_n
    static constexpr int KO = (N_ + 3) / 4;        // k-steps over o_0..o_{n-1}
    static_assert(KE <= 4, "slot_e needs 4 ks + t < 16: degree <= 15");
};

// Column pair owned by element nn of n-tile ni.  Two n-tiles interleave over 16
// consecutive columns so that the C-fragment stores (lane (g,t): columns c0, c0+1,
// c0 = base + 4 t, row g, row pitch L doubles) touch double-banks g L + 4 t + const:
// conflict free per half-warp whenever L is odd (E even), 2 wavefronts per STS.64.
__host__ __device__ constexpr int col_of(int ni, int nn) {
    return 16 * (ni >> 1) + 2 * (ni & 1) + 4 * (nn >> 1) + (nn & 1);
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// --- TMA bulk store (shared::cta -> global), bulk-group completion ----------------
__device__ __forceinline__ void bulk_store(double *gdst, unsigned ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(PENDING) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// --- TMA bulk load (global -> shared::cta), mbarrier completion --------------------
__device__ __forceinline__ void mbar_init(unsigned mbar_s, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar_s), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned mbar_s, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar_s, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(mbar_s), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned sdst, const void *gsrc, unsigned bytes, unsigned mbar_s) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar_s) : "memory");
}

// Folded elevation weights in B-fragment order, register resident for the whole kernel.
// NP = n-tile pairs of the shape: 16 NP column-pair slots cover Lh <= 16 NP, i.e. L <= 32 NP
// (NP = 4: the headline shapes 65 <= L <= 128; NP = 2: 33..64; NP = 1: L <= 32).
template <int N_, int NP> struct BFrags {
    double p[2 * NP][Geom<N_>::KE];
    double q[2 * NP][Geom<N_>::KO > 0 ? Geom<N_>::KO : 1];
};

template <int N_, int NP>
__device__ __forceinline__ void load_bfrags(BFrags<N_, NP> &B, const double *__restrict__ PQ, int L, int LhPad,
                                            int lane) {
    constexpr int NC = N_ + 1;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ni = 0; ni < 2 * NP; ++ni) {
        // Slots Lh <= c <= M hold mirrored weights (plan.cu): valid duplicates.  Slots beyond the
        // row (c > M, only possible for L < 16 NP, i.e. NP = 1 and L < 16) take the weights of
        // column pair 0: their values are valid duplicates for the minimum, their stores are
        // predicated off in the epilogue.
        int col = col_of(ni, g);
        if (col > L - 1) col = 0;
#pragma unroll
        for (int ks = 0; ks < Geom<N_>::KE; ++ks) {
            const int j = 4 * ks + t;
            B.p[ni][ks] = (j <= N_) ? __ldg(PQ + (size_t)j * LhPad + col) : 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < Geom<N_>::KO; ++ks) {
            const int j = 4 * ks + t;
            B.q[ni][ks] = (j < N_) ? __ldg(PQ + (size_t)(NC + j) * LhPad + col) : 0.0;
        }
    }
}

// Where the per-item minima of a launch go (all optional).  f = flattened item index
// b * nitems + item of the launch.
struct MinSinks {
    double *itemmin;                    // [B][pitch] (pitch = nitems unless min_pitch > 0)
    long long min_pitch, nitems;
    double *peer_min[BEZ_MAX_PEERS];    // fused all-gather: the same matrix on other GPUs (NVLink)
    int npeers;
    unsigned *mask;                     // bit f of word f >> 5 = (min < threshold)
    double threshold;
    unsigned long long *list_count;     // compacted (f, min) list of the items with min < threshold,
    long long *list_idx;                //   warp-aggregated atomic append (order unspecified);
    double *list_val;                   //   entries past list_cap are dropped but still counted
    long long list_cap;
};

// v = minimum of item (8 t + g) held by lane (g, t) = 4 g + t  ->  all sinks.  g0 = flattened
// index of the tile's first item (a multiple of 32), cnt = live items of the tile.
__device__ __forceinline__ void emit_minima_item_order(const MinSinks &S, double vi, long long g0, int cnt, int lane);
__device__ __forceinline__ void emit_minima(const MinSinks &S, double v, long long g0, int cnt, int lane) {
    // to item order: lane l takes the minimum of item l (coalesced stores, ballot bit = item)
    const double vi = __shfl_sync(0xffffffffu, v, 4 * (lane & 7) + (lane >> 3));
    emit_minima_item_order(S, vi, g0, cnt, lane);
}

// The same for a warp whose lane l already holds the minimum of item g0 + l (cnt <= 0: none live).
__device__ __forceinline__ void emit_minima_item_order(const MinSinks &S, double vi, long long g0, int cnt, int lane) {
    const bool valid = lane < cnt;
    const long long f = g0 + lane;
    long long dst = f;
    if (S.min_pitch > 0) {
        const long long b = f / S.nitems;
        dst = b * S.min_pitch + (f - b * S.nitems);
    }
    if (valid) {
        if (S.itemmin) S.itemmin[dst] = vi;
        if (S.npeers > 0) {                                 // warp-uniform
#pragma unroll
            for (int q = 0; q < BEZ_MAX_PEERS; ++q)
                if (q < S.npeers) S.peer_min[q][dst] = vi;
        }
    }
    if (S.mask || S.list_count) {                       // warp-uniform
        const bool act = valid && vi < S.threshold;
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        if (S.mask && lane == 0 && cnt > 0) S.mask[g0 >> 5] = bal;
        if (S.list_count && bal) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(S.list_count, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            const long long pos = (long long)base + __popc(bal & ((1u << lane) - 1u));
            if (act && pos < S.list_cap) { S.list_idx[pos] = f; S.list_val[pos] = vi; }
        }
    }
}

// One warp tile: rows = staged (e,o) rows of 32 items, obuf = 2 x [8][L] doubles of
// staging, obuf_s = its shared-window address,
// outg = global address of the tile's first output row (rows contiguous, pitch L).
// base_aligned: the output base address is a multiple of 16 bytes (then every full m-tile
// block is, too).  m-tile mi uses staging buffer mi & 1 (at most one bulk read is left pending).
// MINMODE: per-item minimum -> emit_minima;  STORE = false: the rows are not written at all
// (callers that only want the minima, e.g. the sequential-planning constraint).
//
// Schedule of one m-tile (8 items): the n-tiles are processed in pairs; the DMMAs of
// pair p+1 (4 independent accumulator chains, round-robin over the k-steps) are issued
// before the epilogue of pair p, so the fp64 pipe always has independent work queued
// behind the DADD/STS of the epilogue.
//
// Deferred stores (early_store = false): the proxy fence and the bulk store of m-tile mi are
// issued behind the first DMMA group of m-tile mi + 1.  `fence.proxy.async` lowers to
// MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, which waits for the warp's outstanding shared-memory
// stores; right behind the 32 STS of an epilogue that wait is ~130 cycles of an in-order warp
// (ncu: long-scoreboard samples on the fence), a dozen DMMAs later it is free.
// rows_free(): called once per tile when the staged rows are dead (the A fragments of the last
// m-tile are in registers, a proxy fence has just been executed) -- the kernel starts the TMA
// fetch of the next tile's vehicle rows into the same shared-memory region there.
struct NoRowsHook { __device__ __forceinline__ void operator()() const {} };

template <int N_, int NP, int MINMODE, bool STORE, class RowsFree = NoRowsHook, int MT_UNROLL = 4>
__device__ __forceinline__ void mma_tile(const double *rows, double *obuf, unsigned obuf_s,
                                         const BFrags<N_, NP> &B, double *__restrict__ outg,
                                         const MinSinks &S, long long g0, int cnt,
                                         int L, double beta, int lane, bool base_aligned,
                                         bool early_store = false, RowsFree rows_free = RowsFree()) {
    constexpr int KE = Geom<N_>::KE, KO = Geom<N_>::KO;
    const int g = lane >> 2, t = lane & 3;
    const int M = L - 1;
    double mnv[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) mnv[mi] = INFINITY;
    const double *ar = rows + g * kRowStride + 4 * t;
    double aE[KE], aO[KO > 0 ? KO : 1];
#pragma unroll
    for (int ks = 0; ks < KE; ++ks) aE[ks] = ar[ks];
#pragma unroll
    for (int ks = 0; ks < KO; ++ks) aO[ks] = ar[16 + ks];

    // staged 8 x L block of m-tile mi -> HBM (one TMA bulk store)
    auto store_block = [&](int mi) {
        const unsigned par = (unsigned)mi & 1u;
        fence_async_smem();                         // generic-proxy writes -> visible to the TMA read
        __syncwarp();
        const int nrows = (cnt - 8 * mi) < 8 ? (cnt - 8 * mi) : 8;
        double *dst = outg + (size_t)8 * mi * L;
        const unsigned bytes = (unsigned)(nrows * L) * 8u;
        if (base_aligned && (nrows == 8 || (bytes & 15u) == 0)) {
            if (lane == 0) { bulk_store(dst, obuf_s + par * (unsigned)(64 * L), bytes); bulk_commit(); }
        } else {                                    // odd row count x odd L or unaligned base
            const double *ob = obuf + (size_t)par * 8 * L;
            for (int i = lane; i < nrows * L; i += 32) __stcs(dst + i, ob[i]);
            if (lane == 0) bulk_commit();           // empty group keeps the count in step
        }
    };
    bool rows_released = false;

#pragma unroll(MT_UNROLL)
    for (int mi = 0; mi < 4; ++mi) {
        if (8 * mi >= cnt) {                        // short tile (the last one of a launch)
            mnv[0] = mnv[1]; mnv[1] = mnv[2]; mnv[2] = mnv[3]; mnv[3] = INFINITY;
            continue;
        }
        // staging buffer mi & 1: every tile but the very last of a launch has 4 m-tiles (the
        // tiles run over the flattened item list), so the parity is a compile-time constant;
        // a short tile drains the bulk groups at its end (below)
        const unsigned par = (unsigned)mi & 1u;
        double *ob = obuf + (size_t)par * 8 * L;
        double *of = ob + g * L + 4 * t;           // forward cursor of this lane (column 4 t of row g)
        double *om = ob + g * L + M - 4 * t;       // mirror cursor
        double mnp[NP];                             // one minimum per n-tile pair: short dependency chains
        double C[2][2][4];
        auto mma_pair = [&](int p, double (&c)[2][4]) {
#pragma unroll
            for (int u = 0; u < 2; ++u) { c[u][0] = beta; c[u][1] = beta; c[u][2] = 0.0; c[u][3] = 0.0; }
#pragma unroll
            for (int ks = 0; ks < KE; ++ks) {
#pragma unroll
                for (int u = 0; u < 2; ++u) dmma884(c[u][0], c[u][1], aE[ks], B.p[2 * p + u][ks]);
                if (ks < KO) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) dmma884(c[u][2], c[u][3], aO[ks], B.q[2 * p + u][ks]);
                }
            }
        };
        auto epilogue = [&](int p, const double (&c)[2][4]) {
            double cand[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int cb = 16 * p + 2 * u;              // first column of n-tile 2 p + u
                if (STORE) {
                    const double f0 = c[u][0] + c[u][2], m0 = c[u][0] - c[u][2];
                    const double f1 = c[u][1] + c[u][3], m1 = c[u][1] - c[u][3];
                    if (NP > 1) {
                        // every column slot is a valid output pair (slots >= Lh duplicate their
                        // mirror slot bit for bit, see plan.cu): no liveness test on this path
                        of[cb] = f0; om[-cb] = m0; of[cb + 1] = f1; om[-cb - 1] = m1;
                    } else {                                // L < 16 is possible: slots beyond the row
                        if (cb + 4 * t <= M) { of[cb] = f0; om[-cb] = m0; }
                        if (cb + 4 * t + 1 <= M) { of[cb + 1] = f1; om[-cb - 1] = m1; }
                    }
                }
                if (MINMODE) {                              // min(se+so, se-so) = se - |so|, one DADD
                    cand[u][0] = c[u][0] - fabs(c[u][2]);
                    cand[u][1] = c[u][1] - fabs(c[u][3]);
                }
            }
            if (MINMODE) mnp[p] = dmin(dmin(cand[0][0], cand[0][1]), dmin(cand[1][0], cand[1][1]));
        };

        mma_pair(0, C[0]);
        if (STORE) {
            if (mi > 0 && !early_store) store_block(mi - 1);
            // the bulk read of this staging buffer (issued two m-tiles ago) must be done
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
        }
        if (mi == 3) { rows_free(); rows_released = true; }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            if (p < NP - 1) mma_pair(p + 1, C[(p + 1) & 1]);
            else if (mi < 3) {                              // A fragments of the next m-tile
                const double *an = ar + (size_t)8 * (mi + 1) * kRowStride;
#pragma unroll
                for (int ks = 0; ks < KE; ++ks) aE[ks] = an[ks];
#pragma unroll
                for (int ks = 0; ks < KO; ++ks) aO[ks] = an[16 + ks];
            }
            epilogue(p, C[p & 1]);
        }
        if (MINMODE) {
            double m = mnp[0];
#pragma unroll
            for (int p = 1; p < NP; ++p) m = dmin(m, mnp[p]);
            // mnv[k] = minimum of m-tile k after the fourth pass (static register indices even
            // when the loop is not unrolled)
            mnv[0] = mnv[1]; mnv[1] = mnv[2]; mnv[2] = mnv[3]; mnv[3] = m;
        }
        if (STORE && (early_store || mi == 3 || 8 * (mi + 1) >= cnt)) store_block(mi);
    }
    if (!rows_released) rows_free();                // short tile (the last one of a launch)
    if (STORE && cnt <= 24) {                       // short tile: realign the buffer rotation
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
    // per-item minima: lane (g,t) holds the partial minima of row g of the 4 m-tiles and ends up
    // with the full minimum of m-tile t (item 8 t + g) -- a reduce-scatter over the 4 lanes of a
    // row: 3 shuffles + 3 minima instead of 8 + 8
    if (MINMODE) {
        const bool b0 = t & 1, b1 = t & 2;
        const double r0 = __shfl_xor_sync(0xffffffffu, b0 ? mnv[0] : mnv[1], 1);
        const double r1 = __shfl_xor_sync(0xffffffffu, b0 ? mnv[2] : mnv[3], 1);
        const double a0 = dmin(b0 ? mnv[1] : mnv[0], r0);      // m-tile (t & 1)
        const double a1 = dmin(b0 ? mnv[3] : mnv[2], r1);      // m-tile (t & 1) + 2
        const double r2 = __shfl_xor_sync(0xffffffffu, b1 ? a0 : a1, 2);
        const double v = dmin(b1 ? a1 : a0, r2);               // m-tile t
        emit_minima(S, v, g0, cnt, lane);
    }
}

}  // namespace bezmma
