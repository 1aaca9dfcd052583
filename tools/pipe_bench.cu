// Microbenchmarks that decide the stage-2 design of sq_elev_kernel on sm_100a:
//   (1) DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 .f64) throughput per SM
//   (2) DMMA interleaved with DFMA: do the fp64 tensor and vector pipes overlap?
//   (3) DFMA with constant-bank operands sweeping a 10 KB table (IMC behaviour)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pipe_bench tools/pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// K independent accumulator tiles per warp, m8n8k4
template <int K>
__global__ void k_dmma884(double *out, int iters, double a, double b) {
    double c[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) { c[k][0] = threadIdx.x + k; c[k][1] = k; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) dmma884(c[k][0], c[k][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (double)(t1 - t0);
}
template <int K>
__global__ void k_dmma1688(double *out, int iters, double a, double b) {
    double c[K][4];
    double A[4] = {a, a + 1, a + 2, a + 3}, Bf[2] = {b, b + 1};
#pragma unroll
    for (int k = 0; k < K; ++k) { c[k][0] = threadIdx.x + k; c[k][1] = k; c[k][2] = 1; c[k][3] = 2; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) dmma1688(c[k], A, Bf);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (double)(t1 - t0);
}
template <int K>
__global__ void k_dmma16816(double *out, int iters, double a, double b) {
    double c[K][4];
    double A[8], Bf[4];
#pragma unroll
    for (int q = 0; q < 8; ++q) A[q] = a + q;
#pragma unroll
    for (int q = 0; q < 4; ++q) Bf[q] = b + q;
#pragma unroll
    for (int k = 0; k < K; ++k) { c[k][0] = threadIdx.x + k; c[k][1] = k; c[k][2] = 1; c[k][3] = 2; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) dmma16816(c[k], A, Bf);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (double)(t1 - t0);
}
// K DMMA tiles + F DFMA chains interleaved per iteration
template <int K, int F>
__global__ void k_mix(double *out, int iters, double a, double b) {
    double c[K > 0 ? K : 1][2], f[F > 0 ? F : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) { c[k][0] = threadIdx.x + k; c[k][1] = k; }
#pragma unroll
    for (int k = 0; k < F; ++k) f[k] = threadIdx.x + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < (K > F ? K : F); ++k) {
            if (k < K) dmma884(c[k][0], c[k][1], a, b);
            if (k < F) f[k] = fma(f[k], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += c[k][0] + c[k][1];
#pragma unroll
    for (int k = 0; k < F; ++k) s += f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (double)(t1 - t0);
}

// (3) constant-bank operands: lane = item holds NE row values in registers and
// sweeps NCOL columns x NE weights from __constant__ memory.
constexpr int NE = 21, NCOL = 61;
__constant__ double ctab[NE * NCOL];
template <int ITEMS>
__global__ void k_const(double *out, int iters, double seed) {
    double e[ITEMS][NE];
#pragma unroll
    for (int u = 0; u < ITEMS; ++u)
#pragma unroll
        for (int j = 0; j < NE; ++j) e[u][j] = seed * (threadIdx.x + j + u);
    double tot = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            double s[ITEMS][2];
#pragma unroll
            for (int u = 0; u < ITEMS; ++u) { s[u][0] = 0; s[u][1] = 0; }
#pragma unroll
            for (int j = 0; j < NE; ++j)
#pragma unroll
                for (int u = 0; u < ITEMS; ++u) s[u][j & 1] = fma(e[u][j], ctab[c * NE + j], s[u][j & 1]);
#pragma unroll
            for (int u = 0; u < ITEMS; ++u) tot += s[u][0] - s[u][1];
        }
#pragma unroll
        for (int u = 0; u < ITEMS; ++u) e[u][i % NE] += tot * 1e-30;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (double)(t1 - t0);
}

static double cycles(double *d) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return -1; }
    double cyc;
    cudaMemcpy(&cyc, d + (1 << 20), 8, cudaMemcpyDeviceToHost);
    return cyc;
}

int main() {
    double *d;
    cudaMalloc(&d, ((1 << 20) + 8) * 8);
    const int iters = 2048;
    double h[NE * NCOL];
    for (int i = 0; i < NE * NCOL; ++i) h[i] = 1.0 / (1 + i);
    cudaMemcpyToSymbol(ctab, h, sizeof(h));
#define REPORT(name, kern, K, w, fma_per_inst)                                                        \
    {                                                                                                  \
        kern<<<148, (w) * 32>>>(d, iters, 1.0000001, 1e-9);                                            \
        double cyc = cycles(d);                                                                        \
        printf("%-14s K=%2d %2d warps/SM: %.2f cyc/inst/warp, SM rate %.3f inst/cyc = %.1f FMA/clk/SM\n", name, K, w, \
               cyc / (iters * (double)(K)), (w) * (double)iters * (K) / cyc,                            \
               (w) * (double)iters * (K) / cyc * (fma_per_inst));                                      \
    }
    for (int w : {1, 4, 8, 16}) {
        REPORT("dmma.m8n8k4", k_dmma884<1>, 1, w, 256.0);
        REPORT("dmma.m8n8k4", k_dmma884<4>, 4, w, 256.0);
        REPORT("dmma.m8n8k4", k_dmma884<8>, 8, w, 256.0);
        REPORT("dmma.m16n8k8", k_dmma1688<4>, 4, w, 1024.0);
        REPORT("dmma.m16n8k16", k_dmma16816<4>, 4, w, 2048.0);
    }
    // mix: report elapsed cycles per iteration against the two pure runs
    for (int w : {4, 8, 16}) {
#define MIX(K, F)                                                                                      \
    {                                                                                                  \
        k_mix<K, F><<<148, w * 32>>>(d, iters, 1.0000001, 1e-9);                                       \
        double cyc = cycles(d);                                                                        \
        printf("mix dmma884 x%d + dfma x%d, %2d warps/SM: %.1f cycles per iteration per warp\n", K, F, w, \
               cyc / iters);                                                                           \
    }
        MIX(8, 0) MIX(0, 8) MIX(8, 8) MIX(4, 8) MIX(8, 16)
    }
    for (int w : {4, 8, 12, 16}) {
#define CST(I)                                                                                         \
    {                                                                                                  \
        k_const<I><<<148, w * 32>>>(d, 64, 1e-3);                                                      \
        double cyc = cycles(d);                                                                        \
        double nd = 64.0 * NCOL * NE * I;                                                              \
        printf("const-bank DFMA, %d item(s)/lane, %2d warps/SM: %.2f cyc/DFMA/warp, SM rate %.3f warp-DFMA/cyc\n", I, w, \
               cyc / nd, w * nd / cyc);                                                                \
    }
        CST(1) CST(2)
    }
    return 0;
}
