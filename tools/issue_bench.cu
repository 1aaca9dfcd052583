// Can other warps of the same scheduler issue while a warp streams DMMAs?  (round 2)
// ptxas gives every DMMA.8x8x4 a 16-cycle issue stall; this measures whether that stall blocks
// only the issuing warp (then more warps per scheduler hide stage 1 / the epilogue) or the whole
// SMSP issue port.  8 warps per CTA, one CTA per SM: warps 0-3 (one per SMSP) stream DMMAs,
// warps 4-7 (same SMSPs) run a second instruction stream: nothing / integer ALU / DADD / DFMA
// (register operands) / DFMA (constant-bank operand) / STS.64 / SHFL.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/issue_bench tools/issue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__constant__ double cw[64];

// MODE of the second stream: 0 none, 1 IMAD/LOP chains, 2 DADD, 3 DFMA reg, 4 DFMA const, 5 STS.64, 6 SHFL
template <int MODE>
__global__ void k(double *out, long long *cyc, int iters, int with_dmma, double a, double b) {
    __shared__ double sm[8 * 32 * 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = 0, t1 = 0;
    double s = 0;
    if (warp < 4) {
        if (with_dmma) {
            double c[8][2];
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q][0] = lane + q; c[q][1] = q; }
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int q = 0; q < 8; ++q) dmma884(c[q][0], c[q][1], a, b);
            }
            t1 = clock64();
#pragma unroll
            for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
        }
    } else if (MODE > 0) {
        double f[16];
        int u[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) { f[q] = lane + q; u[q] = lane * 7 + q; }
        double *my = sm + (warp * 32 + lane) * 4;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (MODE == 1) u[q] = u[q] * 3 + (u[q] ^ i);
                if (MODE == 2) f[q] = f[q] + a;
                if (MODE == 3) f[q] = fma(f[q], a, b);
                if (MODE == 4) f[q] = fma(f[q], cw[q], b);
                if (MODE == 5) { my[q & 3] = f[q]; }
                if (MODE == 6) f[q] = __shfl_xor_sync(0xffffffffu, f[q], 1 + (q & 15));
            }
            if (MODE == 5) f[i & 15] += my[(i + 1) & 3];
        }
        t1 = clock64();
#pragma unroll
        for (int q = 0; q < 16; ++q) s += f[q] + u[q];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && lane == 0) cyc[warp] = t1 - t0;
}

template <int MODE>
static void run(const char *name, double *d, long long *c) {
    const int iters = 4096;
    long long h[8];
    double alone2 = 0, dm_alone = 0;
    for (int with = 0; with < 2; ++with) {
        k<MODE><<<148, 256>>>(d, c, iters, with, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        if (!with) alone2 = (double)h[4] / iters;
        else {
            if (MODE == 0) dm_alone = (double)h[0] / iters;
            printf("%-22s second stream alone %7.1f cyc/iter (16 inst) | together: DMMA warp %7.1f cyc per 8 DMMA, "
                   "second warp %7.1f cyc/iter\n", name, alone2, (double)h[0] / iters, (double)h[4] / iters);
        }
    }
    (void)dm_alone;
}

int main() {
    double *d;
    long long *c;
    cudaMalloc(&d, 148 * 256 * 8);
    cudaMalloc(&c, 64);
    double h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-9 * i;
    cudaMemcpyToSymbol(cw, h, sizeof(h));
    run<0>("none", d, c);
    run<1>("IMAD+LOP (ALU)", d, c);
    run<2>("DADD", d, c);
    run<3>("DFMA reg", d, c);
    run<4>("DFMA const-bank", d, c);
    run<5>("STS.64", d, c);
    run<6>("SHFL", d, c);
    return 0;
}
