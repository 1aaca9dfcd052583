// Microbenchmark: DFMA latency / throughput on sm_100a (chains per warp x warps per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void chains(double *out, int iters, double a, double b) {
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = threadIdx.x + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fma(acc[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1 << 20] = (double)(t1 - t0);
}
template <int K>
void run(int warps_per_sm, double *d) {
    int iters = 4096;
    chains<K><<<148, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    double cyc;
    cudaMemcpy(&cyc, d + (1 << 20), 8, cudaMemcpyDeviceToHost);
    printf("K=%2d chains, %2d warps/SM: %.2f cycles per DFMA per warp, SM rate %.3f warp-DFMA/cycle\n", K,
           warps_per_sm, cyc / (iters * K), warps_per_sm * (double)iters * K / cyc);
}
int main() {
    double *d;
    cudaMalloc(&d, ((1 << 20) + 8) * 8);
    for (int w : {1, 4, 8, 12, 16, 32}) {
        run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<16>(w, d);
    }
    return 0;
}
