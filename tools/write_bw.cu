// Microbenchmark: HBM write-only bandwidth on B200 (the C4 kernels write 508 MB per
// evaluation and read ~0.3 MB, so the copy figure in MEASURED_PEAKS.json -- read and
// write in flight together -- is not necessarily reachable by a pure store stream).
//   (a) cudaMemsetAsync   (b) STG.128 streaming stores, grid-stride
//   (c) TMA bulk stores (cp.async.bulk shared->global) of 7744-byte blocks, like the kernel
//   (d) copy (read+write) for reference
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/write_bw tools/write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_store(double2 *out, size_t n2, double v) {
    const double2 x = make_double2(v, v + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        __stcs(out + i, x);
}
__global__ void k_store_plain(double2 *out, size_t n2, double v) {
    const double2 x = make_double2(v, v + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        out[i] = x;
}
__global__ void k_copy(const double2 *in, double2 *out, size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}
// each warp owns a 7744-byte staging block and streams it out with bulk stores
template <int BYTES>
__global__ void k_tma(double *out, size_t nblocks) {
    extern __shared__ __align__(128) double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *buf = sm + warp * (BYTES / 8);
    for (int i = lane; i < BYTES / 8; i += 32) buf[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + warp, nw = (size_t)gridDim.x * (blockDim.x >> 5);
    if (lane == 0) {
        for (size_t b = gw; b < nblocks; b += nw) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + b * (BYTES / 8)), "r"((unsigned)__cvta_generic_to_shared(buf)), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <class F> float timeit(F f, int reps = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t bytes = (size_t)2 << 30;   // 2 GiB >> 126 MB L2
    double *d, *e;
    cudaMalloc(&d, bytes); cudaMalloc(&e, bytes);
    cudaMemset(e, 0, bytes);
    const size_t n2 = bytes / 16;
    float ms;
    ms = timeit([&] { cudaMemsetAsync(d, 0, bytes); });
    printf("cudaMemsetAsync            : %.3f ms  %.0f GB/s\n", ms, bytes / ms * 1e-6);
    for (int bps : {4, 8, 16}) {
        ms = timeit([&] { k_store<<<148 * bps, 256>>>((double2 *)d, n2, 1.0); });
        printf("STG.128 .cs  %2d CTA/SM     : %.3f ms  %.0f GB/s\n", bps, ms, bytes / ms * 1e-6);
    }
    ms = timeit([&] { k_store_plain<<<148 * 8, 256>>>((double2 *)d, n2, 1.0); });
    printf("STG.128 plain 8 CTA/SM     : %.3f ms  %.0f GB/s\n", ms, bytes / ms * 1e-6);
    {
        constexpr int BYTES = 7744;
        const size_t nblocks = bytes / BYTES;
        for (int warps : {4, 8, 12, 16}) {
            cudaFuncSetAttribute(k_tma<BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * BYTES);
            ms = timeit([&] { k_tma<BYTES><<<148, warps * 32, warps * BYTES>>>(d, nblocks); });
            printf("TMA bulk 7744 B, %2d warps/SM: %.3f ms  %.0f GB/s\n", warps, ms, nblocks * (double)BYTES / ms * 1e-6);
        }
    }
    ms = timeit([&] { k_copy<<<148 * 8, 256>>>((const double2 *)e, (double2 *)d, n2); });
    printf("copy (read+write counted)  : %.3f ms  %.0f GB/s\n", ms, 2.0 * bytes / ms * 1e-6);
    ms = timeit([&] { cudaMemcpyAsync(d, e, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D (r+w)       : %.3f ms  %.0f GB/s\n", ms, 2.0 * bytes / ms * 1e-6);
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
