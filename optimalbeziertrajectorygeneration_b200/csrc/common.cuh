// Shared declarations for libbezgpu.so (sm_100a).  See include/bezgpu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "bezgpu.h"

struct bez_plan {
    int n, dim, elev, device;
    int L;      // 2n + elev + 1 : elevated control points per curve
    int Lh;     // (L + 1) / 2   : columns handled by the even/odd elevation
    int LhPad;  // Lh rounded up to a multiple of 32
    // Even/odd folded elevation table, [2n+1][LhPad] (zero padded):
    //   rows 0..n    P[j][i] = (T[j][i] + T[2n-j][i]) / 2   (row n: T[n][i])
    //   rows n+1..2n Q[j][i] = (T[j][i] - T[2n-j][i]) / 2   (j = row-n-1 < n)
    double *d_PQ;
    double *d_T;     // dense elevation matrix [2n+1][L] (FD sweep / generic paths)
    double *d_W;     // product weights [n+1][n+1]
    double *d_E1;    // elevMatrix(n-1, 1) [n][n+1]
    // host copies used to fill by-value kernel parameters
    double h_W[(BEZ_MAX_DEGREE + 1) * (BEZ_MAX_DEGREE + 1)];
    double h_E1lo[BEZ_MAX_DEGREE + 1];   // E1[i][i]   = weight of d_i     in q_i
    double h_E1hi[BEZ_MAX_DEGREE + 1];   // E1[i-1][i] = weight of d_{i-1} in q_i
};

void bez_set_error(const char *fmt, ...);
int bez_cuda_fail(cudaError_t e, const char *what);

#define BEZ_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return bez_cuda_fail(e__, #call); \
    } while (0)

// Entry points run on the device their plan / tables live on and restore the caller's current
// device on return (a host thread that drives several GPUs must not find its device changed).
struct BezDeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit BezDeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = (err == cudaSuccess);
        }
    }
    ~BezDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    BezDeviceGuard(const BezDeviceGuard &) = delete;
    BezDeviceGuard &operator=(const BezDeviceGuard &) = delete;
};
#define BEZ_ON_DEVICE(device)                                                     \
    BezDeviceGuard bez_guard__(device);                                            \
    if (bez_guard__.err != cudaSuccess) return bez_cuda_fail(bez_guard__.err, "cudaSetDevice")

// Launch configuration of a kernel on the *current* device, cached per (kernel, device, shared
// memory size) behind a mutex: opts the kernel in to `shmem` bytes of dynamic shared memory on
// that device (the attribute is per device) and returns the SM count and the resident CTAs per
// SM.  plan.cu.
int bez_kernel_config(const void *func, int threads, size_t shmem, int *sms, int *ctas_per_sm);

#define BEZ_REQUIRE(cond, msg)                 \
    do {                                       \
        if (!(cond)) {                         \
            bez_set_error("%s: %s", __func__, msg); \
            return BEZ_EINVAL;                 \
        }                                      \
    } while (0)

// Row offset of row i in the lexicographic i<j pair list over N items.
__host__ __device__ inline long long bez_pair_row_offset(long long i, long long N) {
    return i * (2 * N - i - 1) / 2;
}

// Inverse of the above: pair index p -> (i, j), i < j.  The row is estimated with a
// single-precision square root (the fp64 sqrt is a ~30-instruction sequence and this runs
// once per 32 outputs rows in kernels that are bound by instruction issue) and then fixed up
// exactly with integer arithmetic, so the result does not depend on the estimate's rounding.
__device__ inline void bez_pair_decode(long long p, int N, int &i, int &j) {
    const float b = 2.0f * (float)N - 1.0f;
    const float disc = (float)((2.0 * (double)N - 1.0) * (2.0 * (double)N - 1.0) - 8.0 * (double)p);
    int ii = (int)((b - __fsqrt_rn(disc > 0.0f ? disc : 0.0f)) * 0.5f);
    if (ii < 0) ii = 0;
    if (ii > N - 2) ii = N - 2;
    while (ii > 0 && bez_pair_row_offset(ii, N) > p) --ii;
    while (ii < N - 2 && bez_pair_row_offset(ii + 1, N) <= p) ++ii;
    i = ii;
    j = (int)(p - bez_pair_row_offset(ii, N)) + ii + 1;
}
