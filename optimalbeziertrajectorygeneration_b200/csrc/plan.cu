// Plan management + error plumbing for libbezgpu.so.  See include/bezgpu.h.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

void bez_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int bez_cuda_fail(cudaError_t e, const char *what) {
    bez_set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

extern "C" const char *bez_last_error(void) { return g_err; }

// See common.cuh.  The dynamic-shared-memory opt-in is a per-device function attribute and this
// library may be driven from several host threads / for several devices of one process.
namespace {
struct KernelConfig {
    const void *func;
    int device, threads;
    size_t shmem;
    int sms, per_sm;
};
std::mutex g_cfg_mutex;
std::vector<KernelConfig> g_cfg;
}  // namespace

int bez_kernel_config(const void *func, int threads, size_t shmem, int *sms, int *ctas_per_sm) {
    int dev = 0;
    BEZ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    for (const KernelConfig &c : g_cfg)
        if (c.func == func && c.device == dev && c.threads == threads && c.shmem == shmem) {
            *sms = c.sms;
            *ctas_per_sm = c.per_sm;
            return BEZ_OK;
        }
    KernelConfig c{func, dev, threads, shmem, 148, 1};
    // the opt-in only ever grows: a smaller request must not lower what an earlier (cached)
    // configuration of the same kernel on this device relies on
    size_t granted = 48 * 1024;
    for (const KernelConfig &o : g_cfg)
        if (o.func == func && o.device == dev && o.shmem > granted) granted = o.shmem;
    if (shmem > granted)
        BEZ_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    BEZ_CUDA(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
    BEZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.per_sm, func, threads, shmem));
    if (c.per_sm < 1) c.per_sm = 1;
    g_cfg.push_back(c);
    *sms = c.sms;
    *ctas_per_sm = c.per_sm;
    return BEZ_OK;
}
extern "C" int bez_version(void) { return 100; }

extern "C" int bez_plan_create(int n, int dim, int elev, int device,
                               const double *h_prodW, const double *h_elevT,
                               const double *h_elev1, bez_plan **out) {
    BEZ_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    BEZ_REQUIRE(h_prodW && h_elevT, "tables are NULL");
    BEZ_REQUIRE(dim >= 1 && dim <= 3, "dim must be 1, 2 or 3");
    BEZ_REQUIRE(elev >= 0, "elev must be >= 0");
    if (n < 1 || n > BEZ_MAX_DEGREE) {
        bez_set_error("bez_plan_create: degree %d outside [1, %d]", n, BEZ_MAX_DEGREE);
        return BEZ_EUNSUPPORTED;
    }
    BEZ_REQUIRE(h_elev1 != nullptr, "elev1 table is NULL");
    BEZ_ON_DEVICE(device);

    bez_plan *p = (bez_plan *)calloc(1, sizeof(bez_plan));
    if (!p) {
        bez_set_error("bez_plan_create: out of host memory");
        return BEZ_ENOMEM;
    }
    p->n = n; p->dim = dim; p->elev = elev; p->device = device;
    const int m = 2 * n;
    p->L = m + elev + 1;
    p->Lh = (p->L + 1) / 2;
    p->LhPad = (p->Lh + 31) / 32 * 32;
    const int L = p->L, M = m + elev;

    memcpy(p->h_W, h_prodW, sizeof(double) * (n + 1) * (n + 1));
    // elevMatrix(n-1,1): q_i = E1[i-1][i]*d_{i-1} + E1[i][i]*d_i   (bezier.py:519, 1141-1147)
    for (int i = 0; i <= n; ++i) {
        p->h_E1lo[i] = (i < n) ? h_elev1[i * (n + 1) + i] : 0.0;
        p->h_E1hi[i] = (i > 0) ? h_elev1[(i - 1) * (n + 1) + i] : 0.0;
    }

    // fold the elevation matrix: column i and its mirror M-i share the weights
    // T[j][M-i] = T[m-j][i], so with e_j = s_j + s_{m-j}, o_j = s_j - s_{m-j}
    //   b_i     = sum_j e_j P[j][i] + sum_j o_j Q[j][i]
    //   b_{M-i} = sum_j e_j P[j][i] - sum_j o_j Q[j][i]
    std::vector<double> PQ((size_t)(m + 1) * p->LhPad, 0.0);
    for (int i = 0; i < p->Lh; ++i) {
        for (int j = 0; j < n; ++j) {
            double a = h_elevT[(size_t)j * L + i], b = h_elevT[(size_t)(m - j) * L + i];
            PQ[(size_t)j * p->LhPad + i] = (a + b) * 0.5;
            PQ[(size_t)(n + 1 + j) * p->LhPad + i] = (a - b) * 0.5;
        }
        PQ[(size_t)n * p->LhPad + i] = h_elevT[(size_t)n * L + i];
    }
    // Padding columns c >= Lh that still lie inside the row (c <= M) get the weights of their
    // mirror column M - c with the odd part negated: the tensor-path kernels then compute
    // every one of their LhPad column slots as a *valid* output pair -- slot c yields
    // (b_c, b_{M-c}) = (mirror, forward) of slot M - c, bit for bit, so the duplicate stores
    // are benign and no liveness test is needed.  The DFMA kernels ignore columns >= Lh.
    for (int i = p->Lh; i < p->LhPad; ++i) {
        const int mi = M - i;
        if (mi < 0 || mi >= p->Lh) continue;
        for (int j = 0; j <= n; ++j) PQ[(size_t)j * p->LhPad + i] = PQ[(size_t)j * p->LhPad + mi];
        for (int j = 0; j < n; ++j)
            PQ[(size_t)(n + 1 + j) * p->LhPad + i] = -PQ[(size_t)(n + 1 + j) * p->LhPad + mi];
    }

    cudaError_t e;
#define PLAN_ALLOC_COPY(dst, src, count)                                              \
    e = cudaMalloc((void **)&(dst), sizeof(double) * (count));                        \
    if (e == cudaSuccess)                                                             \
        e = cudaMemcpy((dst), (src), sizeof(double) * (count), cudaMemcpyHostToDevice); \
    if (e != cudaSuccess) { bez_plan_destroy(p); return bez_cuda_fail(e, "plan table upload"); }
    PLAN_ALLOC_COPY(p->d_PQ, PQ.data(), PQ.size());
    PLAN_ALLOC_COPY(p->d_T, h_elevT, (size_t)(m + 1) * L);
    PLAN_ALLOC_COPY(p->d_W, h_prodW, (size_t)(n + 1) * (n + 1));
    PLAN_ALLOC_COPY(p->d_E1, h_elev1, (size_t)n * (n + 1));
#undef PLAN_ALLOC_COPY
    *out = p;
    return BEZ_OK;
}

extern "C" int bez_plan_destroy(bez_plan *p) {
    if (!p) return BEZ_OK;
    cudaFree(p->d_PQ);
    cudaFree(p->d_T);
    cudaFree(p->d_W);
    cudaFree(p->d_E1);
    free(p);
    return BEZ_OK;
}

extern "C" int bez_plan_info(const bez_plan *p, int *n, int *dim, int *elev, int *L) {
    BEZ_REQUIRE(p != nullptr, "plan is NULL");
    if (n) *n = p->n;
    if (dim) *dim = p->dim;
    if (elev) *elev = p->elev;
    if (L) *L = p->L;
    return BEZ_OK;
}
