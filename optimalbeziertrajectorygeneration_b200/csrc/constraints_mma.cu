// Dispatcher of the tensor-path constraint kernels; the kernels themselves are instantiated
// in constraints_mma_<mode>_<range>.cu (parallel compilation).
#include <stdlib.h>

#include "sq_elev_stage1.cuh"

namespace bezcore {

int bez_sq_elev_mma_pair_a(const bez_plan *, const SqElevArgs &, cudaStream_t);
int bez_sq_elev_mma_pair_b(const bez_plan *, const SqElevArgs &, cudaStream_t);
int bez_sq_elev_mma_pair_c(const bez_plan *, const SqElevArgs &, cudaStream_t);
int bez_sq_elev_mma_pair_d(const bez_plan *, const SqElevArgs &, cudaStream_t);
int bez_sq_elev_mma_speed_a(const bez_plan *, const SqElevArgs &, cudaStream_t);
int bez_sq_elev_mma_speed_b(const bez_plan *, const SqElevArgs &, cudaStream_t);

// BEZGPU_FORCE_DFMA=1 keeps every shape on the column-stationary DFMA kernel (A/B runs and
// tests/test_gpu_constraints.py::test_tensor_path_matches_dfma_path); read on every call.
static bool bez_force_dfma() {
    const char *e = getenv("BEZGPU_FORCE_DFMA");
    return e && e[0] == '1';
}

// BEZGPU_MMA_FLAGS=<int>: kFlag* bits for A/B experiments (tile order, L1 prefetch).
int bez_sq_elev_mma_flags() {
    const char *e = getenv("BEZGPU_MMA_FLAGS");
    return e ? atoi(e) : 0;
}

bool bez_sq_elev_mma_supported(const bez_plan *plan) {
#ifdef BEZ_ONLY_N
    if (plan->n != BEZ_ONLY_N) return false;
#endif
    return plan->n <= 15 && plan->Lh <= 64 && plan->dim >= 2 && !bez_force_dfma();
}

// L > 128: the column-tiled tensor-path variant (sq_elev_mma_wide.cuh); it always writes the rows
bool bez_sq_elev_mma_wide_supported(const bez_plan *plan) {
#ifdef BEZ_ONLY_N
    if (plan->n != BEZ_ONLY_N) return false;
#endif
    return plan->n <= 15 && plan->Lh > 64 && plan->dim >= 2 && !bez_force_dfma();
}

int bez_sq_elev_mma(const bez_plan *plan, const SqElevArgs &A, int mode, cudaStream_t st) {
    const int n = plan->n;
    if (mode == PAIR) {
        if (n <= 6) return bez_sq_elev_mma_pair_a(plan, A, st);
        if (n <= 10) return bez_sq_elev_mma_pair_b(plan, A, st);
        if (n <= 13) return bez_sq_elev_mma_pair_c(plan, A, st);
        return bez_sq_elev_mma_pair_d(plan, A, st);
    }
    return n <= 9 ? bez_sq_elev_mma_speed_a(plan, A, st) : bez_sq_elev_mma_speed_b(plan, A, st);
}

}  // namespace bezcore
