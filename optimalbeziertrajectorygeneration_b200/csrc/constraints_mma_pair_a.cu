// Tensor-path kernels, MODE = PAIR, degrees 1..6 (see sq_elev_mma_kernel.cuh).
#define BEZ_MMA_FN bez_sq_elev_mma_pair_a
#define BEZ_MMA_MODE bezcore::PAIR
#define BEZ_MMA_NLO 1
#define BEZ_MMA_NHI 6
#include "sq_elev_mma_kernel.cuh"
