// Tensor-path kernels, MODE = PAIR, degrees 11..13 (see sq_elev_mma_kernel.cuh).
#define BEZ_MMA_FN bez_sq_elev_mma_pair_c
#define BEZ_MMA_MODE bezcore::PAIR
#define BEZ_MMA_NLO 11
#define BEZ_MMA_NHI 13
#include "sq_elev_mma_kernel.cuh"
