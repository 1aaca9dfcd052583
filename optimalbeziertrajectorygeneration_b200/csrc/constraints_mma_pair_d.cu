// Tensor-path kernels, MODE = PAIR, degrees 14..15 (see sq_elev_mma_kernel.cuh).
#define BEZ_MMA_FN bez_sq_elev_mma_pair_d
#define BEZ_MMA_MODE bezcore::PAIR
#define BEZ_MMA_NLO 14
#define BEZ_MMA_NHI 15
#include "sq_elev_mma_kernel.cuh"
