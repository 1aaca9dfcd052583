// Tensor-path kernels, MODE = PAIR, degrees 7..10 (see sq_elev_mma_kernel.cuh).
#define BEZ_MMA_FN bez_sq_elev_mma_pair_b
#define BEZ_MMA_MODE bezcore::PAIR
#define BEZ_MMA_NLO 7
#define BEZ_MMA_NHI 10
#include "sq_elev_mma_kernel.cuh"
