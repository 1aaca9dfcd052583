// Tensor-path kernels, MODE = SPEED, degrees 10..15 (see sq_elev_mma_kernel.cuh).
#define BEZ_MMA_FN bez_sq_elev_mma_speed_b
#define BEZ_MMA_MODE bezcore::SPEED
#define BEZ_MMA_NLO 10
#define BEZ_MMA_NHI 15
#include "sq_elev_mma_kernel.cuh"
