// Tensor-path kernels, MODE = SPEED, degrees 1..9 (see sq_elev_mma_kernel.cuh).
#define BEZ_MMA_FN bez_sq_elev_mma_speed_a
#define BEZ_MMA_MODE bezcore::SPEED
#define BEZ_MMA_NLO 1
#define BEZ_MMA_NHI 9
#include "sq_elev_mma_kernel.cuh"
