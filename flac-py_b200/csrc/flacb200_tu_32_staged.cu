/* encode kernels with LPC loops unrolled to order 32, staged pipeline (see flacb200_tu.cuh) */
#define FLACB_TU_LMAX 32
#define FLACB_TU_STAGED 1
#include "flacb200_tu.cuh"
