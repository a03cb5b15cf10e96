/* encode kernels with LPC loops unrolled to order 12, one-kernel path (see flacb200_tu.cuh) */
#define FLACB_TU_LMAX 12
#define FLACB_TU_STAGED 0
#include "flacb200_tu.cuh"
