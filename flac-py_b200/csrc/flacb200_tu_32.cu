/* the LPC pipeline, loops unrolled to order 32 (see flacb200_tu.cuh) */
#define FLACB_TU_LMAX 32
#include "flacb200_tu.cuh"
