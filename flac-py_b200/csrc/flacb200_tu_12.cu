/* the LPC pipeline, loops unrolled to order 12 (see flacb200_tu.cuh) */
#define FLACB_TU_LMAX 12
#include "flacb200_tu.cuh"
