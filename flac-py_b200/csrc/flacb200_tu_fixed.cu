/* the fixed-predictor encode kernel (see flacb200_tu.cuh) */
#define FLACB_TU_LMAX 0
#include "flacb200_tu.cuh"
