/*
 * flacb200_launch.h -- the encode kernels are instantiated in four translation units
 * (flacb200_tu_{12,32}_{fused,staged}.cu, one per unroll depth of the LPC loops and per
 * pipeline) so that they compile side by side; this is the interface flacb200.cu calls.
 */
#pragma once

#include <cuda_runtime.h>

#include "flacb200_kernels.cuh"

namespace flacb {

enum KernelKind {
    K_FUSED = 0,     /* encode_frames_kernel<.., 0>: the whole path in one kernel per frame      */
    K_ANALYSE = 1,   /* stage A : analyse_kernel                                                 */
    K_CHOOSE = 2,    /* stage C1: choose_kernel                                                  */
    K_SERIALISE = 3, /* stage C2: encode_frames_kernel<.., 2>                                    */
    K_BACK = 4       /* stage C : encode_frames_kernel<.., 1> (choice + serialisation together)  */
};

/* variant 0: <576, 1> launch bounds, 1: <288, 3> (72 registers) */
#define FLACB_TU_DECL(name)                                                                                       \
    cudaError_t launch_##name(int kind, int variant, unsigned grid, int threads, size_t smem, cudaStream_t stream, \
                              const EncParams &p);                                                                \
    cudaError_t init_##name(const uint16_t *crc_tab, const uint16_t *xpow_lo, const uint16_t *xpow_hi);           \
    size_t static_smem_##name();
FLACB_TU_DECL(12_fused)
FLACB_TU_DECL(12_staged)
FLACB_TU_DECL(32_fused)
FLACB_TU_DECL(32_staged)
#undef FLACB_TU_DECL

} /* namespace flacb */
