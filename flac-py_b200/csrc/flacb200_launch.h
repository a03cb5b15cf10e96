/*
 * flacb200_launch.h -- the encode kernels are instantiated in three translation units
 * (flacb200_tu_fixed.cu: fixed predictors only; flacb200_tu_12.cu / flacb200_tu_32.cu: the LPC
 * pipeline with its loops unrolled to order 12 / 32) so that they compile side by side; this is
 * the interface flacb200.cu calls.
 */
#pragma once

#include <cuda_runtime.h>

#include "flacb200_kernels.cuh"
#include "flacb200_choose_mma.cuh"

namespace flacb {

enum KernelKind {
    K_FIXED = 0,     /* encode_frames_kernel<.., false>: fixed predictors only, one kernel       */
    K_ANALYSE = 1,   /* stage A : analyse_kernel                                                 */
    K_CHOOSE = 2,    /* stage C1: choose_kernel                                                  */
    K_SERIALISE = 3, /* stage C2: encode_frames_kernel<.., true>                                 */
    K_SERIALISE2 = 4,/* stage C2: serialise_kernel<.., true> (flacb200_serialise.cuh)            */
    K_FIXED2 = 5     /* serialise_kernel<.., false>: fixed predictors only, one kernel           */
};

/* serialise_kernel variants: samples per thread, launch bounds (threads, frames per SM), registers that leaves */
constexpr int SER_VARIANTS = 5;
struct SerVariant { int S, max_threads, ctas, regs; };
inline SerVariant ser_variant(int v)
{
    switch (v) {
    case 0: return {16, 288, 3, 72};
    case 1: return {16, 1024, 1, 64};
    case 2: return {8, 576, 1, 112};
    case 3: return {32, 512, 1, 128};
    default: return {12, 384, 3, 56}; /* 12 warps: the same number on each of the SM's four schedulers */
    }
}

/* variant 0: <576, 1> launch bounds, 1: <288, 3> (72 registers), 2: <288, 4> (56 registers, K_CHOOSE with
 * 288 threads only); flavour: lpc_errors_rolled's, K_CHOOSE only */
#define FLACB_TU_DECL(name)                                                                                        \
    cudaError_t launch_##name(int kind, int variant, int flavour, unsigned grid, int threads, size_t smem,         \
                              cudaStream_t stream, const EncParams &p);                                            \
    cudaError_t init_##name(const uint16_t *crc_tab, const uint16_t *xpow_lo, const uint16_t *xpow_hi);            \
    size_t static_smem_##name();                                                                                   \
    size_t static_smem2_##name();
FLACB_TU_DECL(fixed)
FLACB_TU_DECL(12)
FLACB_TU_DECL(32)
#undef FLACB_TU_DECL
cudaError_t launch_mma_12(bool pdl, unsigned grid, int threads, size_t smem, cudaStream_t stream, const EncParams &p, const MmaLayout &ml);
cudaError_t launch_mma_32(bool pdl, unsigned grid, int threads, size_t smem, cudaStream_t stream, const EncParams &p, const MmaLayout &ml);

} /* namespace flacb */
