/*
 * flacb200_tu.cuh -- body of the kernel translation units: instantiates the encode kernels for
 * FLACB_TU_LMAX (unroll depth of the LPC loops; 0: the fixed-predictor kernel) and defines their
 * launchers (flacb200_launch.h).
 */
#define FLACB_TU_KERNELS_ONLY 1
#include <string.h>
#include "flacb200_launch.h"
#include "flacb200_serialise.cuh"
#include "flacb200_choose_mma.cuh"

namespace flacb {

#define FLACB_CAT_(a, b) a##b
#define FLACB_CAT(a, b) FLACB_CAT_(a, b)
#if FLACB_TU_LMAX
#define FLACB_TU_NAME FLACB_TU_LMAX
#else
#define FLACB_TU_NAME fixed
#endif

/* flavour bit 8: launch with programmatic stream serialisation (the kernel may start while the kernel
 * in front of it on the stream is still finishing; it calls pdl_wait() before it reads that kernel's results) */
static bool g_pdl = false;

template <class K>
static cudaError_t launch_one(K kernel, unsigned grid, int threads, size_t smem, cudaStream_t stream, const EncParams &p)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (g_pdl) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3((unsigned)threads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, p);
    }
    kernel<<<grid, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

#if FLACB_TU_LMAX
/* stage C1 on the tensor cores (flacb200_choose_mma.cuh) */
cudaError_t FLACB_CAT(launch_mma_, FLACB_TU_NAME)(bool pdl, unsigned grid, int threads, size_t smem, cudaStream_t stream,
                                                   const EncParams &p, const MmaLayout &ml)
{
    auto kernel = choose_mma_kernel<FLACB_TU_LMAX, 256, 3>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p, ml);
}
#endif

cudaError_t FLACB_CAT(launch_, FLACB_TU_NAME)(int kind, int variant, int flavour, unsigned grid, int threads, size_t smem,
                                               cudaStream_t stream, const EncParams &p)
{
    g_pdl = (flavour & 0x100) != 0;
    flavour &= 0xFF;
#if FLACB_TU_LMAX
    constexpr int LM = FLACB_TU_LMAX;
    switch (kind) {
    case K_ANALYSE:
        if (variant && threads == 288) return launch_one(analyse_kernel<LM, 288, 3, 288>, grid, threads, smem, stream, p);
#if FLACB_TU_LMAX == 32 /* block 16384: 512-thread frames */
        if (!variant && threads == 512) return launch_one(analyse_kernel<LM, 576, 1, 512>, grid, threads, smem, stream, p);
#endif
        return variant ? launch_one(analyse_kernel<LM, 288, 3, 0>, grid, threads, smem, stream, p)
                       : launch_one(analyse_kernel<LM, 576, 1, 0>, grid, threads, smem, stream, p);
    case K_CHOOSE:
        switch (flavour) {
        case 0:
            if (variant == 2 && threads == 288) return launch_one(choose_kernel<LM, 288, 4, 0, 288>, grid, threads, smem, stream, p);
            if (variant && threads == 288) return launch_one(choose_kernel<LM, 288, 3, 0, 288>, grid, threads, smem, stream, p);
            return variant ? launch_one(choose_kernel<LM, 288, 3, 0, 0>, grid, threads, smem, stream, p)
                           : launch_one(choose_kernel<LM, 576, 1, 0, 0>, grid, threads, smem, stream, p);
        case 1:
            if (variant == 2 && threads == 288) return launch_one(choose_kernel<LM, 288, 4, 1, 288>, grid, threads, smem, stream, p);
            if (variant && threads == 288) return launch_one(choose_kernel<LM, 288, 3, 1, 288>, grid, threads, smem, stream, p);
            return variant ? launch_one(choose_kernel<LM, 288, 3, 1, 0>, grid, threads, smem, stream, p)
                           : launch_one(choose_kernel<LM, 576, 1, 1, 0>, grid, threads, smem, stream, p);
        default:
            if (variant == 2 && threads == 288) return launch_one(choose_kernel<LM, 288, 4, 2, 288>, grid, threads, smem, stream, p);
            if (variant && threads == 288) return launch_one(choose_kernel<LM, 288, 3, 2, 288>, grid, threads, smem, stream, p);
#if FLACB_TU_LMAX == 32 /* block 16384, 24-bit: 256-thread frames */
            if (variant && threads == 256) return launch_one(choose_kernel<LM, 288, 3, 2, 256>, grid, threads, smem, stream, p);
#endif
            return variant ? launch_one(choose_kernel<LM, 288, 3, 2, 0>, grid, threads, smem, stream, p)
                           : launch_one(choose_kernel<LM, 576, 1, 2, 0>, grid, threads, smem, stream, p);
        }
    case K_SERIALISE:
        if (variant && threads == 288) return launch_one(encode_frames_kernel<LM, 288, 3, true, 288>, grid, threads, smem, stream, p);
#if FLACB_TU_LMAX == 32
        if (!variant && threads == 512) return launch_one(encode_frames_kernel<LM, 576, 1, true, 512>, grid, threads, smem, stream, p);
#endif
        return variant ? launch_one(encode_frames_kernel<LM, 288, 3, true, 0>, grid, threads, smem, stream, p)
                       : launch_one(encode_frames_kernel<LM, 576, 1, true, 0>, grid, threads, smem, stream, p);
    case K_SERIALISE2: /* serialise_kernel<LMAX, chosen, samples per thread, launch bounds> */
        switch (variant) {
        case 0: return launch_one(serialise_kernel<LM, true, 16, 288, 3>, grid, threads, smem, stream, p);
        case 1: return launch_one(serialise_kernel<LM, true, 16, 1024, 1>, grid, threads, smem, stream, p);
        case 2: return launch_one(serialise_kernel<LM, true, 8, 576, 1>, grid, threads, smem, stream, p);
        case 3: return launch_one(serialise_kernel<LM, true, 32, 512, 1>, grid, threads, smem, stream, p);
        case 4: return launch_one(serialise_kernel<LM, true, 12, 384, 3>, grid, threads, smem, stream, p);
        default: return cudaErrorInvalidValue;
        }
    default: return cudaErrorInvalidValue;
    }
#else
    if (kind == K_FIXED2) {
        switch (variant) {
        case 0: return launch_one(serialise_kernel<4, false, 16, 288, 3>, grid, threads, smem, stream, p);
        case 1: return launch_one(serialise_kernel<4, false, 16, 1024, 1>, grid, threads, smem, stream, p);
        case 2: return launch_one(serialise_kernel<4, false, 8, 576, 1>, grid, threads, smem, stream, p);
        case 3: return launch_one(serialise_kernel<4, false, 32, 512, 1>, grid, threads, smem, stream, p);
        case 4: return launch_one(serialise_kernel<4, false, 12, 384, 3>, grid, threads, smem, stream, p);
        default: return cudaErrorInvalidValue;
        }
    }
    if (kind != K_FIXED) return cudaErrorInvalidValue;
    if (variant && threads == 288) return launch_one(encode_frames_kernel<12, 288, 3, false, 288>, grid, threads, smem, stream, p);
    return variant ? launch_one(encode_frames_kernel<12, 288, 3, false, 0>, grid, threads, smem, stream, p)
                   : launch_one(encode_frames_kernel<12, 576, 1, false, 0>, grid, threads, smem, stream, p);
#endif
}

/* the __constant__ tables of flacb200_kernels.cuh exist once per translation unit */
cudaError_t FLACB_CAT(init_, FLACB_TU_NAME)(const uint16_t *crc_tab, const uint16_t *xpow_lo, const uint16_t *xpow_hi)
{
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_crc16_tab, crc_tab, sizeof(uint16_t) * 256)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_xpow_lo, xpow_lo, sizeof(uint16_t) * 256)) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_xpow_hi, xpow_hi, sizeof(uint16_t) * XPOW_HI);
}

size_t FLACB_CAT(static_smem_, FLACB_TU_NAME)()
{
    cudaFuncAttributes fa;
#if FLACB_TU_LMAX
    if (cudaFuncGetAttributes(&fa, encode_frames_kernel<FLACB_TU_LMAX, 576, 1, true, 0>) != cudaSuccess) return 8192;
#else
    if (cudaFuncGetAttributes(&fa, encode_frames_kernel<12, 576, 1, false, 0>) != cudaSuccess) return 8192;
#endif
    return fa.sharedSizeBytes;
}

size_t FLACB_CAT(static_smem2_, FLACB_TU_NAME)()
{
    cudaFuncAttributes fa;
#if FLACB_TU_LMAX
    if (cudaFuncGetAttributes(&fa, serialise_kernel<FLACB_TU_LMAX, true, 16, 288, 3>) != cudaSuccess) return 12288;
#else
    if (cudaFuncGetAttributes(&fa, serialise_kernel<4, false, 16, 288, 3>) != cudaSuccess) return 12288;
#endif
    return fa.sharedSizeBytes;
}

} /* namespace flacb */
