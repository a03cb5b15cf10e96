/*
 * flacb200_synth.cu -- BENCHMARK / TEST TOOLING, not part of the product library.
 *
 * The deterministic integer-only synthetic PCM of tools/synth/flacb200_synth.py, on the device:
 * every sample is a pure function of (seed, class, sample index, channel), so a B200 writes any
 * slice of a stream straight into HBM -- bit for bit what the numpy generator returns (the tests
 * compare them).  That is how the 41.5 GB corpus of BASELINE.json's config 4 and the hour of
 * config 3 exist at all: SURVEY.md section 8(d), "generate on device per shard".
 *
 * Each function below restates the numpy function of the same name; Python's floor division and
 * arithmetic shifts of negative numbers are reproduced, the sine table comes from the Python side
 * (numpy's sin, rounded there), so that no libm difference can enter.
 */
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef long long i64;

#define LUT_BITS 12
#define SEGMENT (1 << 16)

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ u64 rand_u64(u64 seed, u64 stream, u64 idx)
{
    const u64 base = seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull;
    return mix64(idx * 0x9E3779B97F4A7C15ull + base);
}

__device__ __forceinline__ i64 floordiv(i64 a, i64 b) /* b > 0 */
{
    i64 q = a / b;
    if (a % b != 0 && a < 0) q--;
    return q;
}

__device__ __forceinline__ i64 uniform(u64 seed, u64 stream, u64 idx, i64 lo, i64 hi)
{
    const i64 span = hi - lo + 1;
    const i64 r = (i64)(rand_u64(seed, stream, idx) >> 33); /* 31 bits */
    return ((r * span) >> 31) + lo;
}

__device__ __forceinline__ i64 gauss(u64 seed, u64 stream, u64 idx, i64 sigma)
{
    const u64 r = rand_u64(seed, stream, idx);
    i64 s = 0;
    for (int k = 0; k < 4; k++) s += (i64)((r >> (16 * k)) & 0xFFFFull);
    s -= 2 * 65535;
    return floordiv(s * sigma, 37837);
}

__device__ __forceinline__ i64 sine(const i64 *lut, u64 idx, u64 step, u64 phase0, i64 amp)
{
    const u64 ph = (idx * step + phase0) & 0xFFFFFFFFull;
    return (lut[ph >> (32 - LUT_BITS)] * amp) >> 30;
}

__device__ i64 class_sample(const i64 *lut, int cls, u64 idx, int ch, int bits, int rate, u64 seed)
{
    const i64 fs = ((i64)1 << (bits - 1)) - 1;
    const u64 st = 16ull * (u64)ch;
    switch (cls) {
    case 0: { /* sine */
        const int nparts = 3 + (int)((seed + (u64)ch) % 3ull);
        i64 out = 0;
        for (int p = 0; p < nparts; p++) {
            const u64 f = 110ull * (u64)(p + 1) * (u64)(ch + 2) + 37ull * (u64)p + seed % 13ull;
            const u64 step = (f << 32) / (u64)rate;
            out += sine(lut, idx, step, ((u64)p * 0x1234567ull + (u64)ch * 0x89ABCDEull) & 0xFFFFFFFFull, fs / (2 * nparts));
        }
        return out + uniform(seed, st + 1, idx, -3, 3);
    }
    case 1:
    case 2: { /* lowpass / highpass */
        const i64 sigma = bits <= 16 ? 1000 : ((i64)1 << 18);
        const i64 rho = cls == 1 ? 29491 : -29491; /* +-0.9 in Q15 */
        i64 out = 0, c = (i64)1 << 15;
        for (int k = 0; k < 16; k++) {
            const i64 e = idx >= (u64)k ? gauss(seed, st + 2, idx - (u64)k, sigma) : 0;
            out += (e * c) >> 15;
            c = (c * rho) >> 15;
        }
        return (out * 7) >> 4;
    }
    case 3: return gauss(seed, st + 3, idx, fs / 8); /* white */
    case 4: { /* silence: dither only, +-3 LSB at the 16-bit level; two-tap shaped for wider samples */
        if (bits <= 16) return uniform(seed, st + 4, idx, -3, 3);
        const i64 d = (i64)3 << (bits - 16);
        return uniform(seed, st + 4, idx, -d, d) + (idx >= 1 ? uniform(seed, st + 4, idx - 1, -d, d) : 0);
    }
    default: { /* clip */
        const u64 f = (bits <= 16 ? 1000ull : 1500ull) + 100ull * (u64)ch;
        const u64 step = (f << 32) / (u64)rate;
        i64 x = sine(lut, idx, step, (u64)ch * 0x3243F6Aull, 2 * fs) + gauss(seed, st + 5, idx, 3);
        x = x < -fs - 1 ? -fs - 1 : (x > fs ? fs : x);
        const i64 d = bits <= 16 ? 3 : 12;
        return x + uniform(seed, st + 6, idx, -d, d);
    }
    }
}

__global__ void synth_kernel(uint8_t *dst, u64 start, u64 n, int channels, int bits, int rate, int cls, u64 seed, const i64 *lut)
{
    const u64 total = n * (u64)channels;
    const int cb = (bits + 7) / 8;
    const i64 fs = ((i64)1 << (bits - 1)) - 1;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (u64)gridDim.x * blockDim.x) {
        const u64 i = t / (u64)channels;
        const int ch = (int)(t - i * (u64)channels);
        const u64 idx = start + i;
        const int k = cls == 6 ? (int)((idx / SEGMENT) % 6ull) : cls; /* mixed: the six classes in turn */
        i64 x = class_sample(lut, k, idx, ch, bits, rate, seed);
        x = x < -fs - 1 ? -fs - 1 : (x > fs ? fs : x);
        uint8_t *q = dst + t * (u64)cb;
        const uint32_t v = (uint32_t)(int32_t)x;
        for (int b = 0; b < cb; b++) q[b] = (uint8_t)(v >> (8 * b));
    }
}

static i64 *g_lut[16] = {nullptr};

/* the Q30 sine table of the Python side (2^LUT_BITS entries), once per device */
extern "C" int flacb_synth_set_lut(int device, const int64_t *lut)
{
    if (device < 0 || device >= 16) return 1;
    if (cudaSetDevice(device) != cudaSuccess) return 2;
    if (!g_lut[device] && cudaMalloc(&g_lut[device], sizeof(i64) << LUT_BITS) != cudaSuccess) return 3;
    return cudaMemcpy(g_lut[device], lut, sizeof(i64) << LUT_BITS, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 4;
}

/* sample frames [start, start + n) as packed little-endian PCM (ceil(bits / 8) bytes per sample,
 * interleaved) at dst, device memory of `device`; cls 0..5 = CLASSES, 6 = mixed */
extern "C" int flacb_synth_pcm(int device, void *dst, uint64_t start, uint64_t n, int channels, int bits, int rate, int cls,
                               uint64_t seed)
{
    if (device < 0 || device >= 16 || !g_lut[device]) return 1;
    if (cudaSetDevice(device) != cudaSuccess) return 2;
    if (n == 0) return 0;
    const u64 total = n * (u64)channels;
    u64 blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    synth_kernel<<<(unsigned)blocks, 256>>>((uint8_t *)dst, start, n, channels, bits, rate, cls, seed, g_lut[device]);
    if (cudaGetLastError() != cudaSuccess) return 5;
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 6;
}
