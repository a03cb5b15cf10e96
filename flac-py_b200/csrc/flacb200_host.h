/*
 * flacb200_host.h -- host-side logic shared by the CUDA library (flacb200.cu) and by the CPU
 * SIMT simulator harness of the tests (tests/sim): parameter validation with the reference's
 * error behaviour, the Tukey window, the CRC tables, the launch shape, the stream prologue.
 * Plain C++, no CUDA.
 */
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#define FLACB_HOST_ONLY_DECLS
#include "../../include/flacb200.h"

namespace flacb {

/* encoder.py:423-440 tukey(n, 0.5).  Evaluated on the host with the C library's cos(), the
 * same function Python's math.cos calls, so the table is identical to the reference's.
 * Returns false where the reference divides by zero (nr == 0, i.e. n in 4..7). */
inline bool tukey_window(int n, std::vector<double> &w, int pad)
{
    w.assign((size_t)n + (size_t)pad, 0.0);
    const int nr = (int)floor(0.5 / 2.0 * (double)n) - 1;
    for (int i = 0; i < n; i++) w[(size_t)i] = 1.0;
    if (nr == 0) return false;
    for (int i = 0; i <= nr; i++) {
        w[(size_t)i] = 0.5 - 0.5 * cos(M_PI * (double)i / (double)nr);
        w[(size_t)(n - nr - 1 + i)] = 0.5 - 0.5 * cos(M_PI * (double)(i + nr) / (double)nr);
    }
    return true;
}

/* longest run [lo, hi) around the middle of the window whose values are exactly 1.0 */
inline void window_flat_range(const std::vector<double> &w, int n, int &lo, int &hi)
{
    lo = hi = n / 2;
    if (n <= 0 || w[(size_t)(n / 2)] != 1.0) return;
    while (lo > 0 && w[(size_t)(lo - 1)] == 1.0) lo--;
    while (hi < n && w[(size_t)hi] == 1.0) hi++;
}

/* crc.py:4-14 crc_table(16, 0x18005) */
inline void crc16_table(uint16_t tab[256])
{
    for (int d = 0; d < 256; d++) {
        uint32_t v = (uint32_t)d << 8;
        for (int i = 0; i < 8; i++) {
            v <<= 1;
            if (v & 0x10000u) v ^= 0x18005u;
        }
        tab[d] = (uint16_t)v;
    }
}

/* x^(8*l) mod 0x18005 for l < 256 and x^(8*256*h) for h < n_hi: weights of the parallel CRC */
inline void crc16_xpow_tables(uint16_t lo[256], uint16_t *hi, int n_hi)
{
    auto mulx8 = [](uint32_t a) {
        for (int i = 0; i < 8; i++) {
            a <<= 1;
            if (a & 0x10000u) a ^= 0x18005u;
        }
        return a;
    };
    auto mul = [](uint32_t a, uint32_t b) {
        uint32_t r = 0;
        for (int i = 15; i >= 0; i--) {
            r <<= 1;
            if (r & 0x10000u) r ^= 0x18005u;
            if ((b >> i) & 1u) r ^= a;
        }
        return r;
    };
    uint32_t v = 1;
    for (int l = 0; l < 256; l++) {
        lo[l] = (uint16_t)v;
        v = mulx8(v);
    }
    const uint32_t step = v; /* x^(8*256) */
    uint32_t h = 1;
    for (int i = 0; i < n_hi; i++) {
        hi[i] = (uint16_t)h;
        h = mul(h, step);
    }
}

/* slice-by-4 tables of the word-wise CRC-16: t[j][b] = b * x^(8 j + 16) mod 0x18005, so that
 * four more message bytes w (big endian) take the CRC c to
 * t[3][v >> 24] ^ t[2][(v >> 16) & 255] ^ t[1][(v >> 8) & 255] ^ t[0][v & 255], v = (c << 16) ^ w */
inline void crc16_slice4(uint16_t t[4][256])
{
    crc16_table(t[0]);
    for (int j = 1; j < 4; j++)
        for (int b = 0; b < 256; b++) {
            const uint32_t c = t[j - 1][b];
            t[j][b] = (uint16_t)(((c << 8) ^ t[0][c >> 8]) & 0xFFFFu);
        }
}

/* w[i] = x^(8 i) mod 0x18005: weight of a slice CRC with i bytes behind it */
inline void crc16_weights(uint16_t *w, size_t n)
{
    uint32_t v = 1;
    for (size_t i = 0; i < n; i++) {
        w[i] = (uint16_t)v;
        for (int k = 0; k < 8; k++) {
            v <<= 1;
            if (v & 0x10000u) v ^= 0x18005u;
        }
    }
}

/* Rice bits < N*(bps+3.01); plus parameters, warm-up, coefficients, headers */
inline uint64_t frame_bound(int bs, int channels, int bps, int rice_max)
{
    uint64_t per = ((uint64_t)bs * (uint64_t)(bps + 4) + 7) / 8 + ((uint64_t)1 << rice_max) + 256;
    return per * (uint64_t)channels + 64;
}

struct EncodeArgs {
    uint64_t n_sample_frames;
    int channels, bits, block_size, max_lpc_order, qlp_precision, rice_min, rice_max;
    uint64_t first_frame_number;
    uint32_t flags;
};

/* Parameter checks with the status the reference would raise:
 *   EncoderParameters.__post_init__ asserts (encoder.py:40-43), block-size range
 *   (encoder.py:245-255 ValueError), coded number range (coded_number.py:38 ValueError). */
inline int validate_encode_args(const EncodeArgs &a)
{
    if (a.channels < 1 || a.channels > 8) return FLACB_E_ASSERT;
    if (a.bits < 4 || a.bits > 24) return a.bits >= 4 && a.bits <= 32 ? FLACB_E_UNSUPPORTED : FLACB_E_ASSERT;
    if (a.block_size < 1) return FLACB_E_ASSERT;
    if (a.block_size > 65535) return FLACB_E_VALUE; /* encode_block_size: no code for it */
    if (a.max_lpc_order < 0 || a.max_lpc_order > FLACB_MAX_LPC) return FLACB_E_ASSERT; /* stop <= 33 */
    if (a.qlp_precision < 5) return FLACB_E_ASSERT;                                   /* encoder.py:43 */
    if (a.qlp_precision > 16) return FLACB_E_UNSUPPORTED; /* the 4-bit field cannot hold it */
    if (a.rice_min < 0 || a.rice_max > 15 || a.rice_min > a.rice_max) return FLACB_E_ASSERT;
    const uint64_t n_frames = (a.n_sample_frames + (uint64_t)a.block_size - 1) / (uint64_t)a.block_size;
    if (n_frames && a.first_frame_number + n_frames - 1 >= (1ull << 31)) return FLACB_E_VALUE;
    if (n_frames > 0xFFFFFFF0ull) return FLACB_E_UNSUPPORTED;
    return FLACB_OK;
}

/* threads per frame: enough E-sample runs to cover the block in as few, as full tiles as
 * possible */
inline int pick_threads(int block_size, int e, int maxb)
{
    const int runs = (block_size + e - 1) / e;
    const int tiles = (runs + maxb - 1) / maxb;
    int b = (runs + tiles - 1) / tiles;
    b = (b + 31) / 32 * 32;
    if (b < 64) b = 64;
    if (b > maxb) b = maxb;
    return b;
}

/* Launch shape of the serialising kernel: how many threads per frame, whether every channel's
 * plane stays in shared memory, and how many frames share an SM.  The kernel is latency bound
 * (barriers between its phases), so resident frames per SM come first: up to 3 with the
 * <288, 3> instantiation (72 registers), then resident planes (nothing reloaded or recomputed
 * when the frame is serialised). */
struct LaunchShape {
    int threads, all_planes, ctas_per_sm, variant; /* variant 0: <576,1>, 1: <288,3> */
};

template <class LayoutFn>
inline LaunchShape pick_shape(int block_size, int e, size_t static_smem, size_t smem_optin, size_t smem_per_sm,
                              LayoutFn layout_total)
{
    LaunchShape best = {0, 0, 0, 0};
    for (int variant = 1; variant >= 0; variant--) {
        const int maxb = variant ? 288 : 576;
        const int regs = variant ? 72 : 96;
        for (int all_planes = 1; all_planes >= 0; all_planes--) {
            const int threads = pick_threads(block_size, e, maxb);
            const size_t dyn = layout_total(threads, all_planes);
            if (dyn + static_smem > smem_optin) continue;
            int by_smem = (int)(smem_per_sm / (dyn + static_smem + 1024));
            int by_regs = 65536 / (regs * threads);
            int ctas = by_smem < by_regs ? by_smem : by_regs;
            if (variant && ctas > 3) ctas = 3;
            if (!variant && ctas > 2) ctas = 2;
            if (ctas < 1) continue;
            const int score = ctas * threads * 4 + all_planes;
            const int best_score = best.ctas_per_sm * best.threads * 4 + best.all_planes;
            if (best.threads == 0 || score > best_score) best = {threads, all_planes, ctas, variant};
        }
    }
    return best;
}

/* Host path with host buffers on both sides: first frame of every chunk, then n_frames.  Chunks are
 * whole waves of resident frames (`wave`), about 16 MB of PCM (two waves when that is still a
 * small chunk: one-wave kernels leave nothing to fill their tails with); the first and the last
 * chunk are one wave, because nothing overlaps the first chunk's input copy or the last chunk's
 * kernels and output copy.  want_override > 0 (FLACB200_CHUNK_FRAMES): that many frames per
 * chunk, taken as it is; edge_override > 0 (FLACB200_EDGE_CHUNK_FRAMES): size of the first and
 * last chunk. */
inline std::vector<uint64_t> chunk_plan(uint64_t n_frames, uint64_t wave, uint64_t frame_pcm, long long want_override,
                                        long long edge_override, long long last_override = 0)
{
    std::vector<uint64_t> f0;
    f0.push_back(0);
    if (wave < 1) wave = 1;
    uint64_t want = (16ull << 20) / (frame_pcm ? frame_pcm : 1);
    want = (want + wave / 2) / wave * wave; /* whole waves of resident frames, nearest */
    if (want < wave) want = wave;
    if (want < 2 * wave && 2 * wave * frame_pcm <= (32ull << 20)) want = 2 * wave;
    uint64_t edge = wave;
    if (want_override > 0) {
        want = (uint64_t)want_override;
        if (edge > want) edge = want;
    }
    if (want * 2 <= n_frames) {
        uint64_t first = edge, last = edge;
        if (edge_override > 0) first = last = (uint64_t)edge_override;
        if (last_override > 0) last = (uint64_t)last_override;
        if (first + last >= n_frames) first = last = want;
        uint64_t f = first < want ? first : want;
        while (f + last < n_frames) {
            f0.push_back(f);
            f += want;
        }
        if (n_frames - f0.back() > last && last < want) f0.push_back(n_frames - last);
    }
    f0.push_back(n_frames);
    return f0;
}

/* encoder.py:61-83, :170-189: magic, metadata block header (last, STREAMINFO, 34), STREAMINFO
 * with min = max block size, frame sizes 0, and a zero MD5 */
inline void stream_prologue(uint32_t sample_rate, int channels, int bits, uint64_t total_samples,
                            int block_size, uint8_t out[42])
{
    memset(out, 0, 42);
    memcpy(out, "fLaC", 4);
    out[4] = 0x80;
    out[7] = 34;
    uint8_t *s = out + 8;
    s[0] = (uint8_t)(block_size >> 8);
    s[1] = (uint8_t)block_size;
    s[2] = s[0];
    s[3] = s[1];
    /* s[4..9]: min/max frame size = 0 */
    const uint64_t v = ((uint64_t)(sample_rate & 0xFFFFFu) << 44) | ((uint64_t)((channels - 1) & 7) << 41) |
                       ((uint64_t)((bits - 1) & 31) << 36) | (total_samples & 0xFFFFFFFFFull);
    for (int i = 0; i < 8; i++) s[10 + i] = (uint8_t)(v >> (56 - 8 * i));
}

inline const char *status_text(int status)
{
    switch (status) {
    case FLACB_OK: return "ok";
    case FLACB_E_ASSERT: return "AssertionError: the reference asserts on this input or parameter";
    case FLACB_E_VALUE: return "ValueError: the reference raises ValueError on this input";
    case FLACB_E_ZERODIV: return "ZeroDivisionError: the reference divides by zero on this input";
    case FLACB_E_EOF: return "EOFError: truncated stream";
    case FLACB_E_NOSPACE: return "output buffer too small";
    case FLACB_E_OVERFLOW: return "OverflowError: floating point overflow in the LPC analysis";
    case FLACB_E_KEY: return "KeyError: reserved header code";
    case FLACB_E_NOTIMPL: return "NotImplementedError";
    case FLACB_E_CUDA: return "CUDA runtime error";
    case FLACB_E_UNSUPPORTED: return "parameter combination not supported by the B200 kernels";
    case FLACB_E_RANGE: return "residual does not fit 32 bits";
    case FLACB_E_INTERNAL: return "internal error";
    default: return "unknown status";
    }
}

} /* namespace flacb */
