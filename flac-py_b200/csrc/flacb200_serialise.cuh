/*
 * flacb200_serialise.cuh -- second generation of the serialising stage (C2) and of the
 * fixed-predictor-only encode: serialise_kernel.
 *
 * Same job as encode_frames_kernel (flacb200_kernels.cuh): winner's residual (reference
 * flac/encoder.py:537-548), zig-zag (utils.py:91-94), Rice partition search
 * (encoder.py:632-760), exact subframe sizes, frame placement by decoupled look-back, frame
 * header + CRC-8 (encoder.py:194-234), subframes (encoder.py:553-627), Rice codes
 * (encoder.py:765-806) packed MSB first (binary.py:144-216), padding and CRC-16
 * (encoder.py:159-165, crc.py).  What changed is the work split, because the first kernel spent
 * its time waiting at block barriers (4.6 warps stalled on a barrier per issued instruction,
 * about 65 barriers per stereo frame):
 *
 *   - a thread owns S CONSECUTIVE samples of every channel (the first kernel: runs of 8,
 *     tile-strided).  One block scan per channel gives both the partition sums of every
 *     candidate order and, later, the bit position of every code: no per-tile scans, no
 *     per-tile barriers.
 *   - the residual stays in the thread's registers from its computation to the sums over
 *     (u >> k); it is written back in place of the samples only for the packing phase, where
 *     every thread reads its own words again.  Candidate orders whose partition of a span has
 *     the same parameter share the sum.
 *   - decisions that used to be "thread 0 decides, barrier, everybody reads" are taken by every
 *     warp redundantly from per-warp partial sums: one barrier per reduction instead of two.
 *   - the serial fields (subframe header, warm-up samples, LPC header and coefficients) are
 *     or-ed in by different threads at closed-form bit positions while the others pack codes.
 *   - the whole frame is staged in shared memory when it fits and leaves with one flush;
 *     CRC-16 runs word-wise (slice-by-4 tables in shared memory) over balanced slices, weighted
 *     with x^(8 * bytes behind) from a direct table (one table look-up and one carry-less
 *     multiply per thread instead of two multiplies per 64-byte slice), while warp 0 looks back
 *     for the frame's offset.
 *
 * 14 barriers per stereo frame.  Channel planes are XOR-swizzled on 16-byte units so that a
 * quarter warp's 16-byte accesses at strides of 32 / 64 / 128 bytes hit every bank once.
 *
 * Measured and NOT kept (profiles/README.md, round 2): persistent CTAs with a prefetched ticket,
 * Rice parameters per warp by shuffles instead of per block behind a barrier, one scan barrier
 * for all channels, the previous frame's copy-out deferred into the next frame's round (it does
 * remove the look-back waits, 4-10 retries per frame -> 0.1, but the SM's other frames had been
 * using those cycles anyway): each of them made the kernel 10-15 % slower or changed nothing.
 *
 * Limits (the host falls back to encode_frames_kernel outside them): every plane resident,
 * block <= 1024 * S samples, Rice partition orders <= 8, a staging window that holds at least
 * one worst-case subframe, no slot mode.
 */
#pragma once

#include "flacb200_kernels.cuh"

namespace flacb {

#ifndef FLACB_HOST_ONLY


/* crc.py:17-22 crc8, polynomial 0x107 (common.py:12): table of the byte-wise form */
static __constant__ uint8_t c_crc8_tab[256] = {
    0x00,0x07,0x0E,0x09,0x1C,0x1B,0x12,0x15,0x38,0x3F,0x36,0x31,0x24,0x23,0x2A,0x2D,0x70,0x77,0x7E,0x79,0x6C,0x6B,0x62,0x65,0x48,0x4F,0x46,0x41,0x54,0x53,0x5A,0x5D,
    0xE0,0xE7,0xEE,0xE9,0xFC,0xFB,0xF2,0xF5,0xD8,0xDF,0xD6,0xD1,0xC4,0xC3,0xCA,0xCD,0x90,0x97,0x9E,0x99,0x8C,0x8B,0x82,0x85,0xA8,0xAF,0xA6,0xA1,0xB4,0xB3,0xBA,0xBD,
    0xC7,0xC0,0xC9,0xCE,0xDB,0xDC,0xD5,0xD2,0xFF,0xF8,0xF1,0xF6,0xE3,0xE4,0xED,0xEA,0xB7,0xB0,0xB9,0xBE,0xAB,0xAC,0xA5,0xA2,0x8F,0x88,0x81,0x86,0x93,0x94,0x9D,0x9A,
    0x27,0x20,0x29,0x2E,0x3B,0x3C,0x35,0x32,0x1F,0x18,0x11,0x16,0x03,0x04,0x0D,0x0A,0x57,0x50,0x59,0x5E,0x4B,0x4C,0x45,0x42,0x6F,0x68,0x61,0x66,0x73,0x74,0x7D,0x7A,
    0x89,0x8E,0x87,0x80,0x95,0x92,0x9B,0x9C,0xB1,0xB6,0xBF,0xB8,0xAD,0xAA,0xA3,0xA4,0xF9,0xFE,0xF7,0xF0,0xE5,0xE2,0xEB,0xEC,0xC1,0xC6,0xCF,0xC8,0xDD,0xDA,0xD3,0xD4,
    0x69,0x6E,0x67,0x60,0x75,0x72,0x7B,0x7C,0x51,0x56,0x5F,0x58,0x4D,0x4A,0x43,0x44,0x19,0x1E,0x17,0x10,0x05,0x02,0x0B,0x0C,0x21,0x26,0x2F,0x28,0x3D,0x3A,0x33,0x34,
    0x4E,0x49,0x40,0x47,0x52,0x55,0x5C,0x5B,0x76,0x71,0x78,0x7F,0x6A,0x6D,0x64,0x63,0x3E,0x39,0x30,0x37,0x22,0x25,0x2C,0x2B,0x06,0x01,0x08,0x0F,0x1A,0x1D,0x14,0x13,
    0xAE,0xA9,0xA0,0xA7,0xB2,0xB5,0xBC,0xBB,0x96,0x91,0x98,0x9F,0x8A,0x8D,0x84,0x83,0xDE,0xD9,0xD0,0xD7,0xC2,0xC5,0xCC,0xCB,0xE6,0xE1,0xE8,0xEF,0xFA,0xFD,0xF4,0xF3};

/* encoder.py:194-234 put_frame_header (+ coded_number.py:7-39, crc.py:17-22), table CRC-8 */
static __device__ __noinline__ int build_frame_header_v2(uint32_t flags, int channels, int bs, u64 number, uint8_t *h)
{
    int code = 0, extra = 0;
    switch (bs) { /* common.py:85-105 */
    case 192: code = 1; break;
    case 576: code = 2; break;
    case 1152: code = 3; break;
    case 2304: code = 4; break;
    case 4608: code = 5; break;
    case 256: code = 8; break;
    case 512: code = 9; break;
    case 1024: code = 10; break;
    case 2048: code = 11; break;
    case 4096: code = 12; break;
    case 8192: code = 13; break;
    case 16384: code = 14; break;
    case 32768: code = 15; break;
    default:
        if (bs <= 255) code = 6, extra = 8; /* encoder.py:245-255 */
        else code = 7, extra = 16;
    }
    int len = 0;
    h[len++] = 0xFF; /* sync 0x7FFC (15 bits) + blocking strategy 0 */
    h[len++] = 0xF8;
    h[len++] = (uint8_t)(code << 4); /* sample rate 0000: from STREAMINFO (encoder.py:94) */
    const uint32_t chn = (flags & FLACB_SPEC_CHANNELS) ? (uint32_t)(channels - 1) : 1u; /* encoder.py:95 */
    h[len++] = (uint8_t)(chn << 4); /* sample size 000 + reserved 0 (encoder.py:96) */
    const int nb = 64 - __clzll(number | 1ull) - ((number == 0) ? 1 : 0);
    if (nb <= 7) h[len++] = (uint8_t)number;
    else {
        const int size = nb <= 11 ? 2 : nb <= 16 ? 3 : nb <= 21 ? 4 : nb <= 26 ? 5 : 6;
        for (int i = 0; i < size; i++) {
            const uint8_t six = (uint8_t)((number >> (6 * (size - 1 - i))) & 0x3F);
            h[len++] = i == 0 ? (uint8_t)((((1u << size) - 1u) << (8 - size)) | six) : (uint8_t)(0x80 | six);
        }
    }
    if (extra == 8) h[len++] = (uint8_t)(bs - 1);
    else if (extra == 16) h[len++] = (uint8_t)((bs - 1) >> 8), h[len++] = (uint8_t)(bs - 1);
    uint32_t crc = 0;
    for (int i = 0; i < len; i++) crc = c_crc8_tab[crc ^ h[i]];
    h[len++] = (uint8_t)crc;
    return len;
}

struct SerShared {
    Decision dec[MAX_CH];
    int chan_status[MAX_CH];
    int32_t warm[MAX_CH][FLACB_MAX_LPC];
    uint32_t chan_bit0[MAX_CH + 1];
    uint32_t dec_magic[MAX_CH];         /* ceil(2^32 / partition size) of the chosen order */
    uint32_t sizeA[2][16], sizeB[2][16], any5[2][16], errkey[2], magic_f[2];
    u64 wtot[32];
    uint32_t wtot2[2][32];
    u64 partf[32][5];
    uint32_t part[2][32][12];
    uint32_t frame, frame_bytes, payload_bytes, wbase, crc;
    u64 frame_off;
    int status, writable, have_off;
    int range_flag[MAX_CH];
    int hdr_len;
    uint8_t hdr[16];
    uint16_t crc4[4][256];
    /* where the frame goes */
    uint8_t *o_out;
    u64 o_cap;
    u64 *o_lb_state, *o_frame_off;
    uint32_t o_n_frames;
};

struct SerCtx {
    const EncParams &p;
    SerShared &sh;
    int tid, B, lane, warp, nwarps, n;
    int32_t *planes;
    uint32_t pstride;
    uint8_t *kall;
    uint32_t kall_stride;
    u64 *cumw;
    uint32_t *stg;
    uint32_t stg_words;
    __device__ SerCtx(const EncParams &p_, SerShared &sh_) : p(p_), sh(sh_) {}
};

template <int N>
__device__ __forceinline__ void ld2(const int32_t *x, uint32_t i, int (&v)[N])
{
#pragma unroll
    for (int k = 0; k < N; k += 4) {
        const int4 t = *(const int4 *)(x + swz2(i + (uint32_t)k));
        v[k] = t.x, v[k + 1] = t.y, v[k + 2] = t.z, v[k + 3] = t.w;
    }
}

template <int N>
__device__ __forceinline__ void st2(int32_t *x, uint32_t i, const uint32_t (&v)[N])
{
#pragma unroll
    for (int k = 0; k < N; k += 4) {
        int4 t;
        t.x = (int)v[k], t.y = (int)v[k + 1], t.z = (int)v[k + 2], t.w = (int)v[k + 3];
        *(int4 *)(x + swz2(i + (uint32_t)k)) = t;
    }
}

/* ------------------------------------------------------------------------------------- */
/* frame/channel split (encoder.py:87 batch, :102 de-interleave): a thread takes four sample
 * frames (4 * CH samples, CH * CB 32-bit words), and stores four consecutive samples of every
 * channel with one 16-byte store.                                                          */

template <int CH, int CB>
__device__ __forceinline__ void load_groups(const uint8_t *base, int32_t *planes, uint32_t pstride, int groups, int tid, int B)
{
    constexpr int W = CH * CB; /* 32-bit words per group */
    for (int g = tid; g < groups; g += B) {
        uint32_t w[W + 1];
        const uint32_t *src = (const uint32_t *)base + (size_t)g * W;
        if (W % 4 == 0 && (((uintptr_t)base) & 15) == 0) {
#pragma unroll
            for (int k = 0; k < W; k += 4) {
                const uint4 t = *(const uint4 *)(src + k);
                w[k] = t.x, w[k + 1] = t.y, w[k + 2] = t.z, w[k + 3] = t.w;
            }
        } else if (W % 2 == 0 && (((uintptr_t)base) & 7) == 0) {
#pragma unroll
            for (int k = 0; k < W; k += 2) {
                const uint2 t = *(const uint2 *)(src + k);
                w[k] = t.x, w[k + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) w[k] = src[k];
        }
        w[W] = 0;
#pragma unroll
        for (int c = 0; c < CH; c++) {
            int4 v;
            int vv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int s = j * CH + c; /* sample of the group */
                if (CB == 2) {
                    const uint32_t ww = w[s >> 1];
                    vv[j] = (s & 1) ? (int32_t)ww >> 16 : (int32_t)(int16_t)(ww & 0xFFFFu);
                } else { /* CB == 3 */
                    const int bo = 3 * s, wi = bo >> 2, shb = (bo & 3) * 8;
                    const uint32_t raw = shb == 0 ? w[wi] : (shb == 8 ? w[wi] >> 8 : __funnelshift_r(w[wi], w[wi + 1], (uint32_t)shb));
                    vv[j] = (int32_t)(raw << 8) >> 8;
                }
            }
            v.x = vv[0], v.y = vv[1], v.z = vv[2], v.w = vv[3];
            *(int4 *)(planes + (uint32_t)c * pstride + swz2((uint32_t)(V2_GUARD + 4 * g))) = v;
        }
    }
}

static __device__ __noinline__ void load_planes_v2(const uint8_t *base, int32_t *planes, uint32_t pstride, uint32_t xwords,
                                                    int n, int channels, int cbytes, int tid, int B)
{
    /* zero history in front of and zero samples behind every plane */
    const int pad = (int)xwords - V2_GUARD - n;
    for (int idx = tid; idx < channels * (V2_GUARD + pad); idx += B) {
        const int pl = idx / (V2_GUARD + pad), j = idx - pl * (V2_GUARD + pad);
        const int i = j < V2_GUARD ? j : n + j;
        planes[(uint32_t)pl * pstride + swz2((uint32_t)i)] = 0;
    }
    int done = 0; /* sample frames [0, done) went through a vector path */
    if ((((uintptr_t)base) & 3) == 0) {
        const int groups = n >> 2;
        bool vec = true;
        if (channels == 1 && cbytes == 2) load_groups<1, 2>(base, planes, pstride, groups, tid, B);
        else if (channels == 2 && cbytes == 2) load_groups<2, 2>(base, planes, pstride, groups, tid, B);
        else if (channels == 1 && cbytes == 3) load_groups<1, 3>(base, planes, pstride, groups, tid, B);
        else if (channels == 2 && cbytes == 3) load_groups<2, 3>(base, planes, pstride, groups, tid, B);
        else if (channels == 8 && cbytes == 3) load_groups<8, 3>(base, planes, pstride, groups, tid, B);
        else if (channels == 8 && cbytes == 2) load_groups<8, 2>(base, planes, pstride, groups, tid, B);
        else vec = false;
        if (vec) done = groups << 2;
    }
    const bool aligned = (((uintptr_t)base) & (uintptr_t)(cbytes - 1)) == 0 || cbytes == 3 || cbytes == 1;
    const int total = (n - done) * channels;
    for (int r = tid; r < total; r += B) {
        const int sidx = done * channels + r;
        const int i = sidx / channels, c = sidx - i * channels;
        planes[(uint32_t)c * pstride + swz2((uint32_t)(V2_GUARD + i))] = load_sample(base, (u64)sidx, cbytes, 0, aligned);
    }
}

/* ------------------------------------------------------------------------------------- */
/* frame offset by decoupled look-back (same protocol as resolve_offset of the first kernel) */

static __device__ __noinline__ void look_back_warp(SerShared *shp, int lane)
{
    SerShared &sh = *shp;
    const uint32_t f = sh.frame;
    u64 *lb_state = sh.o_lb_state;
    {
        const u64 mine = (u64)sh.frame_bytes;
        u64 excl = 0;
        bool failed = false;
        if (f > 0) {
            int64_t j = (int64_t)f - 1;
            u64 spins = 0;
            while (true) {
                const int64_t idx = j - lane;
                u64 s = LB_PREFIX; /* before frame 0: an empty prefix */
                if (idx >= 0) s = *(volatile u64 *)&lb_state[idx];
                const uint32_t flag = (uint32_t)(s >> 62);
                const unsigned has_prefix = __ballot_sync(FULL, flag == 2);
                const unsigned invalid = __ballot_sync(FULL, flag == 0);
                const int first_prefix = has_prefix ? __ffs((int)has_prefix) - 1 : 32;
                const unsigned need = first_prefix == 32 ? FULL : ((2u << first_prefix) - 1u);
                if (invalid & need) { /* a predecessor has not published yet: it is resident (ticket order), so it will */
                    if (++spins > (1ull << 31)) { failed = true; break; } /* minutes: only a dead predecessor gets here */
                    __nanosleep(spins < 64 ? 32 : 256);
                    continue;
                }
                u64 v = (lane <= first_prefix) ? (s & LB_VALUE) : 0ull;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
                excl += v;
                if (first_prefix < 32) break;
                j -= 32;
            }
            if (lane == 0) atomicExch(&lb_state[f], LB_PREFIX | ((excl + mine) & LB_VALUE));
        }
        if (lane == 0) {
            sh.frame_off = excl;
            sh.o_frame_off[f] = excl;
            if (f == sh.o_n_frames - 1) sh.o_frame_off[sh.o_n_frames] = excl + mine;
            if (failed && sh.status == 0) sh.status = FLACB_E_INTERNAL;
            const u64 out_cap = sh.o_cap;
            const u64 lim = out_cap - (excl < out_cap ? excl : out_cap);
            sh.writable = (u64)sh.frame_bytes <= lim && sh.status != FLACB_E_INTERNAL && sh.status != FLACB_E_NOSPACE;
            if (!sh.writable && sh.status == 0) sh.status = FLACB_E_NOSPACE;
            sh.have_off = 1;
        }
    }
}

static __device__ __noinline__ void resolve_offset_v2(SerShared *shp, int tid)
{
    if (tid < 32) look_back_warp(shp, tid);
    __syncthreads();
}

/* Flush the first `nbytes` bytes of the staging window to the frame in HBM and fold them into the
 * frame CRC-16 (crc.py:25-30).  The CRC has zero init and no final xor, so crc(A || B) =
 * crc(A) * x^(8|B|) + crc(B): every thread takes a slice of `sw` words, and its slice CRC times
 * x^(8 * bytes behind the slice) is xor-ed into the frame's.  last: nothing is packed afterwards;
 * otherwise the window is rebased at word `wend` (the partial word moves to the front).
 * Starts with a barrier: everything packed so far must be in the window. */
static __device__ __noinline__ void flush_v2(SerShared *shp, uint32_t *stg, uint32_t stg_words, const uint16_t *crc_w, int tid, int B,
                                             uint32_t nbytes, uint32_t wend, int last)
{
    SerShared &sh = *shp;
    const int have_off = sh.have_off; /* read in front of the barrier: look_back_warp sets it behind it */
    __syncthreads();
    const uint32_t byte0 = sh.wbase * 4u;          /* frame-relative */
    auto copy_out = [&]() {
        if (!(sh.writable && nbytes)) return;
        /* staging words hold the stream MSB first; memory wants those bytes in order.  Aligned
         * 32-bit stores for the body, single bytes in front of / behind it. */
        uint8_t *dst = sh.o_out + sh.frame_off + byte0;
        uint32_t head = (4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u;
        if (head > nbytes) head = nbytes;
        const uint32_t nq = (nbytes - head) >> 2;
        const uint32_t tail0 = head + nq * 4u;
#pragma unroll 1
        for (uint32_t idx = (uint32_t)tid; idx < head + (nbytes - tail0); idx += (uint32_t)B) {
            const uint32_t i = idx < head ? idx : tail0 + (idx - head);
            dst[i] = (uint8_t)(stg[i >> 2] >> (24u - 8u * (i & 3u)));
        }
        uint32_t *dstw = (uint32_t *)(dst + head);
        if (head == 0) {
#pragma unroll 2
            for (uint32_t q = (uint32_t)tid; q < nq; q += (uint32_t)B) dstw[q] = __byte_perm(stg[q], 0, 0x0123);
        } else {
            const uint32_t sh8 = 8u * head;
#pragma unroll 2
            for (uint32_t q = (uint32_t)tid; q < nq; q += (uint32_t)B) {
                /* bytes head + 4q .. + 3 of the window: the low (4 - head) bytes of word q, the high `head` bytes of word q + 1 */
                const uint32_t big = (stg[q] << sh8) | (stg[q + 1] >> (32u - sh8));
                dstw[q] = __byte_perm(big, 0, 0x0123);
            }
        }
    };
    /* CRC-16, slice-by-4: c' = T3[v >> 24] ^ T2[v >> 16 & 255] ^ T1[v >> 8 & 255] ^ T0[v & 255], v = (c << 16) ^ word;
     * worker t of nt */
    auto crc_slices = [&](int t, int nt) {
        const uint32_t nwords = nbytes >> 2, rem = nbytes & 3u;
        const uint32_t sw = (nwords + (uint32_t)nt - 1) / (uint32_t)nt;
        const uint32_t w0 = (uint32_t)t * sw;
        const bool tail_owner = rem && (nwords == 0 ? t == 0 : (uint32_t)t == (nwords - 1) / sw);
        if (w0 >= nwords && !tail_owner) return;
        const uint32_t w1 = w0 >= nwords ? w0 : (w0 + sw < nwords ? w0 + sw : nwords);
        uint32_t c = 0;
#pragma unroll 2
        for (uint32_t q = w0; q < w1; q++) {
            const uint32_t v = stg[q] ^ (c << 16);
            c = (uint32_t)sh.crc4[3][v >> 24] ^ (uint32_t)sh.crc4[2][(v >> 16) & 255u] ^ (uint32_t)sh.crc4[1][(v >> 8) & 255u] ^
                (uint32_t)sh.crc4[0][v & 255u];
        }
        uint32_t hi_byte = w1 * 4u;
        if (tail_owner) {
#pragma unroll 1
            for (uint32_t idx = nwords * 4u; idx < nbytes; idx++) {
                const uint32_t byte = (stg[idx >> 2] >> (24u - 8u * (idx & 3u))) & 0xFFu;
                c = ((c << 8) ^ (uint32_t)sh.crc4[0][(c >> 8) ^ byte]) & 0xFFFFu;
            }
            hi_byte = nbytes;
        }
        const uint32_t behind = sh.payload_bytes - (byte0 + hi_byte);
        if (c) atomicXor(&sh.crc, gf16_mul(c, (uint32_t)crc_w[behind]));
    };
    if (last && !have_off && B > 32) {
        /* warp 0 looks back for the frame's offset while the others work on the CRC */
        if (tid < 32) look_back_warp(shp, tid);
        else crc_slices(tid - 32, B - 32);
        __syncthreads();
        copy_out();
        return;
    }
    if (!have_off) resolve_offset_v2(shp, tid); /* block-uniform */
    copy_out();
    crc_slices(tid, B);
    if (last) {
        __syncthreads();
        return;
    }
    const uint32_t carry_idx = wend - sh.wbase;
    const uint32_t carry = carry_idx < stg_words + 4u ? stg[carry_idx] : 0u;
    __syncthreads();
#pragma unroll 1
    for (uint32_t idx = (uint32_t)tid; idx < stg_words + 4u; idx += (uint32_t)B) stg[idx] = idx == 0 ? carry : 0u;
    if (tid == 0) sh.wbase = wend;
    __syncthreads();
}

/* or `nb` bits (1 <= nb <= 32) into the window at frame bit position bp (binary.py:144-216, MSB first) */
__device__ __forceinline__ void or_bits2(uint32_t *stg, uint32_t wbase, uint32_t bp, uint32_t v, int nb)
{
    const uint32_t wi = (bp >> 5) - wbase;
    const u64 x = (u64)v << (64 - (int)(bp & 31u) - nb);
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    atomicOr(&stg[wi], hi);
    if (lo) atomicOr(&stg[wi + 1u], lo);
}

/* sum of the channel's zig-zag residuals [0, pos): per-thread spans of S samples, warp-local
 * exclusive prefixes in cumw, warp totals in wtot; partial spans from the plane (u written back) */
template <int S>
__device__ __forceinline__ u64 cum_v2(const SerCtx &b, const int32_t *x, int pos)
{
    const int g = pos / S;
    const int wg = g >> 5;
    u64 s = 0;
    for (int w = 0; w < b.nwarps; w++) {
        const u64 v = b.sh.wtot[w];
        if (w < wg) s += v;
    }
    if (g < b.B) s += b.cumw[g];
    for (int j = g * S; j < pos; j++) s += (uint32_t)x[swz2((uint32_t)(V2_GUARD + j))];
    return s;
}

/* ------------------------------------------------------------------------------------- */
/* phase 1 of one channel: [fixed-predictor errors and choice,] residual, Rice search, size */

template <int LMAX, bool CHOSEN, int S>
__device__ __forceinline__ void size_channel(SerCtx &b, int c)
{
    SerShared &sh = b.sh;
    const EncParams &p = b.p;
    const int n = b.n;
    const int par = c & 1;
    int32_t *x = b.planes + (uint32_t)c * b.pstride;
    uint8_t *kall = b.kall + (uint32_t)c * b.kall_stride;
    const int i0 = b.tid * S;
    Decision &d = sh.dec[c];
    flacb_subframe_info *inf = p.info ? p.info + ((u64)sh.frame * (u64)p.channels + (u64)c) : nullptr;

    if (b.tid < 16) sh.sizeA[par][b.tid] = 0, sh.sizeB[par][b.tid] = 0, sh.any5[par][b.tid] = 0;
    if (b.tid == 0) sh.errkey[par] = 0xFFFFFFFFu;

    int type, order, warm;
    uint32_t u[S];
    bool range_bad = false;
    if (!CHOSEN || d.type == 0) {
        int a[S + 4];
        ld2<S + 4>(x, (uint32_t)(V2_GUARD + i0 - 4), a);
        if constexpr (!CHOSEN) {
            /* encoder.py:331-359: sum |residual| of the five fixed predictors (common.py:15-21),
             * by repeated differences; warm-up samples carry no residual */
            int r1[S + 3], r2[S + 2], r3[S + 1], r4[S];
#pragma unroll
            for (int j = 0; j < S + 3; j++) r1[j] = a[j + 1] - a[j];
#pragma unroll
            for (int j = 0; j < S + 2; j++) r2[j] = r1[j + 1] - r1[j];
#pragma unroll
            for (int j = 0; j < S + 1; j++) r3[j] = r2[j + 1] - r2[j];
#pragma unroll
            for (int j = 0; j < S; j++) r4[j] = r3[j + 1] - r3[j];
            u64 acc[5] = {0, 0, 0, 0, 0};
            if (i0 >= 4 && i0 + S <= n) {
#pragma unroll
                for (int e0 = 0; e0 < S; e0 += 4) { /* four terms below 2^28 fit 32 bits */
                    uint32_t t[5] = {0, 0, 0, 0, 0};
#pragma unroll
                    for (int e = e0; e < e0 + 4; e++) {
                        t[0] = __sad(a[e + 4], 0, t[0]), t[1] = __sad(r1[e + 3], 0, t[1]), t[2] = __sad(r2[e + 2], 0, t[2]);
                        t[3] = __sad(r3[e + 1], 0, t[3]), t[4] = __sad(r3[e + 1], r3[e], t[4]); /* |a - b| + t: one instruction */
                    }
#pragma unroll
                    for (int o = 0; o < 5; o++) acc[o] += t[o];
                }
            } else if (i0 < n) {
#pragma unroll
                for (int e = 0; e < S; e++) {
                    const int i = i0 + e;
                    if (i < n) {
                        const int r[5] = {a[e + 4], r1[e + 3], r2[e + 2], r3[e + 1], r4[e]};
#pragma unroll
                        for (int o = 0; o < 5; o++)
                            if (i >= o) acc[o] += (u64)(uint32_t)abs(r[o]);
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 5; o++) acc[o] = warp_sum_u64(acc[o]);
            if (b.lane == 0) {
#pragma unroll
                for (int o = 0; o < 5; o++) sh.partf[b.warp][o] = acc[o];
            }
            __syncthreads();
            /* every warp: lane o adds up order o, then the first minimum (encoder.py:352) */
            u64 tot = 0;
            if (b.lane < 5)
                for (int w = 0; w < b.nwarps; w++) tot += sh.partf[w][b.lane];
            u64 fe[5];
#pragma unroll
            for (int o = 0; o < 5; o++) fe[o] = __shfl_sync(FULL, tot, o);
            int fo = 0;
            if (n > 4) {
                u64 least = fe[0];
#pragma unroll
                for (int o = 1; o < 5; o++)
                    if (fe[o] < least) least = fe[o], fo = o;
            }
            type = 0, order = fo, warm = fo;
            if (b.tid == 0) {
                d.type = 0, d.order = fo, d.warm = fo, d.shift = 0, d.ncoef = 0;
                if (inf) {
                    inf->type = 0, inf->order = fo, inf->shift = 0, inf->precision = 0, inf->fixed_order = fo, inf->lpc_order = 0;
                    for (int o = 0; o < 5; o++) inf->fixed_err[o] = (n <= 4 && o > 0) ? -1 : (int64_t)fe[o];
                    for (int m = 0; m < FLACB_MAX_LPC; m++) {
                        inf->lpc_err[m] = 0, inf->lpc_shift[m] = 0, inf->lpc_ncoef[m] = 0;
                        for (int j = 0; j < FLACB_MAX_LPC; j++) inf->lpc_q[m][j] = 0;
                    }
                    for (int l = 0; l <= FLACB_MAX_LPC; l++) inf->autoc[l] = 0.0;
                }
            }
        } else {
            type = 0, order = d.order, warm = d.warm;
        }
        /* the chosen predictor (common.py:15-21) on the thread's samples */
        const int c1 = order, c2 = order == 2 ? -1 : order == 3 ? -3 : order == 4 ? -6 : 0;
        const int c3 = order == 3 ? 1 : order == 4 ? 4 : 0, c4 = order == 4 ? -1 : 0;
#pragma unroll
        for (int e = 0; e < S; e++) {
            const int rr = a[e + 4] - (c1 * a[e + 3] + c2 * a[e + 2] + c3 * a[e + 1] + c4 * a[e]);
            u[e] = ((uint32_t)rr << 1) ^ (uint32_t)(rr >> 31);
        }
        if (i0 < warm) {
#pragma unroll
            for (int e = 0; e < S; e++)
                if (i0 + e < warm && i0 + e < FLACB_MAX_LPC) sh.warm[c][i0 + e] = a[e + 4];
        }
    } else {
        /* encoder.py:537-548 prediction_residual with the chosen quantised coefficients */
        type = 1, order = d.order, warm = d.warm;
        if constexpr (CHOSEN) {
            int h[LMAX + S];
            ld2<LMAX + S>(x, (uint32_t)(V2_GUARD + i0 - LMAX), h);
            const int ncoef = d.ncoef, shift = d.shift;
            uint32_t qa = 0;
            for (int j = 0; j < ncoef; j++) qa += (uint32_t)abs(d.q[j]);
            if (((u64)qa << (p.bps - 1)) <= (1ull << 31) - 1) {
                /* every partial sum fits 32 bits */
#pragma unroll
                for (int e0 = 0; e0 < S; e0 += 4) {
                    int acc[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[e] = 0;
#pragma unroll
                    for (int j = 0; j < LMAX; j++) {
                        if (j < ncoef) {
                            const int cj = d.q[j];
#pragma unroll
                            for (int e = 0; e < 4; e++) acc[e] += h[LMAX + e0 + e - 1 - j] * cj;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const i64 r = (i64)h[LMAX + e0 + e] - (i64)(acc[e] >> shift);
                        const u64 z = ((u64)r << 1) ^ (u64)(r >> 63);
                        if ((z >> 32) != 0 && i0 + e0 + e >= warm && i0 + e0 + e < n) range_bad = true;
                        u[e0 + e] = (uint32_t)z;
                    }
                }
            } else {
#pragma unroll
                for (int e0 = 0; e0 < S; e0 += 4) {
                    i64 acc[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[e] = 0;
#pragma unroll
                    for (int j = 0; j < LMAX; j++) {
                        if (j < ncoef) {
                            const i64 cj = d.q[j];
#pragma unroll
                            for (int e = 0; e < 4; e++) acc[e] += (i64)h[LMAX + e0 + e - 1 - j] * cj;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const i64 r = (i64)h[LMAX + e0 + e] - (acc[e] >> shift);
                        const u64 z = ((u64)r << 1) ^ (u64)(r >> 63);
                        if ((z >> 32) != 0 && i0 + e0 + e >= warm && i0 + e0 + e < n) range_bad = true;
                        u[e0 + e] = (uint32_t)z;
                    }
                }
            }
            if (i0 < warm) {
#pragma unroll
                for (int e = 0; e < S; e++)
                    if (i0 + e < warm && i0 + e < FLACB_MAX_LPC) sh.warm[c][i0 + e] = h[LMAX + e];
            }
        }
    }
    if (i0 < warm || i0 + S > n) {
#pragma unroll
        for (int e = 0; e < S; e++)
            if (i0 + e < warm || i0 + e >= n) u[e] = 0;
    }
    if (range_bad) sh.range_flag[c] = 1;

    /* candidate partition orders (encoder.py:664-667): both tests are monotone in the order */
    int o_first = p.rmin, o_last = -1;
    {
        int o = __ffs(n) - 1;
        if (o > p.rmax) o = p.rmax;
        while (o >= p.rmin && (n >> o) <= warm) o--;
        o_last = o >= p.rmin ? o : -1;
    }
    const bool no_cand = o_last < 0; /* encoder.py:669 assert; laid out with order 0 so that the stream stays well formed */
    if (no_cand) o_first = 0, o_last = 0;
    const int ncand = o_last - o_first + 1;
    const uint32_t psz_f = (uint32_t)n >> o_last;
    const bool aligned = (psz_f % (uint32_t)S) == 0; /* every partition boundary is a span boundary */
    if (b.tid == 0) sh.magic_f[par] = psz_f == 1 ? 0u : 0xFFFFFFFFu / psz_f + 1u;

    /* block scan of the span sums */
    {
        u64 ts = 0;
#pragma unroll
        for (int e = 0; e < S; e++) ts += u[e];
        u64 inc = ts;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u64 t = __shfl_up_sync(FULL, inc, off);
            if (b.lane >= off) inc += t;
        }
        b.cumw[b.tid] = inc - ts;
        if (b.lane == 31) sh.wtot[b.warp] = inc;
    }
    __syncthreads(); /* #1: every thread has its inputs in registers, span sums are visible */
    st2<S>(x, (uint32_t)(V2_GUARD + i0), u); /* residual in place of the samples */
    if (!aligned) __syncthreads();           /* partial spans are read from the plane */

    /* parameter k = floor(log2(sum / len)) of every partition of every candidate order
     * (encoder.py:714-753), heap order: h = 2^o + partition */
    for (uint32_t h = (1u << o_first) + (uint32_t)b.tid; h < (2u << o_last); h += (uint32_t)b.B) {
        const int o = 31 - __clz(h);
        const int pi = (int)(h - (1u << o));
        const int psz = n >> o;
        const int start = pi * psz, end = start + psz;
        const uint32_t len = (uint32_t)(psz - (pi == 0 ? warm : 0));
        const u64 sum = cum_v2<S>(b, x, end) - cum_v2<S>(b, x, start);
        const int ci = o - o_first;
        int k;
        if (sum < (u64)len) {
            /* mean < 1: math domain error (sum == 0, encoder.py:753) or a negative shift count
             * (encoder.py:758).  Lenient policy: parameter 0. */
            atomicMin(&sh.errkey[par], ((uint32_t)ci << 20) | ((uint32_t)pi << 4) | FLACB_E_VALUE);
            k = 0;
        } else {
            k = (63 - __clzll(sum)) - (31 - __clz(len));
            if (((u64)len << k) > sum) k--;
        }
        kall[h] = (uint8_t)k;
        atomicAdd(&sh.sizeA[par][ci], 4u + (k > 14 ? 5u : 4u)); /* encoder.py:718-725 */
        atomicAdd(&sh.sizeB[par][ci], len * (uint32_t)(1 + k));
        if (k > 14) atomicOr(&sh.any5[par][ci], 1u);
    }
    __syncthreads(); /* #2 */

    /* sum over the samples of u >> k for every candidate order (rice_size, encoder.py:756-760).
     * A partition's total is below 2 len, so 32-bit sums are exact. */
    {
        uint32_t mine = 0;
        const uint32_t magic_f = sh.magic_f[par];
        if (aligned) {
            const uint32_t pf = i0 < n ? fast_div((uint32_t)i0, psz_f, magic_f) : 0u;
            uint32_t kprev = 0xFFFFFFFFu, t32 = 0;
#pragma unroll
            for (int j = 0; j < V2_MAXC; j++) {
                if (j < ncand) { /* block-uniform */
                    const int o = o_first + j;
                    const uint32_t k0 = kall[(1u << o) + (pf >> (o_last - o))];
                    if (k0 != kprev) {
                        t32 = 0;
#pragma unroll
                        for (int e = 0; e < S; e++) t32 += u[e] >> k0;
                        kprev = k0;
                    }
                    const uint32_t tot = __reduce_add_sync(FULL, t32);
                    if (b.lane == j) mine = tot;
                }
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < ncand; j++) {
                const int o = o_first + j;
                const uint32_t psz = (uint32_t)n >> o;
                const uint8_t *kk = kall + (1u << o);
                uint32_t t32 = 0;
#pragma unroll
                for (int e = 0; e < S; e++)
                    if (i0 + e < n) t32 += u[e] >> kk[(uint32_t)(i0 + e) / psz];
                t32 = __reduce_add_sync(FULL, t32);
                if (b.lane == j) mine = t32;
            }
        }
        if (b.lane < 12) sh.part[par][b.warp][b.lane] = mine;
    }
    __syncthreads(); /* #3 */
    /* every warp: lane j adds up candidate j; first minimum of the reference's size estimate
     * (with its per-partition "+4", encoder.py:685-695) */
    {
        uint32_t q = 0, sB = 0, a5 = 0, key = 0xFFFFFFFFu;
        if (b.lane < ncand) {
            for (int w = 0; w < b.nwarps; w++) q += sh.part[par][w][b.lane];
            sB = sh.sizeB[par][b.lane];
            a5 = sh.any5[par][b.lane];
            key = ((sh.sizeA[par][b.lane] + sB + q) << 4) | (uint32_t)b.lane;
        }
        const int best = (int)(__reduce_min_sync(FULL, key) & 15u);
        sB = __shfl_sync(FULL, sB, best);
        q = __shfl_sync(FULL, q, best);
        a5 = __shfl_sync(FULL, a5, best);
        if (b.tid == 0) {
            d.porder = o_first + best;
            d.method = a5 ? 5 : 4; /* encoder.py:648-650 */
            /* method (2) + partition order (4) + one parameter per partition + the codes */
            const uint32_t res_bits = 6u + ((uint32_t)d.method << d.porder) + sB + q;
            /* encoder.py:553-627: header byte, warm-up, [precision, shift, coefficients], residual */
            uint32_t bits = 8u + (uint32_t)warm * (uint32_t)p.bps + res_bits;
            if (type == 1) bits += 4u + 5u + (uint32_t)d.ncoef * (uint32_t)p.P;
            d.bits = bits;
            const uint32_t psz = (uint32_t)n >> d.porder;
            sh.dec_magic[c] = psz == 1 ? 0u : 0xFFFFFFFFu / psz + 1u;
            int &st = sh.chan_status[c];
            if (st == 0 && sh.range_flag[c]) st = FLACB_E_RANGE;
            if (st == 0 && no_cand) st = FLACB_E_ASSERT;
            if (st == 0 && sh.errkey[par] != 0xFFFFFFFFu) st = (int)(sh.errkey[par] & 15u);
            if (inf) {
                inf->porder = d.porder;
                inf->method = d.method;
                inf->bits = bits;
                inf->nparams = 1 << d.porder;
                for (int i = 0; i < FLACB_INFO_PARAMS; i++) inf->params[i] = i < (1 << d.porder) ? kall[(1u << d.porder) + i] : 0;
            }
        }
    }
    (void)order;
}

/* ------------------------------------------------------------------------------------- */
/* phase 2 of one channel: every field of the subframe into the staging window              */

template <int S>
__device__ __forceinline__ void pack_channel(SerCtx &b, int c, const uint16_t *crc_w)
{
    SerShared &sh = b.sh;
    const EncParams &p = b.p;
    const int n = b.n;
    const int par = c & 1;
    const Decision &d = sh.dec[c];
    const int32_t *x = b.planes + (uint32_t)c * b.pstride;
    const uint8_t *kk = b.kall + (uint32_t)c * b.kall_stride + (1u << d.porder);
    const int i0 = b.tid * S;
    const uint32_t bit0 = sh.chan_bit0[c], bit_end = sh.chan_bit0[c + 1];
    /* the subframe must lie inside the window: flush what is in front of it when it would not */
    if ((bit_end >> 5) + 2u > sh.wbase + b.stg_words) {
        if ((bit0 >> 5) > sh.wbase) flush_v2(&sh, b.stg, b.stg_words, crc_w, b.tid, b.B, ((bit0 >> 5) - sh.wbase) * 4u, bit0 >> 5, 0);
        if ((bit_end >> 5) + 2u > sh.wbase + b.stg_words) {
            /* larger than the worst case the window was sized for (only possible on out-of-range
             * input): the frame is marked and this subframe is left out */
            if (b.tid == 0 && sh.status == 0) sh.status = FLACB_E_NOSPACE;
            return;
        }
    }
    const uint32_t wbase = sh.wbase;
    const int warm = d.warm, method = d.method;
    const uint32_t res_bit0 = bit0 + 8u + (uint32_t)warm * (uint32_t)p.bps + (d.type ? 9u + (uint32_t)d.ncoef * (uint32_t)p.P : 0u);
    const uint32_t psz = (uint32_t)n >> d.porder, magic = sh.dec_magic[c];
    const uint32_t hv = ((method == 4 ? 0u : 1u) << 4) | (uint32_t)d.porder; /* encoder.py:766-767 */

    uint32_t u[S];
    {
        int ui[S];
        ld2<S>(x, (uint32_t)(V2_GUARD + i0), ui);
#pragma unroll
        for (int e = 0; e < S; e++) u[e] = (uint32_t)ui[e];
    }
    /* parameter of sample i's partition and the fields in front of its code: the parameter at
     * the start of a partition, method + partition order in front of the first one */
    auto param_of = [&](int i, uint32_t &k, uint32_t &pre) {
        const uint32_t pi = fast_div((uint32_t)i, psz, magic);
        const uint32_t pstart = pi * psz < (uint32_t)warm ? (uint32_t)warm : pi * psz;
        k = kk[pi];
        pre = (uint32_t)i == pstart ? (uint32_t)method + (i == warm ? 6u : 0u) : 0u;
    };
    /* the usual span: inside the block, behind the warm-up, in one partition */
    const uint32_t p0 = i0 < n ? fast_div((uint32_t)i0, psz, magic) : 0u;
    const bool usual = i0 >= warm && i0 + S <= n && (uint32_t)(i0 + S) <= (p0 + 1) * psz;
    uint32_t k0 = 0, pre0 = 0, nbits = 0;
    if (usual) {
        param_of(i0, k0, pre0);
#pragma unroll
        for (int e = 0; e < S; e++) nbits += u[e] >> k0;
        nbits += (uint32_t)S * (k0 + 1u) + pre0;
    } else if (i0 < n) {
#pragma unroll
        for (int e = 0; e < S; e++) {
            const int i = i0 + e;
            if (i >= warm && i < n) {
                uint32_t k, pre;
                param_of(i, k, pre);
                nbits += pre + (u[e] >> k) + 1u + k;
            }
        }
    }
    uint32_t inc = nbits;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, off);
        if (b.lane >= off) inc += t;
    }
    if (b.lane == 31) sh.wtot2[par][b.warp] = inc;
    __syncthreads();
    uint32_t cur = res_bit0 + inc - nbits;
    for (int w = 0; w < b.warp; w++) cur += sh.wtot2[par][w];

    if (nbits) {
        if (usual) {
            if (pre0) {
                if (pre0 > (uint32_t)method) or_bits2(b.stg, wbase, cur, hv, 6), cur += 6;
                or_bits2(b.stg, wbase, cur, k0, method); /* encoder.py:791 */
                cur += (uint32_t)method;
            }
            const uint32_t stop = 1u << k0, mask = stop - 1u;
#pragma unroll
            for (int e = 0; e < S; e++) {
                cur += u[e] >> k0; /* the zeros of the unary part: the window is pre-zeroed (encoder.py:802) */
                or_bits2(b.stg, wbase, cur, stop | (u[e] & mask), (int)k0 + 1);
                cur += k0 + 1u;
            }
        } else {
            /* rolled, and from the plane: indexing the register copy would put it in local memory */
#pragma unroll 1
            for (int e = 0; e < S; e++) {
                const int i = i0 + e;
                if (i >= warm && i < n) {
                    uint32_t k, pre;
                    param_of(i, k, pre);
                    if (pre) {
                        if (pre > (uint32_t)method) or_bits2(b.stg, wbase, cur, hv, 6), cur += 6;
                        or_bits2(b.stg, wbase, cur, k, method);
                        cur += (uint32_t)method;
                    }
                    const uint32_t ue = (uint32_t)x[swz2((uint32_t)(V2_GUARD + i))];
                    cur += ue >> k;
                    or_bits2(b.stg, wbase, cur, (1u << k) | (ue & ((1u << k) - 1u)), (int)k + 1);
                    cur += k + 1u;
                }
            }
        }
    }
    /* the serial fields, one per thread from the far end of the block (encoder.py:553-627) */
    {
        const uint32_t smask = p.bps >= 32 ? 0xFFFFFFFFu : ((1u << p.bps) - 1u);
        for (int r = b.B - 1 - b.tid; r <= 2 * FLACB_MAX_LPC; r += b.B) {
            if (r == 0) {
                or_bits2(b.stg, wbase, bit0, d.type ? (0x20u | (uint32_t)(d.order - 1)) << 1 : (0x08u | (uint32_t)d.order) << 1, 8);
                if (d.type) { /* encoder.py:619-622 */
                    const uint32_t lb = bit0 + 8u + (uint32_t)warm * (uint32_t)p.bps;
                    or_bits2(b.stg, wbase, lb, (((uint32_t)(p.P - 1) & 15u) << 5) | ((uint32_t)d.shift & 31u), 9);
                }
            } else if (r <= FLACB_MAX_LPC) {
                if (r - 1 < warm)
                    or_bits2(b.stg, wbase, bit0 + 8u + (uint32_t)(r - 1) * (uint32_t)p.bps, (uint32_t)sh.warm[c][r - 1] & smask, p.bps);
            } else {
                const int j = r - 1 - FLACB_MAX_LPC;
                if (d.type && j < d.ncoef) /* encoder.py:624-625 */
                    or_bits2(b.stg, wbase, bit0 + 8u + (uint32_t)warm * (uint32_t)p.bps + 9u + (uint32_t)j * (uint32_t)p.P,
                             (uint32_t)d.q[j] & ((1u << p.P) - 1u), p.P);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------- */

template <int LMAX, bool CHOSEN, int S, int TB, int MINB>
__global__ void __launch_bounds__(TB, MINB) serialise_kernel(const EncParams p)
{
    FLACB_DYN_SMEM(smem);
    __shared__ SerShared sh;
    FLACB_SIM_POISON(sh);
    SerCtx b(p, sh);
    b.tid = (int)threadIdx.x;
    b.B = (int)blockDim.x;
    b.lane = b.tid & 31;
    b.warp = b.tid >> 5;
    b.nwarps = b.B >> 5;
    const SerLayout &lay = p.slay;
    b.planes = (int32_t *)(smem + lay.planes);
    b.pstride = lay.pstride;
    b.kall = smem + lay.kall;
    b.kall_stride = lay.kall_stride;
    b.cumw = (u64 *)(smem + lay.cumw);
    b.stg = (uint32_t *)(smem + lay.stg);
    b.stg_words = lay.stg_words;

    if (b.tid == 0) {
        /* frames are taken in launch order so that every predecessor of a frame is already
         * running (or done) when the frame looks back for its offset */
        sh.frame = atomicAdd(p.ticket, 1u);
        sh.status = 0, sh.crc = 0, sh.wbase = 0, sh.have_off = 0;
        sh.o_out = p.out, sh.o_cap = p.out_cap;
        sh.o_lb_state = p.lb_state, sh.o_frame_off = p.frame_off, sh.o_n_frames = p.n_frames;
    }
    if (b.tid < MAX_CH) sh.chan_status[b.tid] = 0, sh.range_flag[b.tid] = 0;
    for (int i = b.tid; i < 512; i += b.B) ((uint32_t *)&sh.crc4[0][0])[i] = ((const uint32_t *)p.crc4)[i];
    for (uint32_t idx = (uint32_t)b.tid; idx < b.stg_words + 4u; idx += (uint32_t)b.B) b.stg[idx] = 0;
    __syncthreads();
    const uint32_t f = sh.frame;
    const int n = b.n = (f == p.n_frames - 1) ? p.tail_size : p.block_size;

    /* ---- frame/channel split ---- */
    load_planes_v2(p.pcm + (u64)f * (u64)p.block_size * (u64)p.channels * (u64)p.cbytes, b.planes, b.pstride, lay.xwords, n,
                   p.channels, p.cbytes, b.tid, b.B);
    if constexpr (CHOSEN) {
        /* LPC configurations: the choice of every subframe was made by choose_kernel.  Everything up to
         * here (ticket, tables, the frame's PCM) did not need it: see pdl_wait in flacb200_kernels.cuh */
        pdl_wait();
        for (uint32_t i = (uint32_t)b.tid; i < (uint32_t)p.channels * (sizeof(Decision) / 4); i += (uint32_t)b.B) {
            const uint32_t c = i / (uint32_t)(sizeof(Decision) / 4), k = i - c * (uint32_t)(sizeof(Decision) / 4);
            ((uint32_t *)&sh.dec[c])[k] = ((const uint32_t *)&p.dec[(u64)f * (u64)p.channels + c].d)[k];
        }
    }
    __syncthreads();
    if constexpr (CHOSEN) {
        if (b.tid < p.channels) sh.chan_status[b.tid] = p.dec[(u64)f * (u64)p.channels + (u64)b.tid].status;
    }

    /* ---- phase 1: sizes ---- */
    for (int c = 0; c < p.channels; c++) size_channel<LMAX, CHOSEN, S>(b, c);

    /* ---- frame size -> published; the frame's own offset is looked up at the first flush ---- */
    if (b.tid == 0) {
        for (int c = p.channels - 1; c >= 0; c--) /* the frame's status: first failing channel */
            if (sh.chan_status[c]) sh.status = sh.chan_status[c];
        sh.hdr_len = build_frame_header_v2(p.flags, p.channels, n, p.first_frame_number + f, sh.hdr);
        uint32_t bits = (uint32_t)sh.hdr_len * 8u;
        for (int c = 0; c < p.channels; c++) {
            sh.chan_bit0[c] = bits;
            bits += sh.dec[c].bits;
        }
        sh.chan_bit0[p.channels] = bits;
        sh.payload_bytes = (bits + 7u) >> 3; /* encoder.py:159-160 zero padding */
        sh.frame_bytes = sh.payload_bytes + 2u; /* + CRC-16 */
        p.frame_bytes[f] = sh.frame_bytes;
        atomicExch(&p.lb_state[f], f == 0 ? (LB_PREFIX | (u64)sh.frame_bytes) : (LB_AGG | (u64)sh.frame_bytes));
        for (int i = 0; i < sh.hdr_len; i += 4) {
            const int m = sh.hdr_len - i < 4 ? sh.hdr_len - i : 4;
            uint32_t v = 0;
            for (int k = 0; k < m; k++) v = (v << 8) | sh.hdr[i + k];
            or_bits2(b.stg, 0, 8u * (uint32_t)i, v, 8 * m);
        }
    }
    __syncthreads();

    /* ---- phase 2: serialise ---- */
    for (int c = 0; c < p.channels; c++) pack_channel<S>(b, c, p.crc_w);
    /* encoder.py:159-165: pad to a byte, CRC-16 over the whole frame, append big endian */
    flush_v2(&sh, b.stg, b.stg_words, p.crc_w, b.tid, b.B, sh.payload_bytes - sh.wbase * 4u, (sh.payload_bytes + 3u) >> 2, 1);
    if (b.tid == 0) {
        if (sh.writable) {
            uint8_t *dst = p.out + sh.frame_off + sh.payload_bytes;
            dst[0] = (uint8_t)(sh.crc >> 8);
            dst[1] = (uint8_t)sh.crc;
        }
        p.frame_status[f] = sh.status;
        if (p.stats) {
            for (int c = 0; c < p.channels; c++) {
                if (sh.dec[c].type) atomicAdd(&p.stats[5], 1ull);
                else atomicAdd(&p.stats[sh.dec[c].order], 1ull);
            }
            if (sh.status >= FLACB_E_ASSERT && sh.status <= FLACB_E_ZERODIV) atomicAdd(&p.stats[6], 1ull);
        }
    }
}

#endif /* !FLACB_HOST_ONLY */

} /* namespace flacb */
