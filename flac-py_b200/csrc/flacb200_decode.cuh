/*
 * flacb200_decode.cuh -- device code of the B200 FLAC decode path (BASELINE.json config 5).
 *
 * The reference decoder (flac/decoder.py) walks the stream serially: frame boundaries are not
 * stored anywhere (STREAMINFO frame sizes are 0, no seek table) and a frame's length is only
 * known once its last Rice code has been read.  To decode all frames in parallel:
 *
 *   D1  find_sync_kernel      every byte position that looks like a frame sync (0xFF 0xF8/0xF9,
 *                             decoder.py:136-138) becomes a candidate, kept in stream order
 *   D2  parse_headers_kernel  one thread per candidate: frame header fields (decoder.py:133-262)
 *   D3  decode_frames_kernel  one thread per candidate: subframes, Rice residuals and the serial
 *                             prediction restore (decoder.py:267-421, :453-498), into a planar
 *                             scratch area; records where the frame ends
 *   D4  chain_kernel          which candidates are real: the frames the reference would visit
 *                             are exactly the chain "frame 0, then the candidate that starts
 *                             where the previous one ended"; found by pointer jumping.  False
 *                             candidates (sync patterns inside Rice data) are never on it.
 *   D5  interleave_kernel     undo inter-channel decorrelation (decoder.py:431-451) and write
 *                             interleaved samples at their final offsets
 *
 * Like the reference, CRC-8 / CRC-16 are not verified (decoder.py:128, :175) and a truncated
 * last frame is dropped (decoder.py:104-108).  Quirks of the reference decoder that change
 * results are reproduced (wasted-bits count, M/S reconstruction), see SURVEY.md 9.7.
 */
#pragma once

#include <stdint.h>

#ifndef FLACB_CPU_SIM
#include <cuda_runtime.h>
#endif

#include "../../include/flacb200.h"

namespace flacb {

typedef unsigned long long u64;
typedef long long i64;

constexpr int SYNC_THREADS = 256;
constexpr int SYNC_BYTES_PER_THREAD = 16;
constexpr int SYNC_CHUNK = SYNC_THREADS * SYNC_BYTES_PER_THREAD;

struct Candidate {
    u64 start;        /* byte offset of the sync code */
    u64 end;          /* byte offset just behind the frame's CRC-16; 0 = did not parse */
    u64 tmp_off;      /* first sample of the frame in the planar scratch area */
    u64 out_off;      /* first sample frame in the output (chain frames only) */
    uint32_t next;    /* candidate that starts at `end` (chain link) */
    uint32_t hdr_bits;
    int32_t bs;       /* block size */
    int32_t nch;      /* subframes in the frame */
    int32_t chcode;   /* channel assignment nibble */
    int32_t sample_size;
    int32_t status;   /* FLACB_OK, or what the reference raises inside this frame; E_EOF = truncated */
    int32_t on_chain;
};

struct DecodeResult {
    u64 n_frames, n_sample_frames;
    int32_t status, frame_channels;
    u64 bad_frame;
};

#ifndef FLACB_HOST_ONLY

/* ---- D1 ------------------------------------------------------------------------------ */

/* pass 0: counts[chunk] = candidates in the chunk; pass 1: positions written in order */
__global__ void __launch_bounds__(SYNC_THREADS) find_sync_kernel(const uint8_t *s, u64 len, uint32_t *counts,
                                                                  const u64 *offsets, Candidate *cand, int write)
{
    __shared__ uint32_t warp_tot[SYNC_THREADS / 32];
    const u64 base = (u64)blockIdx.x * SYNC_CHUNK + (u64)threadIdx.x * SYNC_BYTES_PER_THREAD;
    uint32_t hits = 0; /* bit i: position base + i is a candidate */
    if (base < len) {
        uint8_t b[SYNC_BYTES_PER_THREAD + 1];
#pragma unroll
        for (int i = 0; i <= SYNC_BYTES_PER_THREAD; i++) b[i] = base + i < len ? s[base + i] : 0;
#pragma unroll
        for (int i = 0; i < SYNC_BYTES_PER_THREAD; i++)
            if (b[i] == 0xFF && (b[i + 1] & 0xFE) == 0xF8 && base + i + 1 < len) hits |= 1u << i;
    }
    const uint32_t mine = (uint32_t)__popc(hits);
    uint32_t inc = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
    for (int w = 0; w < SYNC_THREADS / 32; w++) {
        if (w < warp) wbase += warp_tot[w];
        tot += warp_tot[w];
    }
    if (!write) {
        if (threadIdx.x == 0) counts[blockIdx.x] = tot;
        return;
    }
    u64 dst = offsets[blockIdx.x] + wbase + inc - mine;
    while (hits) {
        const int i = __ffs((int)hits) - 1;
        hits &= hits - 1;
        Candidate c;
        c.start = base + (u64)i;
        c.end = 0;
        c.tmp_off = 0;
        c.out_off = 0;
        c.next = 0xFFFFFFFFu;
        c.hdr_bits = 0;
        c.bs = 0;
        c.nch = 0;
        c.chcode = 0;
        c.sample_size = 0;
        c.status = FLACB_E_EOF;
        c.on_chain = 0;
        cand[dst++] = c;
    }
}

/* ---- bit reader (binary.py:78-141 Get) -------------------------------------------------- */

struct BitReader {
    const uint8_t *p;
    u64 len;      /* bytes */
    u64 next;     /* next byte to load */
    u64 buf;      /* unread bits, left aligned */
    int avail;
    bool eof;

    __device__ __forceinline__ void init(const uint8_t *s, u64 length, u64 byte_pos)
    {
        p = s, len = length, next = byte_pos, buf = 0, avail = 0, eof = false;
    }
    __device__ __forceinline__ void refill()
    {
        /* aligned 32-bit loads while a whole word is left (keeps 33..64 bits buffered), single
         * bytes in front of the first aligned word and behind the last one */
        for (;;) {
            if (next >= len) return;
            if (next + 4 <= len && (((uintptr_t)(p + next)) & 3) == 0) {
                if (avail > 32) return;
                const uint32_t w = *(const uint32_t *)(p + next);
                buf |= (u64)__byte_perm(w, 0, 0x0123) << (32 - avail); /* the stream is big endian */
                avail += 32;
                next += 4;
            } else {
                if (avail > 56) return;
                buf |= (u64)p[next++] << (56 - avail);
                avail += 8;
            }
        }
    }
    /* n in 0..32 */
    __device__ __forceinline__ uint32_t read(int n)
    {
        if (n == 0) return 0;
        if (avail < n) refill();
        if (avail < n) { /* binary.py:40 EOFError */
            eof = true;
            avail = 0, buf = 0;
            return 0;
        }
        const uint32_t v = (uint32_t)(buf >> (64 - n));
        buf <<= n;
        avail -= n;
        return v;
    }
    /* n in 1..40 (side channels of 32-bit streams need 33) */
    __device__ __forceinline__ i64 read_signed(int n)
    {
        u64 x;
        if (n > 32) {
            const u64 hi = read(n - 32);
            x = (hi << 32) | read(32);
        } else x = read(n);
        return (i64)x - (i64)((x >> (n - 1)) << n); /* binary.py:129-131 sint */
    }
    /* zeros before the next one bit; the one is consumed (decoder.py:416-418) */
    __device__ __forceinline__ u64 unary()
    {
        u64 cnt = 0;
        for (;;) {
            if (avail == 0) {
                refill();
                if (avail == 0) {
                    eof = true;
                    return cnt;
                }
            }
            if (buf == 0) {
                cnt += (u64)avail;
                avail = 0;
                continue;
            }
            const int z = __clzll(buf);
            cnt += (u64)z;
            buf <<= z;
            buf <<= 1;
            avail -= z + 1;
            return cnt;
        }
    }
    __device__ __forceinline__ u64 bit_pos() const { return next * 8 - (u64)avail; }
};

/* ---- D2 ------------------------------------------------------------------------------ */

/* decoder.py:133-185 get_frame_header and the field decoders :188-262.  Returns the status the
 * reference raises, FLACB_E_EOF when the stream ends inside the header. */
__device__ __forceinline__ int parse_frame_header(BitReader &g, int si_bits, Candidate &c)
{
    const uint32_t sync = g.read(15);
    if (g.eof) return FLACB_E_EOF;
    if (sync != 0x7FFC) return FLACB_E_ASSERT;
    (void)g.read(1); /* blocking strategy */
    const int bsc = (int)g.read(4);
    if (g.eof) return FLACB_E_EOF;
    if (!(bsc > 0 && bsc < 15)) return FLACB_E_ASSERT; /* decoder.py:194 (rejects 0b1111 too) */
    const int src = (int)g.read(4);
    if (g.eof) return FLACB_E_EOF;
    if (!(src < 15)) return FLACB_E_ASSERT;
    if (src == 11) return FLACB_E_KEY; /* SAMPLE_RATE_VALUE_DECODING has no 0b1011 */
    const int chc = (int)g.read(4);
    if (g.eof) return FLACB_E_EOF;
    if (!(chc <= 10)) return FLACB_E_ASSERT;
    const int ssc = (int)g.read(3);
    if (g.eof) return FLACB_E_EOF;
    if (ssc == 3) return FLACB_E_ASSERT;
    const uint32_t rsv = g.read(1);
    if (g.eof) return FLACB_E_EOF;
    if (rsv != 0) return FLACB_E_ASSERT;
    const int b0 = (int)g.read(8); /* coded number, decoder.py:251-262 */
    if (g.eof) return FLACB_E_EOF;
    const int follow = b0 >= 0xFE ? 6 : b0 >= 0xFC ? 5 : b0 >= 0xF8 ? 4 : b0 >= 0xF0 ? 3 : b0 >= 0xE0 ? 2 : b0 >= 0xC0 ? 1 : 0;
    for (int i = 0; i < follow; i++) (void)g.read(8);
    if (g.eof) return FLACB_E_EOF;
    int bs;
    if (bsc == 1) bs = 192;
    else if (bsc >= 2 && bsc <= 5) bs = 144 << bsc;
    else if (bsc == 6) bs = (int)g.read(8) + 1;
    else if (bsc == 7) bs = (int)g.read(16) + 1;
    else bs = 1 << bsc;
    if (g.eof) return FLACB_E_EOF;
    if (src == 12) (void)g.read(8);
    else if (src == 13 || src == 14) (void)g.read(16);
    (void)g.read(8); /* CRC-8, never verified (decoder.py:175) */
    if (g.eof) return FLACB_E_EOF;
    const int ss_tab[8] = {0, 8, 12, 0, 16, 20, 24, 32};
    c.sample_size = ssc == 0 ? si_bits : ss_tab[ssc];
    c.bs = bs;
    c.chcode = chc;
    c.nch = chc < 8 ? chc + 1 : 2; /* common.py:215-230 */
    return FLACB_OK;
}

/* Status of a candidate whose header CRC-8 does not match.  The reference never checks the
 * CRC (decoder.py:175), so such a candidate is only skipped on the fast path: if the frame chain
 * ever lands on one, the host repeats the decode without the filter. */
constexpr int STATUS_CRC8_MISMATCH = 101;
/* Same idea for candidates whose block size exceeds, or whose channel count differs from, the
 * first frame's: in a fixed-block-size stream no real frame looks like that, and a garbage
 * "frame" of 65 536 x 8 samples would keep one thread busy for tens of milliseconds. */
constexpr int STATUS_DEFERRED = 102;

__global__ void parse_headers_kernel(const uint8_t *s, u64 len, int si_bits, Candidate *cand, uint32_t n_cand,
                                     uint32_t *sizes, int check_crc8)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    Candidate c = cand[i];
    BitReader g;
    g.init(s, len, c.start);
    c.status = parse_frame_header(g, si_bits, c);
    c.hdr_bits = (uint32_t)(g.bit_pos() - c.start * 8);
    if (c.status == FLACB_OK && check_crc8) {
        /* sync patterns inside Rice data almost never carry a valid header CRC (crc.py:17-22);
         * dropping them here keeps garbage "frames" with huge block sizes off the GPU */
        const uint32_t nb = c.hdr_bits / 8;
        uint32_t crc = 0;
        for (uint32_t i = 0; i + 1 < nb; i++) {
            crc ^= s[c.start + i];
            for (int bit = 0; bit < 8; bit++) crc = (crc & 0x80u) ? ((crc << 1) ^ 0x07u) & 0xFFu : (crc << 1) & 0xFFu;
        }
        if (crc != s[c.start + nb - 1]) c.status = STATUS_CRC8_MISMATCH;
    }
    if (c.status != FLACB_OK) c.bs = 0, c.nch = 0;
    cand[i] = c;
    sizes[i] = (uint32_t)c.bs * (uint32_t)c.nch;
}

/* ---- D3 ------------------------------------------------------------------------------ */

struct History { /* last 32 reconstructed samples, circular */
    i64 h[32];
    int head;
    __device__ __forceinline__ void push(i64 v)
    {
        h[head & 31] = v;
        head++;
    }
    __device__ __forceinline__ i64 back(int j) const { return h[(head - 1 - j) & 31]; } /* x[i-1-j] */
};

/* decoder.py:358-421 get_residual with the prediction of decoder.py:493-498 applied on the fly:
 * out[i] = residual + (sum_j out[i-1-j] * coef[j] >> shift) */
__device__ __forceinline__ int decode_residual(BitReader &g, int bs, int order, const int32_t *coef, int shift, History &hist,
                               int32_t *out, bool is_fixed)
{
    i64 x1 = 0, x2 = 0, x3 = 0, x4 = 0; /* out[pos-1] .. out[pos-4] for the fixed predictors */
    if (is_fixed) {
        if (order >= 1) x1 = hist.back(0);
        if (order >= 2) x2 = hist.back(1);
        if (order >= 3) x3 = hist.back(2);
        if (order >= 4) x4 = hist.back(3);
    }
    const uint32_t m = g.read(2);
    if (g.eof) return FLACB_E_EOF;
    if (m > 1) return FLACB_E_VALUE; /* RiceCodingMethod(x) */
    const int method = m == 0 ? 4 : 5;
    const int porder = (int)g.read(4);
    if (g.eof) return FLACB_E_EOF;
    const int nparts = 1 << porder;
    if (bs % nparts != 0) return FLACB_E_ASSERT;
    if (!((bs >> porder) > order)) return FLACB_E_ASSERT;
    int pos = order;
    for (int p = 0; p < nparts; p++) {
        const int cnt = p == 0 ? (bs >> porder) - order : (bs >> porder);
        const int k = (int)g.read(method);
        if (g.eof) return FLACB_E_EOF;
        const bool escaped = k == (1 << method) - 1; /* decoder.py:404-407 */
        int size = 0;
        if (escaped) {
            size = (int)g.read(5);
            if (g.eof) return FLACB_E_EOF;
            if (size == 0 && cnt > 0) return FLACB_E_VALUE; /* sint(0): negative shift count */
        }
        for (int i = 0; i < cnt; i++) {
            i64 r;
            if (escaped) r = g.read_signed(size);
            else {
                const u64 msb = g.unary();
                const u64 lsb = g.read(k);
                const u64 u = (msb << k) | lsb;
                r = (i64)(u >> 1) ^ -(i64)(u & 1); /* utils.py:87-88 zigzag_decode */
            }
            if (g.eof) return FLACB_E_EOF;
            i64 v;
            if (is_fixed) {
                /* fixed predictors (common.py:15-21) on the last four samples, kept in registers */
                const i64 pred = order == 0 ? 0 : order == 1 ? x1 : order == 2 ? 2 * x1 - x2
                                 : order == 3 ? 3 * x1 - 3 * x2 + x3 : 4 * x1 - 6 * x2 + 4 * x3 - x4;
                v = r + pred;
                x4 = x3, x3 = x2, x2 = x1, x1 = v;
            } else {
                i64 acc = 0;
                for (int j = 0; j < order; j++) acc += hist.back(j) * (i64)coef[j];
                v = r + (acc >> shift);
                hist.push(v);
            }
            out[pos++] = (int32_t)v;
        }
    }
    return FLACB_OK;
}

/* decoder.py:267-355 get_subframe + :453-498 decode_subframe */
__device__ __forceinline__ int decode_subframe(BitReader &g, int bs, int sample_size, int32_t *out)
{
    const uint32_t padbit = g.read(1);
    if (g.eof) return FLACB_E_EOF;
    if (padbit != 0) return FLACB_E_ASSERT;
    const int type = (int)g.read(6);
    if (g.eof) return FLACB_E_EOF;
    if (!(type <= 1 || (type >= 8 && type <= 12) || type >= 32)) return FLACB_E_ASSERT;
    int wasted = 0;
    const uint32_t wflag = g.read(1);
    if (g.eof) return FLACB_E_EOF;
    if (wflag) { /* decoder.py:346-355: counts the zeros (one less than the format says) and
                    never shifts the samples back; reproduced */
        wasted = (int)g.unary();
        if (g.eof) return FLACB_E_EOF;
    }
    const int ss = sample_size - wasted;
    if (type == 0) { /* constant */
        if (ss <= 0 || ss > 40) return FLACB_E_VALUE;
        const i64 v = g.read_signed(ss);
        if (g.eof) return FLACB_E_EOF;
        for (int i = 0; i < bs; i++) out[i] = (int32_t)v;
        return FLACB_OK;
    }
    if (type == 1) { /* verbatim */
        if (ss <= 0 || ss > 40) return FLACB_E_VALUE;
        for (int i = 0; i < bs; i++) {
            out[i] = (int32_t)g.read_signed(ss);
            if (g.eof) return FLACB_E_EOF;
        }
        return FLACB_OK;
    }
    History hist;
    hist.head = 0;
    int32_t coef[FLACB_MAX_LPC];
    int order, shift = 0;
    if (type <= 12) { /* fixed, common.py:15-21 */
        order = type & 7;
        const int32_t fc[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
        for (int j = 0; j < 4; j++) coef[j] = fc[order][j];
    } else order = (type & 31) + 1;
    if (ss <= 0 || ss > 40) {
        if (order > 0) return FLACB_E_VALUE;
    }
    for (int i = 0; i < order; i++) {
        if (i >= bs) return FLACB_E_ASSERT; /* warm-up longer than the block */
        const i64 v = g.read_signed(ss);
        if (g.eof) return FLACB_E_EOF;
        hist.push(v);
        out[i] = (int32_t)v;
    }
    if (type >= 32) {
        int precision = (int)g.read(4);
        if (g.eof) return FLACB_E_EOF;
        if (precision == 15) return FLACB_E_ASSERT;
        precision += 1;
        shift = (int)g.read_signed(5);
        for (int i = 0; i < order; i++) coef[i] = (int32_t)g.read_signed(precision);
        if (g.eof) return FLACB_E_EOF;
        /* a negative shift only fails once the prediction runs (decoder.py:497); the residual is
         * parsed first */
    }
    const int rc = decode_residual(g, bs, order, coef, shift < 0 ? 0 : shift, hist, out, type <= 12);
    if (rc) return rc;
    if (shift < 0) return FLACB_E_VALUE;
    return FLACB_OK;
}

/* One WARP per candidate, lane 0 does the work: Rice parsing is a serial, data dependent walk
 * through the bit stream, and 32 lanes walking 32 different frames diverge at every code (the
 * first version did that and ran 30x slower).  Parallelism comes from the frames: thousands of
 * independent warps, one per SM sub-partition slot. */
constexpr int DECODE_WARPS = 4;

__global__ void __launch_bounds__(32 * DECODE_WARPS) decode_frames_kernel(const uint8_t *s, u64 len, Candidate *cand,
                                                                         uint32_t n_cand, const u64 *tmp_off,
                                                                         int32_t *tmp, int filter)
{
    const uint32_t ci = blockIdx.x * DECODE_WARPS + (threadIdx.x >> 5);
    if (ci >= n_cand || (threadIdx.x & 31) != 0) return;
    Candidate c = cand[ci];
    c.tmp_off = tmp_off[ci];
    if (filter && ci != 0 && c.status == FLACB_OK) {
        /* cand[0] is only read here (its thread writes it back at the very end, same values
         * for these fields) */
        const int bs0 = cand[0].bs, nch0 = cand[0].nch, st0 = cand[0].status;
        const u64 start0 = cand[0].start;
        if (start0 == 0 && (st0 == FLACB_OK || st0 == STATUS_DEFERRED) && bs0 > 0 && (c.bs > bs0 || c.nch != nch0))
            c.status = STATUS_DEFERRED;
    }
    if (c.status == FLACB_OK) {
        BitReader g;
        g.init(s, len, c.start);
        /* skip the header (already validated) */
        uint32_t hb = c.hdr_bits;
        while (hb >= 32) (void)g.read(32), hb -= 32;
        (void)g.read((int)hb);
        int rc = FLACB_OK;
        for (int ch = 0; ch < c.nch && rc == FLACB_OK; ch++) {
            /* the side channel carries one more bit (common.py:182-196) */
            const int extra = (c.chcode == 8 && ch == 1) || (c.chcode == 9 && ch == 0) || (c.chcode == 10 && ch == 1);
            rc = decode_subframe(g, c.bs, c.sample_size + extra, tmp + c.tmp_off + (u64)ch * (u64)c.bs);
        }
        if (rc == FLACB_OK) {
            const int padn = (int)((8 - (g.bit_pos() & 7)) & 7); /* decoder.py:124-126 */
            const uint32_t pad = g.read(padn);
            if (g.eof) rc = FLACB_E_EOF;
            else if (pad != 0) rc = FLACB_E_ASSERT;
        }
        if (rc == FLACB_OK) {
            (void)g.read(16); /* CRC-16, never verified (decoder.py:128) */
            if (g.eof) rc = FLACB_E_EOF;
        }
        c.status = rc;
        c.end = rc == FLACB_OK ? g.bit_pos() >> 3 : 0;
    }
    cand[ci] = c;
}

/* ---- D4 ------------------------------------------------------------------------------ */

/* Single CTA.  next[c] by binary search, reachability from candidate 0 by pointer jumping,
 * then, in chain order, output offsets and the status the reference would end with. */
__global__ void __launch_bounds__(1024) chain_kernel(Candidate *cand, uint32_t n, u64 len, uint32_t *jump,
                                                     uint32_t *reach, DecodeResult *res)
{
    __shared__ u64 warp_tot[32];
    __shared__ u64 carry;
    __shared__ uint32_t first_bad, first_break, last_ok1;
    __shared__ int changed;
    const uint32_t NONE = 0xFFFFFFFFu;
    const int tid = threadIdx.x, T = blockDim.x;
    if (tid == 0) {
        res->n_frames = 0;
        res->n_sample_frames = 0;
        res->status = FLACB_OK;
        res->frame_channels = 0;
        res->bad_frame = 0;
    }
    if (n == 0) return;
    for (uint32_t c = tid; c < n; c += T) {
        uint32_t nx = NONE;
        const Candidate cc = cand[c];
        if (cc.status == FLACB_OK && cc.end > cc.start) {
            uint32_t lo = c + 1, hi = n; /* first candidate with start >= end */
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if (cand[mid].start < cc.end) lo = mid + 1;
                else hi = mid;
            }
            if (lo < n && cand[lo].start == cc.end) nx = lo;
        }
        cand[c].next = nx;
        jump[c] = nx;
        reach[c] = (c == 0 && cc.start == 0) ? 1u : 0u;
    }
    __syncthreads();
    /* round r links every candidate to the one 2^r frames further down its chain, so the set
     * reached from candidate 0 doubles per round; a round without news means it is complete */
    uint32_t *cur = jump, *nxt = jump + n;
    for (int round = 0; round < 32; round++) {
        if (tid == 0) changed = 0;
        __syncthreads();
        for (uint32_t c = tid; c < n; c += T) {
            const uint32_t j = cur[c];
            if (reach[c] && j != NONE && !reach[j]) {
                reach[j] = 1;
                changed = 1;
            }
        }
        __syncthreads();
        for (uint32_t c = tid; c < n; c += T) {
            const uint32_t j = cur[c];
            nxt[c] = j == NONE ? NONE : cur[j];
        }
        const int any = changed;
        __syncthreads();
        uint32_t *t = cur;
        cur = nxt;
        nxt = t;
        if (!any) break;
    }
    /* first chain frame the reference fails in, first chain frame after which sync is lost */
    if (tid == 0) first_bad = NONE, first_break = NONE, last_ok1 = 0;
    __syncthreads();
    for (uint32_t c = tid; c < n; c += T) {
        if (!reach[c]) continue;
        const Candidate cc = cand[c];
        if (cc.status != FLACB_OK) atomicMin(&first_bad, c);
        else if (cc.next == NONE && cc.end + 1 < len) atomicMin(&first_break, c);
    }
    /* a chain link can also point at a candidate whose own parse failed: covered by first_bad */
    __syncthreads();
    const uint32_t stop = first_bad; /* frames with index >= stop are not produced */
    /* exclusive scan of block sizes over chain frames before `stop` */
    if (tid == 0) carry = 0;
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    for (uint32_t base = 0; base < n; base += T) {
        const uint32_t c = base + tid;
        const bool ok = c < n && reach[c] && c < stop && cand[c].status == FLACB_OK &&
                        (first_break == NONE || c <= first_break);
        const u64 v = ok ? (u64)cand[c].bs : 0;
        u64 inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        u64 wb = 0, tot = 0;
        for (int w = 0; w < nw; w++) {
            if (w < warp) wb += warp_tot[w];
            tot += warp_tot[w];
        }
        if (c < n) {
            cand[c].on_chain = ok ? 1 : 0;
            cand[c].out_off = carry + wb + inc - v;
        }
        if (ok) {
            atomicAdd(&res->n_frames, 1ull);
            atomicMax(&last_ok1, c + 1u); /* index + 1 of the last frame produced */
        }
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid == 0) {
        res->n_sample_frames = carry;
        int status = FLACB_OK;
        if (!reach[0]) {
            /* the stream does not start with a frame: decoder.py:137 sync assert, unless there
             * are fewer than two bytes at all */
            status = len >= 2 ? FLACB_E_ASSERT : FLACB_OK;
        } else if (first_bad != NONE && (first_break == NONE || first_bad <= first_break)) {
            const int s = cand[first_bad].status;
            status = s == FLACB_E_EOF ? FLACB_OK : s; /* truncated last frame: dropped silently */
            res->bad_frame = first_bad;
        } else if (first_break != NONE) {
            status = FLACB_E_ASSERT; /* bytes follow the frame but not a sync code */
            res->bad_frame = first_break;
        }
        res->status = status;
    }
    __syncthreads();
    /* channel count of the last frame produced (what decoder.py's zip(*channels) yields) */
    if (tid == 0) res->frame_channels = last_ok1 ? cand[last_ok1 - 1].nch : 0;
}

/* ---- D5 ------------------------------------------------------------------------------ */

__global__ void __launch_bounds__(256) interleave_kernel(const Candidate *cand, const int32_t *tmp, int32_t *out,
                                                         u64 cap_samples, int out_channels)
{
    const Candidate c = cand[blockIdx.x];
    if (!c.on_chain || c.nch != out_channels) return;
    const u64 total = (u64)c.bs * (u64)c.nch;
    if ((c.out_off + (u64)c.bs) * (u64)c.nch > cap_samples) return;
    const int32_t *src = tmp + c.tmp_off;
    int32_t *dst = out + c.out_off * (u64)c.nch;
    if (c.chcode < 8) {
        for (u64 idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const u64 i = idx / (u64)c.nch, ch = idx - i * (u64)c.nch;
            dst[idx] = src[ch * (u64)c.bs + i];
        }
    } else {
        /* decoder.py:436-449 */
        for (int i = threadIdx.x; i < c.bs; i += blockDim.x) {
            const i64 c0 = src[i], c1 = src[c.bs + i];
            i64 l, r;
            if (c.chcode == 8) l = c0, r = c0 - c1;          /* L_S */
            else if (c.chcode == 9) l = c0 + c1, r = c1;     /* S_R */
            else r = c0 - (c1 >> 1), l = r + c1;             /* M_S as the reference computes it */
            dst[2 * (u64)i] = (int32_t)l;
            dst[2 * (u64)i + 1] = (int32_t)r;
        }
    }
}

#endif /* !FLACB_HOST_ONLY */

} /* namespace flacb */
