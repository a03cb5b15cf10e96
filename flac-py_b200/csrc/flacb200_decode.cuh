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
 *   D3  decode_frames_kernel  one thread per candidate: subframe headers, warm-up samples and
 *                             Rice residuals (decoder.py:267-421) into a planar scratch area;
 *                             records where the frame ends and what every subframe needs for
 *       restore_kernel        the serial prediction restore (decoder.py:453-498): one thread per
 *                             subframe, history and coefficients in registers
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

/* what restore_kernel needs to know about one subframe (slot candidate * 8 + channel) */
struct SubRec {
    u64 off;        /* first sample of the subframe in the planar scratch area */
    int32_t bs;
    int32_t kind;   /* 0: nothing to restore (constant, verbatim, not decoded), 1: fixed, 2: LPC */
    int32_t order, shift;
    int32_t coef[FLACB_MAX_LPC];
};
constexpr int SUBREC_SLOTS = 8; /* subframes per candidate */

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
    uint32_t pre; /* the word at `next`, loaded ahead of its use (state-machine decoder) */
    bool pre_ok;

    __device__ __forceinline__ void init(const uint8_t *s, u64 length, u64 byte_pos)
    {
        p = s, len = length, next = byte_pos, buf = 0, avail = 0, eof = false, pre = 0, pre_ok = false;
    }
    __device__ __forceinline__ void refill()
    {
        pre_ok = false; /* reloads from memory whatever had been fetched ahead */
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

/* One LANE per candidate frame, 32 frames per warp.  Rice parsing is a serial, data dependent
 * walk through the bit stream; 32 lanes walking 32 different frames with ordinary nested loops
 * diverge at every code and end up serialised (the first version: 30x slower).  So the frame
 * decode is written as a state machine: every lane is in one of a few states and a warp-wide
 * loop advances all lanes by one step per iteration.  Practically all iterations find all lanes
 * in ST_RICE (one Rice code per step, at most one refill per step), so the warp stays
 * converged; header fields, warm-up samples, partition parameters are rare steps. */

enum {
    ST_SUB_HDR = 0, /* subframe header byte, wasted bits (decoder.py:318-355) */
    ST_WASTED,      /* unary wasted-bits count in progress */
    ST_CONST,       /* constant subframe value */
    ST_VERBATIM,    /* one verbatim sample per step */
    ST_WARMUP,      /* one warm-up sample per step */
    ST_LPC_HDR,     /* precision, shift */
    ST_COEF,        /* one coefficient per step */
    ST_RES_HDR,     /* coding method, partition order (decoder.py:358-384) */
    ST_PART_HDR,    /* Rice parameter / escape size of the next partition */
    ST_RICE,        /* one Rice coded residual per step (decoder.py:414-421) */
    ST_ESCAPED,     /* one escaped (raw) residual per step (decoder.py:404-407) */
    ST_FRAME_END,   /* padding + CRC-16 */
    ST_DONE
};

struct LaneState {
    BitReader g;
    int state, status;
    int ch, nch, chcode, bs, sample_size; /* frame */
    int ss, type, order, shift, precision; /* subframe */
    int pos, idx;                         /* sample position in the subframe; warm-up / coef index */
    int method, porder, parts_left, part_left, k, esc_size;
    u64 cnt;                              /* zeros of an unfinished unary run */
    int32_t *out;                         /* this channel's plane in the scratch area */
};

__device__ __forceinline__ void lane_fail(LaneState &st, int status)
{
    st.status = status;
    st.state = ST_DONE;
}

/* word-wise refill: keeps 33..64 bits buffered; bytes only in front of the first aligned word and
 * behind the last one.  At most one load per call on the fast path. */
__device__ __forceinline__ void lane_refill(BitReader &g)
{
    if (g.avail <= 32) {
        if (g.pre_ok) {
            g.buf |= (u64)__byte_perm(g.pre, 0, 0x0123) << (32 - g.avail);
            g.avail += 32;
            g.next += 4;
            g.pre_ok = false;
        } else {
            g.refill(); /* ends of the stream, unaligned start */
        }
    }
    /* fetch the next word now; it is only needed a step or two later, so its latency overlaps the
     * decoding of the bits already buffered */
    if (!g.pre_ok && g.next + 4 <= g.len && (((uintptr_t)(g.p + g.next)) & 3) == 0) {
        g.pre = *(const uint32_t *)(g.p + g.next);
        g.pre_ok = true;
    }
}

__device__ __forceinline__ void lane_store(LaneState &st, i64 r)
{
    /* the residual only; restore_kernel adds the prediction (decoder.py:493-498).  The low 32 bits
     * are all it needs: the sums wrap modulo 2^32, which leaves the stored int32 samples exact */
    st.out[st.pos++] = (int32_t)r;
    if (--st.part_left == 0) st.state = st.parts_left > 0 ? ST_PART_HDR : ST_SUB_HDR + 100; /* 100+: subframe done */
}

__device__ __forceinline__ void lane_step(LaneState &st, SubRec *recs, int32_t *tmp, u64 tmp_off)
{
    BitReader &g = st.g;
    switch (st.state) {
    case ST_RICE: {
        lane_refill(g);
        if (g.buf == 0) { /* the unary run continues behind the buffered bits */
            if (g.avail == 0 && g.next >= g.len) return lane_fail(st, FLACB_E_EOF);
            st.cnt += (u64)g.avail;
            g.avail = 0;
            return;
        }
        const int z = __clzll(g.buf);
        const u64 msb = st.cnt + (u64)z;
        st.cnt = 0;
        g.buf <<= z;
        g.buf <<= 1;
        g.avail -= z + 1;
        lane_refill(g);
        const u64 lsb = g.read(st.k);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        const u64 u = (msb << st.k) | lsb;
        lane_store(st, (i64)(u >> 1) ^ -(i64)(u & 1)); /* utils.py:87-88 */
        break;
    }
    case ST_ESCAPED: {
        lane_refill(g);
        const i64 r = g.read_signed(st.esc_size);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        lane_store(st, r);
        break;
    }
    case ST_PART_HDR: {
        lane_refill(g);
        const int first = st.parts_left == (1 << st.porder);
        st.part_left = first ? (st.bs >> st.porder) - st.order : (st.bs >> st.porder);
        st.parts_left--;
        st.k = (int)g.read(st.method);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        if (st.k == (1 << st.method) - 1) { /* escaped partition */
            st.esc_size = (int)g.read(5);
            if (g.eof) return lane_fail(st, FLACB_E_EOF);
            if (st.esc_size == 0 && st.part_left > 0) return lane_fail(st, FLACB_E_VALUE); /* sint(0) */
            st.state = ST_ESCAPED;
        } else st.state = ST_RICE;
        if (st.part_left == 0) st.state = st.parts_left > 0 ? ST_PART_HDR : ST_SUB_HDR + 100;
        break;
    }
    case ST_RES_HDR: {
        lane_refill(g);
        const uint32_t m = g.read(2);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        if (m > 1) return lane_fail(st, FLACB_E_VALUE); /* RiceCodingMethod(x) */
        st.method = m == 0 ? 4 : 5;
        st.porder = (int)g.read(4);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        const int nparts = 1 << st.porder;
        if (st.bs % nparts != 0) return lane_fail(st, FLACB_E_ASSERT);
        if (!((st.bs >> st.porder) > st.order)) return lane_fail(st, FLACB_E_ASSERT);
        st.parts_left = nparts;
        st.cnt = 0;
        st.state = ST_PART_HDR;
        break;
    }
    case ST_WARMUP: {
        if (st.idx >= st.order) {
            st.state = st.type >= 32 ? ST_LPC_HDR : ST_RES_HDR;
            break;
        }
        if (st.ss <= 0 || st.ss > 40) return lane_fail(st, FLACB_E_VALUE);
        if (st.idx >= st.bs) return lane_fail(st, FLACB_E_ASSERT); /* warm-up longer than the block */
        lane_refill(g);
        const i64 v = g.read_signed(st.ss);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        st.out[st.idx++] = (int32_t)v;
        st.pos = st.idx;
        break;
    }
    case ST_LPC_HDR: {
        lane_refill(g);
        st.precision = (int)g.read(4);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        if (st.precision == 15) return lane_fail(st, FLACB_E_ASSERT);
        st.precision += 1;
        st.shift = (int)g.read_signed(5);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        st.idx = 0;
        st.state = ST_COEF;
        break;
    }
    case ST_COEF: {
        if (st.idx >= st.order) {
            st.state = ST_RES_HDR;
            break;
        }
        lane_refill(g);
        recs[st.ch < SUBREC_SLOTS ? st.ch : 0].coef[st.idx++] = (int32_t)g.read_signed(st.precision);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        break;
    }
    case ST_SUB_HDR: {
        if (st.ch >= st.nch) {
            st.state = ST_FRAME_END;
            break;
        }
        lane_refill(g);
        const uint32_t padbit = g.read(1);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        if (padbit != 0) return lane_fail(st, FLACB_E_ASSERT);
        st.type = (int)g.read(6);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        if (!(st.type <= 1 || (st.type >= 8 && st.type <= 12) || st.type >= 32)) return lane_fail(st, FLACB_E_ASSERT);
        const uint32_t wflag = g.read(1);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        /* the side channel carries one more bit (common.py:182-196) */
        const int extra = (st.chcode == 8 && st.ch == 1) || (st.chcode == 9 && st.ch == 0) || (st.chcode == 10 && st.ch == 1);
        st.ss = st.sample_size + extra;
        st.out = tmp + tmp_off + (u64)st.ch * (u64)st.bs;
        st.pos = 0, st.idx = 0, st.shift = 0, st.cnt = 0;
        st.order = st.type >= 32 ? (st.type & 31) + 1 : st.type >= 8 ? (st.type & 7) : 0;
        st.state = wflag ? ST_WASTED : (st.type == 0 ? ST_CONST : st.type == 1 ? ST_VERBATIM : ST_WARMUP);
        break;
    }
    case ST_WASTED: {
        /* decoder.py:346-355: counts the zeros (one less than the format says) and never shifts the
         * samples back; reproduced */
        lane_refill(g);
        if (g.buf == 0) {
            if (g.avail == 0 && g.next >= g.len) return lane_fail(st, FLACB_E_EOF);
            st.cnt += (u64)g.avail;
            g.avail = 0;
            return;
        }
        const int z = __clzll(g.buf);
        const u64 wasted = st.cnt + (u64)z;
        st.cnt = 0;
        g.buf <<= z;
        g.buf <<= 1;
        g.avail -= z + 1;
        st.ss -= (int)(wasted > 1000 ? 1000 : wasted);
        st.state = st.type == 0 ? ST_CONST : st.type == 1 ? ST_VERBATIM : ST_WARMUP;
        break;
    }
    case ST_CONST: {
        if (st.ss <= 0 || st.ss > 40) return lane_fail(st, FLACB_E_VALUE);
        lane_refill(g);
        const i64 v = g.read_signed(st.ss);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        for (int i = 0; i < st.bs; i++) st.out[i] = (int32_t)v;
        st.state = ST_SUB_HDR + 100;
        break;
    }
    case ST_VERBATIM: {
        if (st.ss <= 0 || st.ss > 40) return lane_fail(st, FLACB_E_VALUE);
        if (st.pos >= st.bs) {
            st.state = ST_SUB_HDR + 100;
            break;
        }
        lane_refill(g);
        st.out[st.pos++] = (int32_t)g.read_signed(st.ss);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        break;
    }
    case ST_FRAME_END: {
        const int padn = (int)((8 - (g.bit_pos() & 7)) & 7); /* decoder.py:124-126 */
        lane_refill(g);
        const uint32_t pad = g.read(padn);
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        if (pad != 0) return lane_fail(st, FLACB_E_ASSERT);
        lane_refill(g);
        (void)g.read(16); /* CRC-16, never verified (decoder.py:128) */
        if (g.eof) return lane_fail(st, FLACB_E_EOF);
        st.status = FLACB_OK;
        st.state = ST_DONE;
        break;
    }
    default:
        break;
    }
    if (st.state == ST_SUB_HDR + 100) { /* subframe complete */
        if (st.type >= 32 && st.shift < 0) return lane_fail(st, FLACB_E_VALUE); /* >> negative, decoder.py:497 */
        if (st.ch < SUBREC_SLOTS) {
            SubRec &r = recs[st.ch];
            r.off = (u64)(st.out - tmp);
            r.bs = st.bs;
            r.order = st.order;
            r.shift = st.shift;
            r.kind = st.type >= 32 ? 2 : (st.type >= 8 && st.order > 0) ? 1 : 0;
        }
        st.ch++;
        st.state = ST_SUB_HDR;
    }
}

/* The hot loop: Rice codes of the current partition for as long as nothing unusual happens (whole
 * aligned words left in the stream, no unary run of 32 or more zeros, parameter below 27).  The
 * bit window is two big-endian 32-bit words and a cursor, fields are cut out with funnel shifts,
 * the word behind the window is fetched a code or two before it is needed: about 30 32-bit
 * instructions per code, the same for every lane, so a warp's lanes stay converged here.
 * Everything else (partition and subframe headers, warm-up, the ends of the stream, escapes,
 * long zero runs) is a step of the general state machine above. */
constexpr int RICE_BURST = 32;

#ifdef FLACB_CPU_SIM
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (shift & 31)) >> 32);
}
#endif

__device__ __forceinline__ void rice_fast(LaneState &st)
{
    BitReader &g = st.g;
    const int k = st.k;
    if (st.cnt != 0 || k > 26) return;
    /* absolute bit position -> aligned word and offset inside it */
    const u64 bitpos = g.next * 8 - (u64)g.avail;
    const uintptr_t addr = (uintptr_t)g.p + (uintptr_t)(bitpos >> 3);
    const uint32_t *wp = (const uint32_t *)(addr & ~(uintptr_t)3);
    const int64_t wbyte = (int64_t)((uintptr_t)wp - (uintptr_t)g.p); /* byte offset of the window in the stream */
    if (wbyte < 0 || (u64)wbyte + 12 > g.len) return;
    /* an iteration moves the window by two words at most; three words are read ahead */
    const u64 words_left = (g.len - (u64)wbyte) / 4;
    int n = st.part_left < RICE_BURST ? st.part_left : RICE_BURST;
    if ((u64)n > (words_left - 3) / 2) n = (int)((words_left - 3) / 2);
    if (n <= 0) return;
    uint32_t o = (uint32_t)(bitpos - (u64)wbyte * 8);
    /* The word fetched ahead is used one refill later (about a code): enough for a line that is in L1, not for
     * one that has to come from L2 or HBM -- and with several lanes per warp some lane starts a new line in
     * every third iteration, and the whole warp waits (28 % of the kernel's stall samples).  So the lines a
     * burst or two ahead are asked for now. */
#ifndef FLACB_CPU_SIM
    if ((u64)wbyte + 512 <= g.len) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"((const uint8_t *)wp + 128));
        asm volatile("prefetch.global.L1 [%0];" ::"l"((const uint8_t *)wp + 256));
    }
#endif
    uint32_t w0 = __byte_perm(wp[0], 0, 0x0123), w1 = __byte_perm(wp[1], 0, 0x0123), pre = wp[2];
    wp += 3;
    int32_t *out = st.out + st.pos;
    int done = 0;
    for (; done < n; done++) {
        const uint32_t top = __funnelshift_l(w1, w0, o);
        if (top == 0) break; /* 32 zeros or more: the general path counts them */
        const uint32_t z = (uint32_t)__clz((int)top);
        o += z + 1;
        if (o >= 32) {
            w0 = w1, w1 = __byte_perm(pre, 0, 0x0123), pre = *wp++;
            o -= 32;
        }
        const uint32_t lsb = (__funnelshift_l(w1, w0, o) >> 1) >> (31 - k);
        o += (uint32_t)k;
        if (o >= 32) {
            w0 = w1, w1 = __byte_perm(pre, 0, 0x0123), pre = *wp++;
            o -= 32;
        }
        const uint32_t u = (z << k) | lsb; /* z < 32, k < 27, and z << k < 2^31 ... fits */
        out[done] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); /* utils.py:87-88 */
    }
    if (done == 0) return;
    /* back to the reader's form: the window's remaining bits, left aligned; next = the word behind it */
    g.buf = (((u64)w0 << 32) | (u64)w1) << o;
    g.avail = 64 - (int)o;
    g.next = (u64)((uintptr_t)wp - (uintptr_t)g.p) - 4; /* wp is one behind `pre` */
    g.pre_ok = false;
    st.pos += done;
    st.part_left -= done;
    if (st.part_left == 0) st.state = st.parts_left > 0 ? ST_PART_HDR : ST_SUB_HDR + 100;
}

/* frames per warp: fewer than 32 trades idle lanes for less divergence between the lanes' states
 * and more warps per SM (the kernel is bound by the latency of one lane's instruction stream:
 * refill, count zeros, shift, read k bits, store) */
constexpr int DECODE_LANES = 8;  /* the simulator's value; the library picks 2..32 from the number of candidates (flacb200.cu) */
constexpr int DECODE_WARPS = 4;  /* warps per CTA */

__global__ void __launch_bounds__(32 * DECODE_WARPS) decode_frames_kernel(const uint8_t *s, u64 len, Candidate *cand, uint32_t n_cand,
                                                                          const u64 *tmp_off, int32_t *tmp, SubRec *recs,
                                                                          int filter, int lanes)
{
    const int lane = (int)threadIdx.x & 31;
    const uint32_t wid = blockIdx.x * (uint32_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t ci = lane < lanes && (u64)wid * (u64)lanes + (u64)lane < (u64)n_cand ? wid * (uint32_t)lanes + (uint32_t)lane : n_cand;
    LaneState st;
    st.state = ST_DONE;
    st.status = FLACB_E_EOF;
    Candidate c;
    u64 my_off = 0;
    if (ci < n_cand) {
        c = cand[ci];
        my_off = tmp_off[ci];
        c.tmp_off = my_off;
        if (filter && ci != 0 && c.status == FLACB_OK) {
            /* cand[0] is only read here (its lane writes it back at the very end, same values
             * for these fields) */
            const int bs0 = cand[0].bs, nch0 = cand[0].nch, st0 = cand[0].status;
            const u64 start0 = cand[0].start;
            if (start0 == 0 && (st0 == FLACB_OK || st0 == STATUS_DEFERRED) && bs0 > 0 && (c.bs > bs0 || c.nch != nch0))
                c.status = STATUS_DEFERRED;
        }
        if (c.status == FLACB_OK) {
            st.g.init(s, len, c.start + c.hdr_bits / 8); /* the header is byte aligned and already parsed */
            st.state = ST_SUB_HDR;
            st.ch = 0, st.nch = c.nch, st.chcode = c.chcode, st.bs = c.bs, st.sample_size = c.sample_size;
            st.type = 0, st.order = 0, st.shift = 0, st.cnt = 0;
        }
    }
    const bool active = ci < n_cand && c.status == FLACB_OK;
    SubRec *my_recs = recs + (u64)(ci < n_cand ? ci : 0) * SUBREC_SLOTS;
    while (__any_sync(0xffffffffu, st.state != ST_DONE)) {
        bool general = st.state != ST_DONE;
        if (st.state == ST_RICE) {
            const int before = st.part_left;
            rice_fast(st);
            general = st.state != ST_RICE || st.part_left == before; /* no progress: the general path takes the code */
        }
        if (general) lane_step(st, my_recs, tmp, my_off);
    }
    if (ci < n_cand) {
        if (active) {
            c.status = st.status;
            c.end = st.status == FLACB_OK ? st.g.bit_pos() >> 3 : 0;
        }
        cand[ci] = c;
    }
}

/* ---- prediction restore (decoder.py:453-498 _decode_prediction) -------------------------- */

/* result[i] += (sum_j result[i-1-j] * c[j]) >> shift, serially: one thread per subframe, the last
 * W samples and the coefficients in registers (W >= order, coefficients beyond the order are
 * zero).  The term of the newest sample is added last, so the chain from one sample to the next
 * is short.  Residuals are read eight at a time, several chunks ahead of the arithmetic (the plane is
 * rewritten in place). */
constexpr int RESTORE_CHUNK = 8;
#ifndef FLACB_RESTORE_AHEAD
#define FLACB_RESTORE_AHEAD 4
#endif

template <int W>
__device__ __forceinline__ void restore_run(int32_t *x, int bs, int order, int shift, const int (&ci)[W])
{
    /* The sum of up to 32 products of a 32-bit sample and a 15-bit coefficient is an integer below 2^51: fp64
     * multiply-adds carry it exactly, one instruction per tap (the 64-bit integer form costs three), and the chain
     * from one sample to the next is convert, multiply-add, scale, floor-convert, add. */
    double c[W], h[W]; /* h[j] = x[i - 1 - j] */
#pragma unroll
    for (int j = 0; j < W; j++) c[j] = (double)ci[j], h[j] = j < order ? (double)x[order - 1 - j] : 0.0;
    const double scale = __longlong_as_double((long long)(1023 - shift) << 52); /* 2^-shift */
    /* residuals are fetched RESTORE_AHEAD chunks (about 1500 cycles of arithmetic) before their turn: the planes
     * come from DRAM, every lane from a line of its own, and the chunk behind a store would otherwise wait there */
    constexpr int RESTORE_AHEAD = W > 16 ? 2 : FLACB_RESTORE_AHEAD; /* the 32-tap instance has no registers to spare */
    int q[RESTORE_AHEAD][RESTORE_CHUNK];
#pragma unroll
    for (int k = 0; k < RESTORE_AHEAD; k++)
#pragma unroll
        for (int e = 0; e < RESTORE_CHUNK; e++) q[k][e] = order + k * RESTORE_CHUNK + e < bs ? x[order + k * RESTORE_CHUNK + e] : 0;
    for (int i0 = order; i0 < bs; i0 += RESTORE_CHUNK * RESTORE_AHEAD) {
#pragma unroll
        for (int k = 0; k < RESTORE_AHEAD; k++) {
            const int ik = i0 + k * RESTORE_CHUNK;
            if (ik >= bs) break;
            int cur[RESTORE_CHUNK];
#pragma unroll
            for (int e = 0; e < RESTORE_CHUNK; e++) cur[e] = q[k][e];
#pragma unroll
            for (int e = 0; e < RESTORE_CHUNK; e++) {
                const int in = ik + RESTORE_CHUNK * RESTORE_AHEAD + e;
                q[k][e] = in < bs ? x[in] : 0;
            }
#pragma unroll
            for (int e = 0; e < RESTORE_CHUNK; e++) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int j = W - 1; j >= 1; j--) {
                    if (j & 1) a1 = __fma_rn(h[j], c[j], a1);
                    else a0 = __fma_rn(h[j], c[j], a0);
                }
                const double acc = __fma_rn(h[0], c[0], __dadd_rn(a0, a1));
                const int32_t v = (int32_t)((i64)cur[e] + __double2ll_rd(__dmul_rn(acc, scale)));
#pragma unroll
                for (int j = W - 1; j > 0; j--) h[j] = h[j - 1];
                h[0] = (double)v;
                cur[e] = v;
            }
#pragma unroll
            for (int e = 0; e < RESTORE_CHUNK; e++)
                if (ik + e < bs) x[ik + e] = cur[e];
        }
    }
}

template <int W>
__device__ __forceinline__ void restore_any(int32_t *x, const SubRec &r)
{
    int c[W];
    int shift = 0;
    if (r.kind == 1) {
        /* the fixed predictors are LPC filters with shift 0 (common.py:15-21); their sums wrap
         * modulo 2^32 like the int32 samples they are stored in */
        const int tab[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
#pragma unroll
        for (int j = 0; j < W; j++) c[j] = j < 4 ? tab[r.order][j] : 0;
    } else {
#pragma unroll
        for (int j = 0; j < W; j++) c[j] = j < r.order ? r.coef[j] : 0;
        shift = r.shift < 0 ? 0 : r.shift;
    }
    restore_run<W>(x, r.bs, r.order, shift, c);
}

/* thread t: channel t / n_cand of candidate t % n_cand, so that a warp's lanes hold subframes of
 * the same channel of neighbouring candidates: all of them real for the stream's channels, none
 * for the rest (those warps leave at once) */
__global__ void __launch_bounds__(64) restore_kernel(const SubRec *recs, uint32_t n_cand, int32_t *tmp)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = t < (u64)n_cand * SUBREC_SLOTS;
    const uint32_t ch = (uint32_t)(t / n_cand), ci = (uint32_t)(t - (u64)ch * n_cand);
    const SubRec &r = recs[in_range ? (u64)ci * SUBREC_SLOTS + ch : 0];
    const bool work = in_range && r.kind != 0;
    /* one filter length for the whole warp (the longest order among its subframes, the others' taps zero):
     * lanes that took different lengths ran one after the other, each with the full cost of moving the samples */
#ifdef FLACB_RESTORE_DIVERGENT
    const int wmax = work ? r.order : 0;
#else
    const int wmax = (int)__reduce_max_sync(0xffffffffu, work ? (unsigned)r.order : 0u);
#endif
    if (!work) return;
    int32_t *x = tmp + r.off;
    if (wmax <= 4) restore_any<4>(x, r); /* every fixed predictor */
    else if (wmax <= 8) restore_any<8>(x, r);
    else if (wmax <= 12) restore_any<12>(x, r);
    else if (wmax <= 16) restore_any<16>(x, r);
    else restore_any<32>(x, r);
}

/* ---- D4 ------------------------------------------------------------------------------ */

/* Single CTA.  next[c] by binary search, reachability from candidate 0 by pointer jumping,
 * then, in chain order, output offsets and the status the reference would end with. */
__global__ void __launch_bounds__(1024) chain_kernel(Candidate *cand, uint32_t n, u64 len, uint32_t *jump,
                                                     uint32_t *reach, DecodeResult *res)
{
    __shared__ u64 warp_tot[32];
    __shared__ u64 carry;
    __shared__ uint32_t first_bad, first_break, first_change, last_ok1;
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
    if (tid == 0) first_bad = NONE, first_break = NONE, first_change = NONE, last_ok1 = 0;
    __syncthreads();
    const int nch0 = cand[0].nch;
    for (uint32_t c = tid; c < n; c += T) {
        if (!reach[c]) continue;
        const Candidate cc = cand[c];
        if (cc.status == FLACB_OK && cc.nch != nch0) atomicMin(&first_change, c);
        if (cc.status != FLACB_OK) atomicMin(&first_bad, c);
        else if (cc.next == NONE && cc.end + 1 < len) atomicMin(&first_break, c);
    }
    /* a chain link can also point at a candidate whose own parse failed: covered by first_bad */
    __syncthreads();
    /* The output is one interleaved array: a chain frame with another channel count than frame 0
     * (the reference yields rows of another width from there on, decoder.py:52-55) ends the
     * decode with FLACB_E_UNSUPPORTED; the frames in front of it are produced. */
    const bool change_stop = first_change < first_bad && (first_break == NONE || first_change <= first_break);
    const uint32_t stop = change_stop ? first_change : first_bad; /* frames with index >= stop are not produced */
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
        } else if (change_stop) {
            status = FLACB_E_UNSUPPORTED;
            res->bad_frame = first_change;
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
