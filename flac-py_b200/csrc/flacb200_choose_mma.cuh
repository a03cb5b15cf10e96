/*
 * flacb200_choose_mma.cuh -- stage C1 on the tensor cores: choose_mma_kernel.
 *
 * The reference evaluates EVERY LPC order on EVERY sample (flac/encoder.py:387-398 over
 * prediction_residual, encoder.py:537-548): L (L + 1) / 2 multiply-adds per sample, 528 at
 * L = 32.  For one block of samples that is a matrix product: with X[i][k] = x[i - 32 + k]
 * (the 32 samples in front of sample i, a Hankel matrix) and Q[k][m] = q_m[31 - k] (the quantised
 * coefficients of order m + 1, zero beyond the order), P = X Q holds the prediction sum of
 * every order for every sample.  Samples of up to 24 bits are split into byte planes
 * (x = hi 2^16 + mid 2^8 + lo, hi signed, mid and lo unsigned), coefficients of up to 8 bits
 * are bytes, and each plane goes through `mma.sync.m16n8k32` with 32-bit integer accumulation
 * (|sum| <= 32 * 255 * 128 < 2^21: exact); the planes are put together again afterwards.  A warp
 * takes 16 samples x 32 orders x 32 taps per step: 12 tensor-core instructions where the
 * multiply-add pipe needed 16.5 warp instructions PER SAMPLE.
 *
 * This is integer work on the legacy tensor-core path (mma.sync; no tcgen05 / TMEM: the tiles are
 * tiny, the operands are built from bytes in shared memory and the result is consumed at once by
 * the |residual| epilogue, which is now the larger part: shift, subtract, absolute value and
 * accumulate per sample and order).  Results are bit-identical to the scalar flavours.
 *
 * Used for samples wider than 16 bits with coefficients of at most 8 bits (configs 3 and 4);
 * FLACB200_FLAVOUR=3 forces it for 16-bit samples as well (two planes).
 */
#pragma once

#include "flacb200_kernels.cuh"

namespace flacb {

constexpr int MMA_GB = 48;   /* zero bytes in front of a byte plane (32 taps of history + alignment slack), a multiple of 16 */
constexpr int MMA_PADB = 48; /* bytes behind it */

struct MmaLayout {
    uint32_t bpl, bstride; /* byte planes of one channel: plane p at bpl + p * bstride */
    uint32_t red, qb, lpc, lpc_stride, total; /* qb: the current channel's coefficient bytes, [32 orders][32 taps] */
    /* what the kernel keeps of a subframe's record (no Levinson matrix): offsets inside its lpc_stride bytes */
    uint32_t c_autoc, c_qc, c_qmeta, c_fixed;
    int planes;
};

__host__ __device__ inline MmaLayout make_layout_choose_mma(int nmax, int channels, int bps, int rmax, int L, int threads)
{
    const SmemLayout rec = make_layout(nmax, channels, bps, rmax, L, threads, 0);
    MmaLayout l;
    l.planes = bps <= 8 ? 1 : bps <= 16 ? 2 : 3;
    l.bstride = align16((uint32_t)(MMA_GB + nmax + MMA_PADB));
    uint32_t off = 0;
    l.bpl = off; off += l.bstride * (uint32_t)l.planes;
    const uint32_t thr = (uint32_t)((threads + 31) / 32 * 32);
    l.red = off; off += align16((thr / 32) * 32u * 8u);
    l.qb = off; off += 32u * 32u;
    l.c_autoc = 0;
    l.c_qc = rec.o_levprod - rec.o_autoc;
    l.c_qmeta = l.c_qc + (rec.o_qmeta - rec.o_qc);
    l.c_fixed = l.c_qmeta + (rec.o_fixed - rec.o_qmeta);
    l.lpc_stride = align16(l.c_fixed + 64);
    l.lpc = off; off += l.lpc_stride * (uint32_t)channels;
    l.total = off;
    return l;
}

#ifndef FLACB_HOST_ONLY

/* D = A (16 x 32, row major, bytes) * B (32 x 8, column major, signed bytes) + C, 32-bit sums.
 * Fragments as PTX defines them for mma.m16n8k32 with 8-bit operands: with g = lane / 4, t = lane % 4
 *   a0: row g,     k = 4 t .. 4 t + 3        a1: row g + 8, same k
 *   a2: row g,     k = 16 + 4 t .. + 3       a3: row g + 8, same k
 *   b0: k = 4 t .. 4 t + 3, column g         b1: k = 16 + 4 t .. + 3, column g
 *   c0, c1: row g, columns 2 t, 2 t + 1      c2, c3: row g + 8, same columns
 * (lowest byte of a register = lowest k). */
template <bool A_SIGNED>
__device__ __forceinline__ void mma_bytes(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
#ifdef FLACB_CPU_SIM
    /* the simulator: every lane publishes its fragments, then computes its own four results */
    sim::State &s = sim::st();
    const int w = s.cur >> 5, lane = s.cur & 31;
    static uint32_t fa[64][32][4], fb[64][32][2];
    for (int i = 0; i < 4; i++) fa[w][lane][i] = a[i];
    for (int i = 0; i < 2; i++) fb[w][lane][i] = b[i];
    sim::syncwarp();
    const int g = lane >> 2, t = lane & 3;
    auto A = [&](int row, int k) {
        const int l2 = (row & 7) * 4 + ((k & 15) >> 2), reg = (row >> 3) + 2 * (k >> 4);
        const uint32_t byte = (fa[w][l2][reg] >> (8 * (k & 3))) & 0xFFu;
        return A_SIGNED ? (int)(int8_t)byte : (int)byte;
    };
    auto B = [&](int k, int col) {
        const int l2 = col * 4 + ((k & 15) >> 2), reg = k >> 4;
        return (int)(int8_t)((fb[w][l2][reg] >> (8 * (k & 3))) & 0xFFu);
    };
    for (int i = 0; i < 4; i++) {
        const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1);
        int acc = d[i];
        for (int k = 0; k < 32; k++) acc += A(row, k) * B(k, col);
        d[i] = acc;
    }
    sim::syncwarp();
#else
    if (A_SIGNED)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
#endif
}

/* four consecutive bytes of a byte plane starting at byte index i (any alignment) */
__device__ __forceinline__ uint32_t bytes4(const uint8_t *pl, int i)
{
    const uint32_t *w = (const uint32_t *)(pl + (i & ~3));
    return __funnelshift_r(w[0], w[1], 8u * (uint32_t)(i & 3));
}

/* encoder.py:387-398: sum |residual| of every LPC order 1..L on the channel whose byte planes are in
 * shared memory -> sh.lpc_err[0..L).  qb: [32 orders][32] coefficient bytes, qb[m - 1][31 - j] = q_m[j]. */
template <int NP>
__device__ void lpc_errors_mma(Blk &b, const uint8_t *bpl, uint32_t bstride, int L, const uint8_t *qb, u64 *red)
{
    Shared &sh = b.sh;
    const int n = b.n;
    const int g = b.lane >> 2, t = b.lane & 3;
    const int NT = (L + 7) >> 3; /* column tiles of 8 orders */
    const bool short_taps = L <= 16; /* taps 16..31 (k < 16) meet zero coefficients only */
    /* The planes reassemble to a prediction sum of up to 8 NP + 13 bits.  Once shifted it fits 32 bits whenever
     * sum |q| * 2^(bps-1) >> shift stays below 2^30 (every sane filter): the residual is then 32-bit work on the
     * low word of a 64-bit funnel shift.  Anything else takes the full 64-bit path (block-uniform choice). */
    bool fits = true;
    for (int m = 0; m < L; m++)
        if (b.qncoef[m] != 0 && (((u64)b.qabs[m] << (b.p.bps - 1)) >> b.qshift[m]) >= (1ull << 30)) fits = false;
#ifdef FLACB_CPU_SIM
    if (getenv("FLACB_SIM_MMA_WIDE")) fits = false; /* the tests walk the 64-bit path too */
#endif
    /* the thread's order columns: tile nt -> orders nt * 8 + 2 t + {1, 2} */
    u64 acc[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) acc[nt][0] = 0, acc[nt][1] = 0;
    int shift[4][2];
    bool live[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int m1 = nt * 8 + 2 * t + j; /* order - 1 */
            live[nt][j] = m1 < L && b.qncoef[m1 < L ? m1 : 0] != 0;
            shift[nt][j] = m1 < L ? b.qshift[m1] : 0;
        }
    }
    const int ntiles = (n + 15) >> 4;
    for (int tile = b.warp; tile < ntiles; tile += b.nwarps) {
        const int s0 = tile << 4;
        /* A: rows g, g + 8 = samples s0 + g, s0 + g + 8; column k = sample row - 32 + k */
        uint32_t a[NP][4];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const uint8_t *pl = bpl + (uint32_t)p * bstride + MMA_GB;
            const int i0 = s0 + g - 32 + 4 * t;
            a[p][0] = short_taps ? 0u : bytes4(pl, i0);
            a[p][1] = short_taps ? 0u : bytes4(pl, i0 + 8);
            a[p][2] = bytes4(pl, i0 + 16);
            a[p][3] = bytes4(pl, i0 + 24);
        }
        /* the rows' own samples */
        int xr[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int i = s0 + g + 8 * r;
            int v = (int)(int8_t)bpl[(uint32_t)(NP - 1) * bstride + MMA_GB + i];
#pragma unroll
            for (int p = NP - 2; p >= 0; p--) v = (v << 8) | (int)bpl[(uint32_t)p * bstride + MMA_GB + i];
            xr[r] = v;
        }
        const bool interior = s0 >= 32 && s0 + 16 <= n;
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            if (nt < NT) { /* block-uniform */
                uint32_t bf[2];
                const uint32_t *qrow = (const uint32_t *)(qb + (nt * 8 + g) * 32);
                bf[0] = qrow[t];
                bf[1] = qrow[4 + t];
                int c[NP][4];
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    c[p][0] = c[p][1] = c[p][2] = c[p][3] = 0;
                    if (p == NP - 1) mma_bytes<true>(c[p], a[p], bf);
                    else mma_bytes<false>(c[p], a[p], bf);
                }
                /* epilogue: prediction sum -> residual -> |residual| (encoder.py:544-548, :398) */
                if (fits && interior) { /* warp-uniform: nothing to mask, 32-bit residuals */
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        uint32_t t32 = 0;
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            int hi = c[NP - 1][2 * r + j]; /* planes 1.. as one 32-bit sum (< 2^28) */
#pragma unroll
                            for (int p = NP - 2; p >= 1; p--) hi = (hi << 8) + c[p][2 * r + j];
                            const i64 pred = NP > 1 ? (i64)hi * 256 + (i64)c[0][2 * r + j] : (i64)hi;
                            t32 = __sad(xr[r], (int)(uint32_t)(u64)(pred >> shift[nt][j]), t32);
                        }
                        acc[nt][j] += (u64)t32;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        u64 t64 = 0;
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const int i = s0 + g + 8 * r, m = nt * 8 + 2 * t + j + 1;
                            i64 pred = c[NP - 1][2 * r + j];
#pragma unroll
                            for (int p = NP - 2; p >= 0; p--) pred = pred * 256 + (i64)c[p][2 * r + j];
                            const i64 rr = (i64)xr[r] - (pred >> shift[nt][j]);
                            if (i >= m && i < n) t64 += (u64)(rr < 0 ? -rr : rr);
                        }
                        acc[nt][j] += t64;
                    }
                }
            }
        }
    }
    /* sum over the eight lanes that hold the same columns, then over the warps */
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            u64 v = acc[nt][j];
            v += __shfl_xor_sync(FULL, v, 4);
            v += __shfl_xor_sync(FULL, v, 8);
            v += __shfl_xor_sync(FULL, v, 16);
            if (g == 0) red[b.warp * 32 + nt * 8 + 2 * t + j] = live[nt][j] ? v : 0;
        }
    }
    __syncthreads();
    if (b.tid < L) {
        u64 s = 0;
        for (int w = 0; w < b.nwarps; w++) s += red[w * 32 + b.tid];
        sh.lpc_err[b.tid] = s;
    }
    __syncthreads();
}

/* channel c of the frame as byte planes: plane p holds byte p of every sample (the top plane signed) */
static __device__ __noinline__ void load_byte_planes(const uint8_t *base, uint8_t *bpl, uint32_t bstride, int planes, int n,
                                                     int channels, int cbytes, int c, int tid, int B)
{
    for (int idx = tid; idx < planes * (MMA_GB + MMA_PADB); idx += B) {
        const int p = idx / (MMA_GB + MMA_PADB), j = idx - p * (MMA_GB + MMA_PADB);
        bpl[(uint32_t)p * bstride + (uint32_t)(j < MMA_GB ? j : n + j)] = 0;
    }
    const bool aligned = (((uintptr_t)base) & (uintptr_t)(cbytes - 1)) == 0 || cbytes == 3 || cbytes == 1;
    /* four samples per thread and round: the loads of a round are issued together (one after the other they each
     * waited for HBM), and a round's bytes of a plane leave as one 32-bit store */
    const int n4 = n & ~3;
    for (int i = 4 * tid; i < n4; i += 4 * B) {
        int32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = load_sample(base, (u64)(i + k) * (u64)channels + (u64)c, cbytes, 0, aligned);
        for (int p = 0; p < planes; p++) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) w |= (((uint32_t)v[k] >> (8 * p)) & 0xFFu) << (8 * k);
            *(uint32_t *)(bpl + (uint32_t)p * bstride + (uint32_t)(MMA_GB + i)) = w;
        }
    }
    for (int i = n4 + tid; i < n; i += B) {
        const int32_t v = load_sample(base, (u64)i * (u64)channels + (u64)c, cbytes, 0, aligned);
        for (int p = 0; p < planes; p++) bpl[(uint32_t)p * bstride + (uint32_t)(MMA_GB + i)] = (uint8_t)((uint32_t)v >> (8 * p));
    }
}

template <int LMAX, int TB, int MINB>
__global__ void __launch_bounds__(TB, MINB) choose_mma_kernel(const EncParams p, const MmaLayout ml)
{
    FLACB_DYN_SMEM(smem);
    __shared__ Shared sh;
    FLACB_SIM_POISON(sh);
    Blk b(p, sh);
    b.tid = (int)threadIdx.x;
    b.B = (int)blockDim.x;
    b.lane = b.tid & 31;
    b.warp = b.tid >> 5;
    b.nwarps = b.B >> 5;
    const uint32_t f = blockIdx.x;
    if (b.tid == 0) sh.frame = f;
    if (b.tid < MAX_CH) sh.chan_status[b.tid] = 0;
    b.n = (f == p.n_frames - 1) ? p.tail_size : p.block_size;
    b.ntiles = 0;
    const SmemLayout &lay = p.lay; /* offsets inside a coefficient record */
    b.ldL = lpc_row_stride(p.L);
    uint8_t *bpl = smem + ml.bpl;
    u64 *red = (u64 *)(smem + ml.red);
    const uint8_t *frame_pcm = p.pcm + (u64)f * (u64)p.block_size * (u64)p.channels * (u64)p.cbytes;
    pdl_launch_dependents();
    /* the frame's PCM does not depend on the stages in front: channel 0 is split while they finish */
    load_byte_planes(frame_pcm, bpl, ml.bstride, ml.planes, b.n, p.channels, p.cbytes, 0, b.tid, b.B);
    pdl_wait();
    for (int c = 0; c < p.channels; c++) {
        const uint32_t *rec = (const uint32_t *)(p.pre + ((u64)f * (u64)p.channels + (u64)c) * (u64)p.pre_stride);
        uint32_t *loc = (uint32_t *)(smem + ml.lpc + (uint32_t)c * ml.lpc_stride);
        /* autocorrelation (for the introspection record), then everything behind the Levinson matrix */
        for (uint32_t i = (uint32_t)b.tid; i < (lay.o_levprod - lay.o_autoc) / 4; i += (uint32_t)b.B) loc[i] = rec[lay.o_autoc / 4 + i];
        for (uint32_t i = (uint32_t)b.tid; i < (lay.o_fixed + 64 - lay.o_qc) / 4; i += (uint32_t)b.B) loc[ml.c_qc / 4 + i] = rec[lay.o_qc / 4 + i];
    }
    SmemLayout clay = lay; /* the same record, compacted */
    clay.o_autoc = ml.c_autoc, clay.o_levprod = ml.c_autoc, clay.o_ldc = ml.c_autoc, clay.o_qc = ml.c_qc, clay.o_qmeta = ml.c_qmeta, clay.o_fixed = ml.c_fixed;
    __syncthreads();
    for (int c = 0; c < p.channels; c++) {
        unsigned char *q = smem + ml.lpc + (uint32_t)c * ml.lpc_stride;
        bind_lpc_record(b, q, clay, p.L);
        if (b.tid == 0) sh.cur_ch = c;
        /* coefficient bytes, taps reversed, zero beyond the order (and for orders without coefficients):
         * qb[m1][31 - j] = q_{m1 + 1}[j] */
        uint8_t *qb = smem + ml.qb;
        for (int idx = b.tid; idx < 32 * 32; idx += b.B) {
            const int m1 = idx >> 5, j = 31 - (idx & 31);
            int v = 0;
            if (m1 < p.L && j < b.qncoef[m1]) v = b.qc[m1 * b.ldL + j];
            qb[idx] = (uint8_t)(int8_t)v;
        }
        if (c > 0) load_byte_planes(frame_pcm, bpl, ml.bstride, ml.planes, b.n, p.channels, p.cbytes, c, b.tid, b.B);
        __syncthreads();
        if (p.L > 0 && *b.lev_stat == 0) {
            if (ml.planes == 3) lpc_errors_mma<3>(b, bpl, ml.bstride, p.L, qb, red);
            else if (ml.planes == 2) lpc_errors_mma<2>(b, bpl, ml.bstride, p.L, qb, red);
            else lpc_errors_mma<1>(b, bpl, ml.bstride, p.L, qb, red);
        }
        analyse_choose<LMAX, -1>(b, nullptr, c, nullptr);
        SubDecision *out = p.dec + ((u64)f * (u64)p.channels + (u64)c);
        const uint32_t *src = (const uint32_t *)&sh.dec[c];
        for (uint32_t i = (uint32_t)b.tid; i < sizeof(Decision) / 4; i += (uint32_t)b.B) ((uint32_t *)&out->d)[i] = src[i];
        if (b.tid == 0) out->status = sh.chan_status[c];
        __syncthreads(); /* the planes and the reduction scratch are free again */
    }
}

#endif /* !FLACB_HOST_ONLY */

} /* namespace flacb */
