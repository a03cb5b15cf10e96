/*
 * flacb200_kernels.cuh -- device code of the B200 FLAC encode path.
 *
 * LPC configurations (max_lpc_order > 0) run as four kernels; every stage is one CTA per frame
 * except the second, the hand-over between them is a small record per subframe in HBM (HBM is
 * not the bound of this path: the reference's semantics cost hundreds of integer / fp64
 * operations per sample against three bytes of traffic, so re-reading the PCM per stage is free
 * and small single-purpose kernels win on occupancy and instruction cache):
 *
 *   A   analyse_kernel           frame/channel split (reference flac/encoder.py:87, :102),
 *                                fixed-predictor errors, orders 0-4 (encoder.py:331-359),
 *                                windowed fp64 autocorrelation (encoder.py:368-371, :443-450)
 *   B   lpc_coefficients_kernel  Levinson-Durbin and coefficient quantisation of every order
 *                                (encoder.py:453-534), one warp per subframe
 *   C1  choose_kernel            integer residual error of every LPC order on every sample
 *                                (encoder.py:387-398, :537-548), best order, fixed-vs-LPC
 *                                choice (encoder.py:135-157)
 *   C2  encode_frames_kernel     winner's residual in place of the samples, Rice partition
 *                                search (encoder.py:632-760) -> exact size of every subframe;
 *                                frame size -> byte offset of the frame in the stream by a
 *                                decoupled look-back over a device-wide array of frame sizes
 *                                (single pass, no compaction copy); frame header + CRC-8
 *                                (encoder.py:194-234), subframes (encoder.py:553-627), Rice
 *                                codes (encoder.py:765-806) placed MSB first
 *                                (binary.py:144-216) in a shared memory window, flushed to HBM;
 *                                CRC-16 (encoder.py:159-165, crc.py)
 *
 * Fixed predictors only (max_lpc_order == 0): encode_frames_kernel<.., false> does everything.
 *
 * Work split inside a CTA: thread t of tile k owns the E consecutive samples [(k*B + t)*E, +E),
 * moved between shared memory and registers with 16-byte accesses (swz).  Consecutive samples
 * per thread give the filters a register-resident history window.
 *
 * The same source compiles for the CPU SIMT simulator of tests/sim (FLACB_CPU_SIM), which is
 * test infrastructure: it lets the kernel logic be checked against the golden vectors on a
 * machine without a GPU.
 */
#pragma once

#include <stdint.h>

#ifdef FLACB_CPU_SIM
#include "cpu_sim.h"
#else
#include <cuda_runtime.h>
#define FLACB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define FLACB_SIM_POISON(obj) ((void)0) /* the CPU simulator poisons static shared memory here */
#endif

#include "../../include/flacb200.h"

namespace flacb {

typedef unsigned long long u64;
typedef long long i64;

/* Programmatic dependent launch (sm_90+): the stages of an encode follow one another on one stream;
 * a stage lets the next one start its CTAs as soon as all of its own are running
 * (pdl_launch_dependents), and the next one does what does not depend on its predecessor -- taking
 * a frame, loading and splitting its PCM -- before it waits for the predecessor's results
 * (pdl_wait: the whole predecessor grid is complete and its memory visible).  The ragged end of
 * one stage overlaps the start of the next.  Without the launch attribute both are no-ops. */
#if defined(FLACB_CPU_SIM) || defined(FLACB_HOST_ONLY)
static inline void pdl_wait() {}
static inline void pdl_launch_dependents() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

constexpr int E = 8;            /* consecutive samples per thread per tile */
constexpr int GUARD = 64;       /* zero history in front of every channel plane (a multiple of 16: keeps the swizzle phase) */
constexpr int TAILPAD = 48;     /* zero samples behind a plane (window reads run past the end) */
#ifndef FLACB_STG_WORDS
#define FLACB_STG_WORDS 4096
#endif
constexpr int STG_WORDS = FLACB_STG_WORDS; /* staging window of the bit packer, 32-bit words */
constexpr int MAXB = 576;       /* threads per CTA, upper bound */
constexpr int MAX_CH = 8;
constexpr int RED_STRIDE = 40;  /* u64 slots per warp in the reduction scratch */
constexpr int MAX_CAND = 16;    /* Rice partition orders 0..15 */
constexpr int CRC_SLICE = 64;   /* bytes per thread in the parallel CRC-16 */
constexpr int XPOW_HI = 8192;   /* x^(8*256*h) table entries: frames up to 2 MiB */
constexpr unsigned FULL = 0xffffffffu;

/* Optional per-phase cycle counters (development builds, -DFLACB_PHASE_TIMING): thread 0 of every
 * CTA adds the cycles between two marks to stats[8 + phase]; tools/perf_sweep.py prints them. */
#ifdef FLACB_PHASE_TIMING
#define FLACB_MARK(b, k)                                                      \
    do {                                                                      \
        if ((b).tid == 0) {                                                   \
            const long long t_ = clock64();                                   \
            atomicAdd(&(b).p.stats[8 + (k)], (u64)(t_ - (b).sh.t_last));      \
            (b).sh.t_last = t_;                                               \
        }                                                                     \
    } while (0)
#else
#define FLACB_MARK(b, k) do { } while (0)
#endif

/* Helpers that take the per-thread block state (Blk) stay inline: an out-of-line call forces that
 * state into local memory, which cost 40 % on the fixed-predictor path when measured.  The two
 * helpers that are out of line (flush_window_impl, resolve_offset) take plain arguments. */

/* look-back state flags (top two bits of a 64-bit word, value below) */
constexpr u64 LB_AGG = 1ull << 62;
constexpr u64 LB_PREFIX = 2ull << 62;
constexpr u64 LB_VALUE = (1ull << 62) - 1;

/* plane index -> shared memory word.  A thread owns runs of 8 consecutive samples and moves them
 * with 16-byte accesses; a quarter warp's eight accesses are 8 words apart and would only touch
 * half of the banks, twice.  Flipping bit 2 in every other group of 32 words (swapping the two
 * halves of every run there) spreads them over all 32 banks without padding.  Aligned groups of
 * four words stay contiguous, so x[swz(i) .. swz(i) + 3] is samples i .. i + 3 for i % 4 == 0. */
__host__ __device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 4u); }

struct SmemLayout {
    uint32_t xs, xwords, res, cum, kall, stg, stg_words, fb, pst, red, lpc, lpc_stride, total;
    /* per-channel LPC front end, offsets inside a channel's lpc area */
    uint32_t o_autoc, o_levprod, o_ldc, o_qc, o_qmeta, o_fixed;
    uint32_t chan_stride; /* bytes between the running-sum / parameter areas of two channels */
    int xs_all;
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

/* row stride of the per-order coefficient tables: rows are padded to a multiple of 8 so that the
 * fixed-width filter bodies of lpc_errors can read a whole (zero padded) row */
__host__ __device__ inline int lpc_row_stride(int L) { return (L + 7) & ~7; }

/* Shared memory plan.  all_planes: every channel plane resident; otherwise one plane that is
 * (re)loaded per channel and phase, which leaves room for more CTAs per SM.  `threads` sizes the
 * staging window: a tile of threads * E samples packs into at most (bps + 4) bits per sample
 * (SURVEY.md 9.5), so a window of that size never needs a second pass. */
__host__ __device__ inline SmemLayout make_layout(int nmax, int channels, int bps, int rmax, int L, int threads,
                                                   int all_planes)
{
    SmemLayout l;
    l.xwords = ((uint32_t)(GUARD + nmax + TAILPAD) + 7u) & ~7u;
    const uint32_t plane = l.xwords * 4;
    int tile = threads * E < nmax ? threads * E : nmax;
    uint32_t words = ((uint32_t)tile * (uint32_t)(bps + 4) + 31) / 32 + 96; /* + serial header part */
    /* better still: the whole frame (one flush, at the end), up to the cap */
    const uint32_t per = ((uint32_t)nmax * (uint32_t)(bps + 4) + 7) / 8 + (1u << rmax) + 256;
    const uint32_t frame_words = (per * (uint32_t)channels + 64 + 3) / 4 + 96;
    if (words < frame_words) words = frame_words;
    if (words < 512) words = 512;
    if (words > (uint32_t)STG_WORDS) words = (uint32_t)STG_WORDS;
    l.stg_words = words;
    /* all_planes: every channel's plane, running sums and Rice parameters resident (nothing is
     * reloaded or recomputed when the frame is serialised); otherwise one of each.  The zig-zag
     * residual of a channel replaces its samples in the plane (residual_and_cum). */
    l.xs_all = all_planes >= 1;
    uint32_t off = 0;
    l.xs = off; off += l.xs_all ? plane * (uint32_t)channels : plane;
    const uint32_t cum_b = align16(((uint32_t)nmax / E + 4) * 8);
    const uint32_t kall_b = align16(2u << rmax);
    l.chan_stride = cum_b + kall_b;
    l.res = l.xs + GUARD * 4; /* sample i and residual i share a word */
    l.cum = off;
    l.kall = off + cum_b;
    off += l.chan_stride * (l.xs_all ? (uint32_t)channels : 1u);
    l.stg = off; off += align16((words + 4) * 4);
    /* reduction scratch; the packer's per-thread arrays live in the same bytes (never live together) */
    const uint32_t thr = (uint32_t)((threads + 31) / 32 * 32);
    uint32_t red_b = align16((thr / 32) * RED_STRIDE * 8);
    const uint32_t pack_b = align16(thr * 4) + align16((thr + 1) * 4);
    if (red_b < pack_b) red_b = pack_b;
    l.red = off; l.fb = off; l.pst = off + align16(thr * 4); off += red_b;
    /* LPC front end of every channel (autocorrelation, Levinson and quantised coefficients of
     * every order, fixed-predictor errors): all channels run their Levinson recursion at once */
    uint32_t o = 0;
    l.o_autoc = o; o += align16((uint32_t)(L + 1) * 8);
    l.o_levprod = o; o += align16((uint32_t)(L > 0 ? L : 1) * 8);
    l.o_ldc = o; o += align16((uint32_t)(L * lpc_row_stride(L)) * 8);
    l.o_qc = o; o += align16((uint32_t)(L * lpc_row_stride(L)) * 4);
    l.o_qmeta = o; o += align16((uint32_t)(L > 0 ? L : 1) * 16 + 16); /* shift, ncoef, stat, abs per order + lev_stat */
    l.o_fixed = o; o += 64;
    l.lpc_stride = o;
    l.lpc = off; off += o * (uint32_t)channels;
    l.total = off;
    return l;
}

/* Shared memory plan of analyse_kernel (stage A of the staged pipeline): sample planes, the
 * windowed samples of the two tapered ends, per-warp partial sums.  Only the fields that
 * fixed_errors / autocorrelation read are filled; the o_* offsets and lpc_stride describe the
 * per-subframe record in HBM and equal make_layout's. */
__host__ __device__ inline SmemLayout make_layout_front(int nmax, int channels, int bps, int rmax, int L, int threads,
                                                         int all_planes)
{
    SmemLayout l = make_layout(nmax, channels, bps, rmax, L, threads, 0);
    l.xs_all = all_planes >= 1;
    const uint32_t plane = l.xwords * 4;
    uint32_t off = 0;
    l.xs = off; off += l.xs_all ? plane * (uint32_t)channels : plane;
    /* both tapered ends of a Tukey(0.5) window: n/2 + 152 samples at most (autocorrelation()
     * computes on the fly if a window ever needs more) */
    const uint32_t cnt = nmax >= 1024 ? (uint32_t)nmax / 2 + 256 : (uint32_t)nmax + 160;
    const uint32_t wx_b = align16((cnt + (cnt >> 3) + 2) * 8);
    l.res = off; l.cum = off; l.kall = off + wx_b; /* wx_bytes = kall - res */
    l.chan_stride = wx_b;
    off += wx_b;
    const uint32_t autoc_words = (uint32_t)((threads + 31) / 32) * (FLACB_MAX_LPC + 1) * 6 + 8;
    l.stg_words = autoc_words;
    l.stg = off; off += align16((autoc_words + 4) * 4);
    const uint32_t thr = (uint32_t)((threads + 31) / 32 * 32);
    l.red = off; l.fb = off; l.pst = off; off += align16((thr / 32) * RED_STRIDE * 8);
    l.lpc = off; /* no LPC area in shared memory: the records live in HBM */
    l.total = off;
    return l;
}

/* Shared memory plan of choose_kernel (stage C1): sample planes, reduction scratch and the
 * coefficient records of the frame's subframes. */
__host__ __device__ inline SmemLayout make_layout_choose(int nmax, int channels, int bps, int rmax, int L, int threads,
                                                          int all_planes)
{
    SmemLayout l = make_layout(nmax, channels, bps, rmax, L, threads, 0);
    l.xs_all = all_planes >= 1;
    const uint32_t plane = l.xwords * 4;
    uint32_t off = 0;
    l.xs = off; off += l.xs_all ? plane * (uint32_t)channels : plane;
    l.res = l.cum = l.kall = l.stg = off;
    l.stg_words = 0;
    const uint32_t thr = (uint32_t)((threads + 31) / 32 * 32);
    l.red = off; l.fb = off; l.pst = off; off += align16((thr / 32) * RED_STRIDE * 8);
    l.lpc = off; off += l.lpc_stride * (uint32_t)channels;
    l.total = off;
    return l;
}

/* ---- shared memory plan of serialise_kernel (flacb200_serialise.cuh) ---- */
constexpr int V2_GUARD = 32;   /* zero history in front of a plane (>= FLACB_MAX_LPC, a multiple of 32) */
constexpr int V2_MAXC = 9;     /* candidate partition orders 0..8 */
constexpr int V2_MAX_RICE = 8;

/* plane index -> shared memory word: 16-byte units, eight to a 128-byte row, the unit index
 * xor-ed with the row number.  Aligned groups of four words stay contiguous. */
__host__ __device__ __forceinline__ uint32_t swz2(uint32_t i)
{
    const uint32_t u = i >> 2;
    return ((u ^ ((u >> 3) & 7u)) << 2) | (i & 3u);
}

struct SerLayout {
    uint32_t xwords;    /* words of one plane: guard + threads * S */
    uint32_t pstride;   /* words between two planes: = 4 (mod 32), so that the same sample of two channels lies 4 banks apart */
    uint32_t planes, kall, kall_stride, cumw, partf, stg, stg_words, total;
    int S;
};

/* worst-case bits of one subframe, SURVEY.md 9.5 (the per-channel part of frame_bound) */
__host__ __device__ inline uint32_t subframe_bound_bytes(int bs, int bps, int rmax)
{
    return ((uint32_t)bs * (uint32_t)(bps + 4) + 7) / 8 + (1u << rmax) + 256;
}

/* budget: dynamic shared memory this frame may use.  stg_words == 0: does not fit. */
__host__ __device__ inline SerLayout make_ser_layout(int nmax, int channels, int bps, int rmax, int threads, int S, uint32_t budget)
{
    SerLayout l;
    l.S = S;
    l.xwords = (uint32_t)(V2_GUARD + threads * S);
    l.pstride = l.xwords + 4u;
    uint32_t off = 0;
    l.planes = off; off += align16(l.pstride * 4u * (uint32_t)channels);
    l.kall_stride = align16(2u << rmax);
    l.kall = off; off += l.kall_stride * (uint32_t)channels;
    l.cumw = off; off += align16((uint32_t)threads * 8u);
    l.partf = off; off += align16((uint32_t)channels * (uint32_t)((threads + 31) / 32) * 5u * 8u); /* fixed-predictor errors per warp */
    l.stg = off;
    const uint32_t per = subframe_bound_bytes(nmax, bps, rmax);
    const uint32_t want = (per * (uint32_t)channels + 64 + 3) / 4 + 16; /* the whole frame */
    const uint32_t least = (per + 3) / 4 + 32;                        /* one subframe + the partial word in front */
    uint32_t words = 0;
    if (budget >= off + 64) {
        const uint32_t room = (budget - off - 32) / 4;
        words = want <= room ? want : room;
        if (words < least) words = 0;
    }
    l.stg_words = words;
    l.total = off + align16((words + 4) * 4);
    return l;
}

struct SubDecision;

struct EncParams {
    const uint8_t *pcm;        /* interleaved, cbytes per sample, little endian */
    u64 n_sample_frames;
    int channels, bps, cbytes, block_size;
    int L, P, rmin, rmax;
    u64 first_frame_number;
    uint32_t flags;
    uint32_t n_frames;
    int tail_size;             /* size of the last frame (== block_size when it is full) */
    const double *win_full;    /* tukey(block_size), zero padded by 64 */
    const double *win_tail;    /* tukey(tail_size), zero padded by 64 */
    int win_full_bad, win_tail_bad; /* ZeroDivisionError of encoder.py:437 (n in 4..7) */
    int flat_full_lo, flat_full_hi;  /* win_full[i] == 1.0 exactly for i in [lo, hi) */
    int flat_tail_lo, flat_tail_hi;
    uint8_t *out;
    u64 out_cap;
    u64 slot_stride;           /* != 0: slot mode (frame f at f * slot_stride) */
    uint32_t *frame_bytes;     /* [n_frames] */
    u64 *frame_off;            /* [n_frames + 1] exclusive prefix (look-back mode) */
    int32_t *frame_status;     /* [n_frames] */
    u64 *lb_state;             /* [n_frames] look-back words, zeroed before launch */
    uint32_t *ticket;          /* zeroed before launch */
    u64 *stats;                /* device counters: [0..4] fixed order histogram, [5] lpc, [6] ood */
    flacb_subframe_info *info; /* optional, device memory */
    int all_planes;            /* every channel plane resident in shared memory */
    SmemLayout lay;            /* make_layout(...) of this launch, computed once on the host */
    /* LPC pipeline: per subframe one "LPC front end" record in HBM (autocorrelation, Levinson and
     * quantised coefficients of every order, fixed-predictor errors), laid out by the o_* offsets
     * of make_layout, pre_stride bytes apart */
    uint8_t *pre;
    uint32_t pre_stride;
    SubDecision *dec;          /* [n_frames * channels] choices of choose_kernel for encode_frames_kernel<.., true> */
    /* serialise_kernel (flacb200_serialise.cuh) */
    SerLayout slay;
    const uint16_t *crc4;      /* [4][256] slice-by-4 CRC-16 tables */
    const uint16_t *crc_w;     /* x^(8 i) mod 0x18005 for i up to the frame bound */
};

#ifndef FLACB_HOST_ONLY

/* CRC-16 helper tables: x^(8*l) mod P for l < 256 and x^(8*256*h) mod P (P = 0x18005) */
/* (one copy per translation unit: flacb200_launch.h) */
static __constant__ uint16_t c_xpow_lo[256];
static __constant__ uint16_t c_xpow_hi[XPOW_HI];
static __constant__ uint16_t c_crc16_tab[256];

struct Decision {
    int32_t type;   /* 0 fixed, 1 lpc */
    int32_t order;  /* order field of the header */
    int32_t warm;   /* warm-up samples == predictor order used by the Rice search */
    int32_t shift, ncoef;
    int32_t porder, method;
    uint32_t bits;  /* serialised subframe size */
    int32_t q[FLACB_MAX_LPC];
};

/* what choose_kernel hands to the serialising kernel for one subframe */
struct SubDecision {
    Decision d;
    int32_t status;
};

struct Shared {
    u64 lpc_err[FLACB_MAX_LPC];
    u64 qsum[MAX_CAND];
    Decision dec[MAX_CH];
    int chan_status[MAX_CH]; /* first error of each channel, in the reference's evaluation order */
    int32_t warm[MAX_CH][FLACB_MAX_LPC]; /* warm-up samples (the planes become residuals) */
    int cur_ch;
    uint32_t wmax_hi; /* autocorrelation: high word of the largest |windowed sample| of the tapered ends (0 between uses) */
    uint32_t cand_o[MAX_CAND], cand_magic[MAX_CAND];
    uint32_t sizeA[MAX_CAND], sizeB[MAX_CAND], any5[MAX_CAND];
    uint32_t errkey;
    int ncand;
    /* packer */
    uint32_t pos, wbase, crc;
    u64 frame_off;
    uint32_t frame_bytes, payload_bytes;
    int status, writable, range_flag, have_off;
    long long t_last; /* FLACB_PHASE_TIMING */
    /* where the frame goes (copied from the launch parameters for the out-of-line helpers) */
    uint8_t *o_out;
    u64 o_cap, o_slot_stride;
    u64 *o_lb_state, *o_frame_off;
    uint32_t o_n_frames;
    uint32_t frame;
    uint32_t scan_warp[33];
    u64 scan_warp64[33];
    uint16_t crc16_tab[256];
};

struct Blk {
    const EncParams &p;
    Shared &sh;
    int tid, B, lane, warp, nwarps;
    int n;      /* block size of this frame */
    int ntiles;
    int32_t *xs;
    uint32_t xwords;
    uint32_t *res;
    u64 *cum;
    uint8_t *kall;
    uint32_t *stg;
    uint32_t *fb;
    uint32_t *pst;
    u64 *red;
    /* LPC front end of the current channel (shared memory, see make_layout) */
    double *autoc, *lev_prod;
    double *ldc;   /* [L][L] Levinson coefficients after iteration k */
    int32_t *qc;   /* [L][L] quantised coefficients per order */
    int32_t *qshift, *qncoef, *qstat, *lev_stat;
    uint32_t *qabs;
    u64 *fixed_err;
    int ldL;       /* row stride of ldc / qc */
    uint32_t stg_words;
    uint32_t wx_bytes; /* bytes of the residual + running-sum area (scratch of the autocorrelation) */
    const double *win;
    int flat_lo, flat_hi;
    __device__ Blk(const EncParams &p_, Shared &sh_) : p(p_), sh(sh_) {}
};

__device__ __forceinline__ void set_status(Blk &b, int code)
{
    /* first error in the reference's evaluation order wins: per channel in the order of the
     * stages (callers run in that order), then over the channels in ascending order */
    int &st = b.sh.chan_status[b.sh.cur_ch];
    if (st == 0) st = code;
}

/* ------------------------------------------------------------------------------------- */
/* 16-byte shared-memory moves of a thread's runs: N consecutive words starting at logical index
 * i (i % 4 == 0, N % 4 == 0) of a swizzled array                                           */

template <int N>
__device__ __forceinline__ void ld_run(const int32_t *x, uint32_t i, int (&v)[N])
{
#pragma unroll
    for (int k = 0; k < N; k += 4) {
        const int4 t = *(const int4 *)(x + swz(i + (uint32_t)k));
        v[k] = t.x, v[k + 1] = t.y, v[k + 2] = t.z, v[k + 3] = t.w;
    }
}

template <int N>
__device__ __forceinline__ void st_run(int32_t *x, uint32_t i, const int (&v)[N])
{
#pragma unroll
    for (int k = 0; k < N; k += 4) {
        int4 t;
        t.x = v[k], t.y = v[k + 1], t.z = v[k + 2], t.w = v[k + 3];
        *(int4 *)(x + swz(i + (uint32_t)k)) = t;
    }
}

/* ------------------------------------------------------------------------------------- */
/* block-wide primitives                                                                  */

/* warp sums through the integer reduction unit (REDUX): three 21-bit pieces, each of which
 * sums to less than 2^26 over 32 lanes.  Exact for |v| < 2^62. */
__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#ifdef FLACB_WARPSUM_SHFL
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
#endif
    const uint32_t p0 = (uint32_t)v & 0x1FFFFFu, p1 = (uint32_t)(v >> 21) & 0x1FFFFFu, p2 = (uint32_t)(v >> 42);
    return (u64)__reduce_add_sync(FULL, p0) + ((u64)__reduce_add_sync(FULL, p1) << 21) +
           ((u64)__reduce_add_sync(FULL, p2) << 42);
}

__device__ __forceinline__ i64 warp_sum_i64(i64 v)
{
#ifdef FLACB_WARPSUM_SHFL
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
#endif
    const uint32_t p0 = (uint32_t)v & 0x1FFFFFu, p1 = (uint32_t)(v >> 21) & 0x1FFFFFu;
    const int p2 = (int)(v >> 42); /* arithmetic: v = p2 * 2^42 + p1 * 2^21 + p0 */
    return (i64)__reduce_add_sync(FULL, p0) + ((i64)__reduce_add_sync(FULL, p1) << 21) +
           (i64)__reduce_add_sync(FULL, p2) * ((i64)1 << 42);
}

template <int K>
__device__ void block_reduce_u64(Blk &b, u64 (&v)[K], u64 *out)
{
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = warp_sum_u64(v[k]);
    if (b.lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) b.red[b.warp * K + k] = v[k];
    }
    __syncthreads();
    if (b.tid < K) {
        u64 s = 0;
        for (int w = 0; w < b.nwarps; w++) s += b.red[w * K + b.tid];
        out[b.tid] = s;
    }
    __syncthreads();
}

static __device__ uint32_t block_excl_scan_u32(Blk &b, uint32_t v, uint32_t &total)
{
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, off);
        if (b.lane >= off) inc += t;
    }
    if (b.lane == 31) b.sh.scan_warp[b.warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < b.nwarps; w++) {
        uint32_t x = b.sh.scan_warp[w];
        if (w < b.warp) base += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

static __device__ u64 block_excl_scan_u64(Blk &b, u64 v, u64 &total)
{
    u64 inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        u64 t = __shfl_up_sync(FULL, inc, off);
        if (b.lane >= off) inc += t;
    }
    if (b.lane == 31) b.sh.scan_warp64[b.warp] = inc;
    __syncthreads();
    u64 base = 0, tot = 0;
    for (int w = 0; w < b.nwarps; w++) {
        u64 x = b.sh.scan_warp64[w];
        if (w < b.warp) base += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

/* ------------------------------------------------------------------------------------- */
/* phase 0: frame/channel split (encoder.py:87 batch, :102 de-interleave)                 */

__device__ __forceinline__ int32_t load_sample(const uint8_t *base, u64 s, int cbytes, int bps,
                                               bool aligned)
{
    int32_t v;
    if (cbytes == 2) {
        if (aligned) v = ((const int16_t *)base)[s];
        else v = (int16_t)((uint32_t)base[2 * s] | ((uint32_t)base[2 * s + 1] << 8));
    } else if (cbytes == 3) {
        const uint8_t *q = base + 3 * s;
        uint32_t u = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
        v = (int32_t)(u << 8) >> 8;
    } else if (cbytes == 4) {
        if (aligned) v = ((const int32_t *)base)[s];
        else {
            const uint8_t *q = base + 4 * s;
            v = (int32_t)((uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) |
                          ((uint32_t)q[3] << 24));
        }
    } else {
        v = (int8_t)base[s];
    }
    (void)bps;
    return v;
}

/* c_only < 0: all channels into their planes (xs_all); otherwise channel c_only into plane 0.
 * CH: compile-time channel count for the common layouts (0 = any, divides at run time).
 * Plain arguments, no block state: the dispatcher below is out of line (one copy per kernel
 * instead of one per call site and channel count: 85 KB of code in the first staged build). */
template <int CH>
__device__ __forceinline__ void load_planes_t(const uint8_t *base, int32_t *xs, uint32_t xwords, int n, int channels,
                                              int cbytes, int c_only, int tid, int B)
{
    const int ch = CH ? CH : channels;
    const bool aligned = (((uintptr_t)base) & (uintptr_t)(cbytes - 1)) == 0 || cbytes == 3 || cbytes == 1;
    const int planes = c_only < 0 ? ch : 1;
    /* zero guard and tail pad */
    for (int idx = tid; idx < planes * (GUARD + TAILPAD); idx += B) {
        int pl = idx / (GUARD + TAILPAD);
        int j = idx - pl * (GUARD + TAILPAD);
        int i = j < GUARD ? j : n + j; /* logical index GUARD + n + (j - GUARD) */
        xs[(uint32_t)pl * xwords + swz((uint32_t)i)] = 0;
    }
    const int total = n * ch;
    int done = 0; /* samples [0, done) are handled by a vector path, the rest one by one */
    auto put = [&](int sidx, int32_t v) {
        /* sample `sidx` of the interleaved frame -> its plane */
        const int i = sidx / ch, c = sidx - i * ch; /* shifts for the compile-time counts */
        if (c_only < 0) xs[(uint32_t)c * xwords + swz((uint32_t)(GUARD + i))] = v;
        else if (c == c_only) xs[swz((uint32_t)(GUARD + i))] = v;
    };
    if (cbytes == 2 && (((uintptr_t)base) & 15) == 0) {
        /* 16-bit: eight samples per 16-byte load; mono and stereo go to their planes with
         * 16-byte stores (the four / eight samples of a channel are consecutive) */
        const int groups = total >> 3;
        const uint4 *src = (const uint4 *)base;
        for (int g = tid; g < groups; g += B) {
            const uint4 w = src[g];
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            int v[8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                v[2 * k] = (int32_t)(int16_t)(ww[k] & 0xFFFFu);
                v[2 * k + 1] = (int32_t)ww[k] >> 16;
            }
            if (CH == 1) {
                st_run<8>(xs, (uint32_t)(GUARD + 8 * g), v);
            } else if (CH == 2) {
                const int ev[4] = {v[0], v[2], v[4], v[6]}, od[4] = {v[1], v[3], v[5], v[7]};
                if (c_only < 0) {
                    st_run<4>(xs, (uint32_t)(GUARD + 4 * g), ev);
                    st_run<4>(xs + xwords, (uint32_t)(GUARD + 4 * g), od);
                } else {
                    st_run<4>(xs, (uint32_t)(GUARD + 4 * g), c_only == 0 ? ev : od);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) put(8 * g + k, v[k]);
            }
        }
        done = groups << 3;
    } else if (cbytes == 3 && (((uintptr_t)base) & 3) == 0) {
        /* 24-bit: four samples per three 32-bit loads */
        const int groups = total >> 2;
        const uint32_t *src = (const uint32_t *)base;
        for (int g = tid; g < groups; g += B) {
            const uint32_t w0 = src[3 * g], w1 = src[3 * g + 1], w2 = src[3 * g + 2];
            put(4 * g, (int32_t)(w0 << 8) >> 8);
            put(4 * g + 1, (int32_t)(((w0 >> 24) | (w1 << 8)) << 8) >> 8);
            put(4 * g + 2, (int32_t)(((w1 >> 16) | (w2 << 16)) << 8) >> 8);
            put(4 * g + 3, (int32_t)w2 >> 8);
        }
        done = groups << 2;
    }
    for (int sidx = done + tid; sidx < total; sidx += B) put(sidx, load_sample(base, (u64)sidx, cbytes, 0, aligned));
}

static __device__ __noinline__ void load_planes_impl(const uint8_t *base, int32_t *xs, uint32_t xwords, int n, int channels,
                                                      int cbytes, int c_only, int tid, int B)
{
    switch (channels) {
    case 1: load_planes_t<1>(base, xs, xwords, n, channels, cbytes, c_only, tid, B); break;
    case 2: load_planes_t<2>(base, xs, xwords, n, channels, cbytes, c_only, tid, B); break;
    case 8: load_planes_t<8>(base, xs, xwords, n, channels, cbytes, c_only, tid, B); break;
    default: load_planes_t<0>(base, xs, xwords, n, channels, cbytes, c_only, tid, B); break;
    }
    __syncthreads();
}

__device__ __forceinline__ void load_planes(Blk &b, int c_only)
{
    const EncParams &p = b.p;
    const uint8_t *base = p.pcm + (u64)b.sh.frame * (u64)p.block_size * (u64)p.channels * (u64)p.cbytes;
    load_planes_impl(base, b.xs, b.xwords, b.n, p.channels, p.cbytes, c_only, b.tid, b.B);
}

/* ------------------------------------------------------------------------------------- */
/* encoder.py:331-359 encode_subframe_fixed: sum |residual| of orders 0..4                */

__device__ __forceinline__ void fixed_residuals(const int (&a)[E + 4], int e, int (&r)[5])
{
    const int x0 = a[e + 4], x1 = a[e + 3], x2 = a[e + 2], x3 = a[e + 1], x4 = a[e];
    r[0] = x0;                                  /* common.py:15-21 coefficient table */
    r[1] = x0 - x1;
    r[2] = x0 - 2 * x1 + x2;
    r[3] = x0 - 3 * x1 + 3 * x2 - x3;
    r[4] = x0 - 4 * x1 + 6 * x2 - 4 * x3 + x4;
}

static __device__ void fixed_errors(Blk &b, const int32_t *x)
{
    u64 acc[5] = {0, 0, 0, 0, 0};
    for (int tile = 0; tile < b.ntiles; tile++) {
        const int i0 = (tile * b.B + b.tid) * E;
        if (i0 < b.n) {
            int a[E + 4];
            ld_run<E + 4>(x, (uint32_t)(GUARD + i0 - 4), a);
            if (i0 >= 4 && i0 + E <= b.n) { /* nothing to mask; |r| < 2^28 (bits <= 24): a run's eight fit 32 bits */
                uint32_t t32[5] = {0, 0, 0, 0, 0};
                /* the predictors are successive differences (common.py:15-21 are the binomial rows): four
                 * subtractions per sample serve all five orders; the last difference goes into |a - b| + t */
                int r1[E + 3], r2[E + 2], r3[E + 1];
#pragma unroll
                for (int j = 0; j < E + 3; j++) r1[j] = a[j + 1] - a[j];
#pragma unroll
                for (int j = 0; j < E + 2; j++) r2[j] = r1[j + 1] - r1[j];
#pragma unroll
                for (int j = 0; j < E + 1; j++) r3[j] = r2[j + 1] - r2[j];
#pragma unroll
                for (int e = 0; e < E; e++) {
                    t32[0] = __sad(a[e + 4], 0, t32[0]); /* |r| + t: one instruction */
                    t32[1] = __sad(r1[e + 3], 0, t32[1]);
                    t32[2] = __sad(r2[e + 2], 0, t32[2]);
                    t32[3] = __sad(r3[e + 1], 0, t32[3]);
                    t32[4] = __sad(r3[e + 1], r3[e], t32[4]);
                }
#pragma unroll
                for (int o = 0; o < 5; o++) acc[o] += (u64)t32[o];
            } else {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const int i = i0 + e;
                    if (i < b.n) {
                        int r[5];
                        fixed_residuals(a, e, r);
#pragma unroll
                        for (int o = 0; o < 5; o++)
                            if (i >= o) acc[o] += (u64)(uint32_t)(r[o] < 0 ? -r[o] : r[o]);
                    }
                }
            }
        }
    }
    block_reduce_u64<5>(b, acc, b.fixed_err);
}

/* ------------------------------------------------------------------------------------- */
/* encoder.py:368-371, :443-450: windowed autocorrelation.  The reference sums with Python's
 * builtin sum(), which is compensated (Neumaier) from 3.12 on and returns the correctly
 * rounded exact sum on every lag we measured; a double-double accumulation in any order does
 * the same.  Products and window multiplications are single IEEE operations (no FMA).       */

__device__ __forceinline__ void dd_add(double &s, double &c, double x)
{
    const double t = __dadd_rn(s, x);
    const double bb = __dsub_rn(t, s);
    const double err = __dadd_rn(__dsub_rn(s, __dsub_rn(t, bb)), __dsub_rn(x, bb));
    s = t;
    c = __dadd_rn(c, err);
}

/* A thread's 8 samples are "flat" when every window value they touch (up to lag L ahead) is
 * exactly 1.0 and no product is cut off at the end of the block: then w = x, the products are
 * integers below 2^53 and their sum is exact in int64.  Half of a Tukey(0.5) window is flat. */
__device__ __forceinline__ bool autoc_flat(const Blk &b, int i0, int L)
{
    return i0 >= b.flat_lo && i0 + E - 1 + L < b.flat_hi && i0 + E - 1 < b.n - L - 1;
}

template <int LG, int LGI>
__device__ void autocorrelation(Blk &b, const int32_t *x, int L)
{
    const int n = b.n;
    /* per-warp partial results live in the (idle) staging window of the packer:
     * [warp][lag] exact integer part, then [warp][lag] double-double part.  One barrier for all. */
    i64 *part_i = (i64 *)b.stg;
    double2 *part_d = (double2 *)(part_i + ((b.nwarps * (FLACB_MAX_LPC + 1) + 1) & ~1)); /* 16-byte aligned */
    /* ---- integer part ---- */
    for (int lag0 = 0; lag0 <= L; lag0 += LGI) {
        i64 acc[LGI];
#pragma unroll
        for (int l = 0; l < LGI; l++) acc[l] = 0;
        bool any = false;
        for (int tile = 0; tile < b.ntiles; tile++) {
            const int i0 = (tile * b.B + b.tid) * E;
            if (i0 < n && autoc_flat(b, i0, L)) {
                any = true;
                /* lag0 is 0 (LGI = 13, L <= 12) or a multiple of 8 (LGI = 8): aligned runs */
                int xa[E], xb[(E + LGI - 1 + 3) & ~3];
                ld_run<E>(x, (uint32_t)(GUARD + i0), xa);
                ld_run<(E + LGI - 1 + 3) & ~3>(x, (uint32_t)(GUARD + i0 + lag0), xb);
#pragma unroll
                for (int l = 0; l < LGI; l++) {
                    if (lag0 + l <= L) {
#pragma unroll
                        for (int e = 0; e < E; e++) acc[l] += (i64)xa[e] * (i64)xb[e + l];
                    }
                }
            }
        }
        const bool warp_any = __ballot_sync(FULL, any) != 0; /* warp-uniform */
#pragma unroll
        for (int l = 0; l < LGI; l++) {
            if (lag0 + l <= L) {
                const i64 t = warp_any ? warp_sum_i64(acc[l]) : 0;
                if (b.lane == 0) part_i[b.warp * (FLACB_MAX_LPC + 1) + lag0 + l] = t; /* one writer per slot */
            }
        }
    }
    /* ---- windowed part, double-double ---- */
    /* Windowed samples w = fl(x * win) of the two tapered ends, computed once and kept as doubles
     * in the (still unused) residual + running-sum area: index i -> i + i/8 keeps the stride-8
     * reads of a half-warp on distinct 8-byte banks.  Front range [0, A), back range [Bq, n+pad). */
    double *wxs = (double *)b.res;
    int A = b.flat_lo + 2 * E + L;
    int Bq = b.flat_hi - L - 2 * E;
    if (Bq > n - L - 2 * E - 2) Bq = n - L - 2 * E - 2;
    Bq = Bq < 0 ? 0 : (Bq / E) * E;
    const int top = n + E + L + 8; /* exclusive; window and samples are zero padded behind n */
    if (A >= Bq) A = top, Bq = top; /* no flat middle: one range */
    const int count = A + (top - Bq);
    const bool stored = (size_t)(count + (count >> 3) + 2) * 8 <= b.wx_bytes;
    /* warp sum of LG double-double partials, transposing as it goes: after the first two exchanges
     * (lanes 16 and 8 apart) a lane holds one lag's sum over four lanes, three ordinary butterfly
     * steps finish it: 6 exchanges for 4 lags instead of 20 */
    static_assert(LG == 4, "the transposed warp sum is written for four lags per pass");
    auto warp_sum4 = [&](const double (&s)[LG], const double (&c)[LG], int lag0) {
        const bool b4 = (b.lane & 16) != 0, b3 = (b.lane & 8) != 0;
        auto xchg = [&](double &ks, double &kc, double ss, double sc, int off) {
            const double s2 = __shfl_xor_sync(FULL, ss, off);
            const double c2 = __shfl_xor_sync(FULL, sc, off);
            dd_add(ks, kc, s2);
            kc = __dadd_rn(kc, c2);
        };
        /* lanes 16 apart: the lower lane keeps lags 0, 1 and takes the partner's, the upper 2, 3 */
        double k0s = b4 ? s[2] : s[0], k0c = b4 ? c[2] : c[0], k1s = b4 ? s[3] : s[1], k1c = b4 ? c[3] : c[1];
        xchg(k0s, k0c, b4 ? s[0] : s[2], b4 ? c[0] : c[2], 16);
        xchg(k1s, k1c, b4 ? s[1] : s[3], b4 ? c[1] : c[3], 16);
        /* lanes 8 apart: the lower keeps the first of its two lags */
        double ks = b3 ? k1s : k0s, kc = b3 ? k1c : k0c;
        xchg(ks, kc, b3 ? k0s : k1s, b3 ? k0c : k1c, 8);
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) xchg(ks, kc, ks, kc, off);
        const int lag = lag0 + (b4 ? 2 : 0) + (b3 ? 1 : 0);
        if ((b.lane & 7) == 0 && lag <= L) part_d[b.warp * (FLACB_MAX_LPC + 1) + lag] = make_double2(ks, kc);
    };
    if (stored) {
        /* The last sample never enters a product (encoder.py:449, range(len - lag - 1): neither as the
         * left nor as the right factor), and there are zeros behind it: with w[n - 1] stored as zero
         * no product needs a range check.  (Adding a zero product changes neither part of a sum.) */
        uint32_t mhi = 0;
        for (int t = b.tid; t < count; t += b.B) {
            const int i = t < A ? t : Bq + (t - A);
            const double w = i >= n - 1 ? 0.0 : __dmul_rn((double)x[swz((uint32_t)(GUARD + i))], b.win[i]);
            wxs[t + (t >> 3)] = w;
            const uint32_t hi = (uint32_t)__double2hiint(w) & 0x7FFFFFFFu; /* |w|'s high word: monotone in |w| */
            mhi = hi > mhi ? hi : mhi;
        }
        mhi = __reduce_max_sync(FULL, mhi);
        if (b.lane == 0 && mhi) atomicMax(&b.sh.wmax_hi, mhi);
        __syncthreads();
        /* The tapered runs (everything autoc_flat() refuses) are the runs [0, RA) and [RB, nruns),
         * T of them: thread t takes the runs number t, t + B, ... of that list, Q each -- and the
         * R = T mod B runs that are left over (two of 290 at block 4608: one more full round would
         * leave the block waiting for the one warp that holds them) are done by a warp each with
         * a lane per LAG, eight products per lane instead of eight per lag. */
        const int nruns = (n + E - 1) / E;
        int RA = (b.flat_lo + E - 1) / E;
        int lim = b.flat_hi - L - E;
        if (lim > n - L - E - 1) lim = n - L - E - 1;
        int RB = lim < 0 ? 0 : lim / E + 1;
        if (RA < 0) RA = 0;
        if (RA >= RB || RA >= nruns) RA = nruns, RB = nruns;
        if (RB > nruns) RB = nruns;
        const int T = RA + (nruns - RB);
        int Q = T / b.B, R = T - Q * b.B;
        if (R > 0 && (R + b.nwarps - 1) / b.nwarps >= L + 1) Q += 1, R = 0; /* too many for that: a round like the others */
        const int Tmain = R ? Q * b.B : T;
        auto run_pos = [&](int tau) { /* position of the run's first sample in wxs */
            const int i0 = (tau < RA ? tau : RB + (tau - RA)) * E;
            return i0 < A ? i0 : A + (i0 - Bq);
        };
        /* A thread's sum of one lag is kept against a bias 1.5 * 2^k with 2^(k-1) above anything the sum can
         * reach (k from the largest |w| of the frame's tapered ends and the thread's number of products): every
         * a + p then rounds to the same fixed grid, the part of p that a took, (a + p) - a, is exact, and so
         * is the rest p - that (a's rounding error; |a| > |p|: three operations where the symmetric TwoSum
         * needs six).  The rests are far below the grid and add up in an ordinary double, as the error
         * terms of the double-double form did.  (sum, rests) leave the loop as the same kind of pair. */
        const uint32_t wm = b.sh.wmax_hi;
        double bias = 1.5;
        if ((wm >> 20) != 0) {
            const int nprod = (Q + 1) * E;
            const int k = 2 * ((int)(wm >> 20) - 1023 + 1) + (32 - __clz(nprod)) + 1;
            bias = __hiloint2double((int)(((uint32_t)(k + 1023) << 20) | 0x80000u), 0);
        }
        for (int lag0 = 0; lag0 <= L; lag0 += LG) {
            double s[LG], c[LG];
#pragma unroll
            for (int l = 0; l < LG; l++) s[l] = bias, c[l] = 0.0;
            for (int tau = b.tid; tau < Tmain; tau += b.B) {
                const int t0 = run_pos(tau); /* a thread's reads stay inside one range */
                double wa[E], wb[E + LG - 1];
#pragma unroll
                for (int e = 0; e < E; e++) wa[e] = wxs[(t0 + e) + ((t0 + e) >> 3)];
#pragma unroll
                for (int j = 0; j < E + LG - 1; j++) wb[j] = wxs[(t0 + lag0 + j) + ((t0 + lag0 + j) >> 3)];
#pragma unroll
                for (int l = 0; l < LG; l++) {
                    if (lag0 + l <= L) {
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double pr = __dmul_rn(wa[e], wb[e + l]);
                            const double t = __dadd_rn(s[l], pr);
                            c[l] = __dadd_rn(c[l], __dsub_rn(pr, __dsub_rn(t, s[l])));
                            s[l] = t;
                        }
                    }
                }
            }
#pragma unroll
            for (int l = 0; l < LG; l++) s[l] = __dsub_rn(s[l], bias);
            warp_sum4(s, c, lag0);
        }
        if (R) {
            __syncwarp(); /* this warp's partial sums are in part_d */
            for (int j = b.warp; j < R; j += b.nwarps) {
                const int t0 = run_pos(Q * b.B + j);
                for (int lag = b.lane; lag <= L; lag += 32) {
                    double ss = 0.0, cc = 0.0;
#pragma unroll
                    for (int e = 0; e < E; e++)
                        dd_add(ss, cc, __dmul_rn(wxs[(t0 + e) + ((t0 + e) >> 3)], wxs[(t0 + e + lag) + ((t0 + e + lag) >> 3)]));
                    double2 v = part_d[b.warp * (FLACB_MAX_LPC + 1) + lag];
                    dd_add(v.x, v.y, ss);
                    v.y = __dadd_rn(v.y, cc);
                    part_d[b.warp * (FLACB_MAX_LPC + 1) + lag] = v;
                }
            }
        }
    } else {
        /* the windowed ends do not fit the scratch area (very large blocks): windowed on the fly */
        for (int lag0 = 0; lag0 <= L; lag0 += LG) {
            double s[LG], c[LG];
#pragma unroll
            for (int l = 0; l < LG; l++) s[l] = 0.0, c[l] = 0.0;
            for (int tile = 0; tile < b.ntiles; tile++) {
                const int i0 = (tile * b.B + b.tid) * E;
                if (i0 < n && !autoc_flat(b, i0, L)) {
                    double wa[E], wb[E + LG - 1];
#pragma unroll
                    for (int e = 0; e < E; e++)
                        wa[e] = __dmul_rn((double)x[swz((uint32_t)(GUARD + i0 + e))], b.win[i0 + e]);
#pragma unroll
                    for (int j = 0; j < E + LG - 1; j++) {
                        const int idx = i0 + lag0 + j;
                        wb[j] = __dmul_rn((double)x[swz((uint32_t)(GUARD + idx))], b.win[idx]);
                    }
#pragma unroll
                    for (int l = 0; l < LG; l++) {
                        const int lag = lag0 + l;
                        if (lag <= L) {
#pragma unroll
                            for (int e = 0; e < E; e++) {
                                /* range(len - lag - 1): the last product is dropped (encoder.py:449) */
                                if (i0 + e < n - lag - 1) dd_add(s[l], c[l], __dmul_rn(wa[e], wb[e + l]));
                            }
                        }
                    }
                }
            }
            warp_sum4(s, c, lag0);
        }
    }
    __syncthreads();
    /* ---- a warp per lag combines the warps' partial sums; rounded once at the end ---- */
    for (int lag = b.warp; lag <= L; lag += b.nwarps) {
        double S = 0.0, C = 0.0;
        if (b.lane < b.nwarps) {
            const double2 v = part_d[b.lane * (FLACB_MAX_LPC + 1) + lag];
            S = v.x, C = v.y;
            /* the exact integer part, split into two doubles that each hold it exactly */
            const i64 iv = part_i[b.lane * (FLACB_MAX_LPC + 1) + lag];
            const i64 hi = iv & ~(i64)0x7FF, lo = iv & (i64)0x7FF;
            dd_add(S, C, (double)hi);
            dd_add(S, C, (double)lo);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double s2 = __shfl_xor_sync(FULL, S, off);
            const double c2 = __shfl_xor_sync(FULL, C, off);
            dd_add(S, C, s2);
            C = __dadd_rn(C, c2);
        }
        if (b.lane == 0) b.autoc[lag] = __dadd_rn(S, C);
    }
    if (b.tid == 0) b.sh.wmax_hi = 0; /* for the next channel: every thread has read it long ago */
    __syncthreads();
}

/* ------------------------------------------------------------------------------------- */
/* encoder.py:453-479 levinson_durbin, run once to the maximum order: the coefficient vector
 * after iteration k equals levinson_durbin(autoc[:k+2]) bit for bit.  Strict IEEE operation
 * order; `lambda ** 2` is evaluated as lambda*lambda (glibc pow differs by 1 ulp on ~0.08% of
 * arguments; documented in DESIGN.md).  One warp.                                            */

static __device__ void levinson(Blk &b, int L)
{
    /* warp 0: lane j keeps coefs[j] (j < 32; coefs[32] only appears in the last iteration of
     * L = 32 and is handled on the side).  The products of the lambda sum are formed in parallel,
     * the sum itself runs in the reference's order (every partial sum is rounded), the symmetric
     * pair update is one step per lane. */
    const int lane = b.lane;
    double cj = lane == 0 ? 1.0 : 0.0; /* coefs[lane] */
    double error = b.autoc[0];
    int stat = 0;
#pragma unroll 1
    for (int k = 0; k < L; k++) {
        if (lane <= k) b.lev_prod[lane] = __dmul_rn(cj, b.autoc[k + 1 - lane]);
        __syncwarp();
        double lambda = 0.0;
#pragma unroll 4
        for (int j = 0; j <= k; j++) lambda = __dsub_rn(lambda, b.lev_prod[j]); /* encoder.py:467 */
        __syncwarp();
        if (error == 0.0) { stat = FLACB_E_ZERODIV; break; }                                /* encoder.py:469 */
        lambda = __ddiv_rn(lambda, error);
        /* encoder.py:471-474: coefs[i] <- coefs[i] + lambda * coefs[k+1-i] for i in 0..k+1 */
        const int partner = k + 1 - lane;
        double cp = __shfl_sync(FULL, cj, partner & 31);
        if (partner < 0 || partner > 31) cp = 0.0; /* coefs[32] is still zero when it is read */
        const double c0 = __shfl_sync(FULL, cj, 0);
        if (lane <= k + 1) cj = __dadd_rn(cj, __dmul_rn(lambda, cp));
        const double c32 = __dadd_rn(0.0, __dmul_rn(lambda, c0)); /* new coefs[32] when k + 1 == 32 */
        const double l2 = __dmul_rn(lambda, lambda);
        if (isinf(l2) && !isinf(lambda)) { stat = FLACB_E_OVERFLOW; break; }
        error = __dmul_rn(error, __dsub_rn(1.0, l2));                                       /* encoder.py:476 */
        if (lane >= 1 && lane <= k + 1) b.ldc[k * b.ldL + lane - 1] = cj;
        if (k + 1 == 32 && lane == 0) b.ldc[k * b.ldL + 31] = c32;
    }
    if (lane == 0) *b.lev_stat = stat;
}

/* encoder.py:482-534 quantize_lpc_coefficients for order m (one thread per order) */
static __device__ void quantize_order(Blk &b, int m, int precision)
{
    const double *c = b.ldc + (m - 1) * b.ldL;
    double cmax = 0.0;
    bool any_nan = false;
    for (int i = 0; i < m; i++) {
        const double a = fabs(c[i]);
        if (a != a) any_nan = true;
        if (a > cmax) cmax = a;
    }
    b.qncoef[m - 1] = 0;
    b.qshift[m - 1] = 0;
    b.qabs[m - 1] = 0;
    b.qstat[m - 1] = 0;
    if (any_nan || !(cmax > 0.0)) { b.qstat[m - 1] = FLACB_E_ASSERT; return; } /* encoder.py:496 */
    if (isinf(cmax)) { b.qstat[m - 1] = FLACB_E_OVERFLOW; return; }
    int shift = precision - ilogb(cmax) - 2; /* floor(log2(cmax)) == exponent, encoder.py:503 */
    if (shift > 15) shift = 15;
    else if (shift < -16) { b.qstat[m - 1] = FLACB_E_ASSERT; return; } /* encoder.py:508 */
    if (shift < 0) return; /* encoder.py:523-532: loop never appends -> ([], 0) */
    const i64 qmax = ((i64)1 << (precision - 1)) - 1;
    const i64 qmin = -((i64)1 << (precision - 1));
    const double scale = (double)((i64)1 << shift);
    double error = 0.0;
    uint32_t asum = 0;
    for (int i = 0; i < m; i++) {
        error = __dadd_rn(error, __dmul_rn(c[i], scale));
        const double r = rint(error); /* round(): half to even */
        const i64 qi = r < (double)qmin ? qmin : (r > (double)qmax ? qmax : (i64)r);
        error = __dsub_rn(error, (double)qi);
        b.qc[(m - 1) * b.ldL + i] = (int32_t)qi;
        asum += (uint32_t)(qi < 0 ? -qi : qi);
    }
    b.qncoef[m - 1] = m;
    b.qshift[m - 1] = shift;
    b.qabs[m - 1] = asum;
}

/* ------------------------------------------------------------------------------------- */
/* encoder.py:387-398: sum |residual| of every LPC order 1..L (prediction_residual,
 * encoder.py:537-548).  History window in registers, coefficients from shared memory.
 * Built for the instruction cache:
 * the orders run in a rolled loop and share a few fixed-width filter bodies (width = order
 * rounded up, coefficient rows zero padded), so the hot code is a few KB that every warp runs
 * L times per tile instead of 100+ KB of straight-line code that every warp runs once.
 * FLAVOUR 0: 16-bit samples and coefficients of at most 8 bits -> two multiply-adds per
 *            instruction (dp2a on sample pairs / coefficient bytes), 32-bit sums are exact;
 * FLAVOUR 1: 32-bit sums proven exact for every order on the host (L * 2^(P-1) * 2^(bps-1) <= 2^30);
 * FLAVOUR 2: 32-bit or 64-bit sums per order: 32-bit whenever sum|q| * 2^(bps-1) <= 2^30 proves
 *            them exact.
 * (The first version unrolled every order: 100+ KB of straight-line code per kernel, and the
 * instruction cache misses were the largest stall reason of the whole encode.)               */

#ifdef FLACB_CPU_SIM
static inline int __dp2a_lo(int a, int bq, int c)
{
    return c + (int)(int16_t)(a & 0xFFFF) * (int)(int8_t)(bq & 0xFF) + (int)(int16_t)((unsigned)a >> 16) * (int)(int8_t)((bq >> 8) & 0xFF);
}
static inline int __dp2a_hi(int a, int bq, int c)
{
    return c + (int)(int16_t)(a & 0xFFFF) * (int)(int8_t)((bq >> 16) & 0xFF) + (int)(int16_t)((unsigned)a >> 16) * (int)(int8_t)((bq >> 24) & 0xFF);
}
#endif

/* sum |x[i] - (pred >> shift)| over the thread's E samples; lo/hi: first / one past the last valid
 * sample of the run (0 / E for interior runs) */
template <class ACC, bool TIGHT = false>
__device__ __forceinline__ u64 lpc_abs_sum(const int (&xe)[E], const ACC (&acc)[E], int shift, bool interior, int lo, int hi)
{
    u64 tot = 0;
    if (sizeof(ACC) == 4 && interior && TIGHT) {
        /* 16-bit samples against coefficient bytes: |r| < 2^28, the run's eight fit 32 bits; |a - b| + c is
         * one instruction (VABSDIFF) */
        uint32_t t32 = 0;
#pragma unroll
        for (int e = 0; e < E; e++) t32 = __sad(xe[e], (int)(acc[e] >> shift), t32);
        tot = t32;
    } else if (sizeof(ACC) == 4 && interior) {
#pragma unroll
        for (int e = 0; e < E; e += 2) { /* |r| < 2^31 each: pairs fit 32 bits */
            uint32_t t32 = __sad(xe[e], (int)(acc[e] >> shift), 0u);
            t32 = __sad(xe[e + 1], (int)(acc[e + 1] >> shift), t32);
            tot += (u64)t32;
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; e++) {
            if (e >= lo && e < hi) {
                const i64 r = (i64)xe[e] - (i64)(acc[e] >> shift);
                tot += (u64)(r < 0 ? -r : r);
            }
        }
    }
    return tot;
}

/* one order with a filter body of width W (W >= order, coefficient row zero padded to W) */
template <int LMAX, int W, class ACC>
__device__ __forceinline__ u64 lpc_body(const int (&h)[LMAX + E], const int32_t *q, int shift, bool interior, int lo, int hi)
{
    ACC acc[E];
#pragma unroll
    for (int e = 0; e < E; e++) acc[e] = 0;
#pragma unroll
    for (int j4 = 0; j4 < W; j4 += 4) {
        const int4 c4 = *(const int4 *)(q + j4);
        const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const ACC cj = (ACC)cc[jj];
#pragma unroll
            for (int e = 0; e < E; e++) acc[e] += (ACC)h[LMAX + e - 1 - (j4 + jj)] * cj;
        }
    }
    int xe[E];
#pragma unroll
    for (int e = 0; e < E; e++) xe[e] = h[LMAX + e];
    return lpc_abs_sum<ACC>(xe, acc, shift, interior, lo, hi);
}

/* dp2a flavour: pk[k] = (h[k] << 16) | (h[k + 1] & 0xFFFF); a coefficient word holds q[4t..4t+3] as
 * bytes; pair (q[2t], q[2t+1]) meets samples (x[i-1-2t], x[i-2-2t]) = (h[k+1], h[k]), k = LMAX+e-2-2t */
template <int LMAX, int W>
__device__ __forceinline__ u64 lpc_body_dp2a(const int (&pk)[LMAX + E - 1], const int (&xe)[E], const uint32_t *q8, int shift,
                                             bool interior, int lo, int hi)
{
    int acc[E];
#pragma unroll
    for (int e = 0; e < E; e++) acc[e] = 0;
#pragma unroll
    for (int t4 = 0; t4 < W / 4; t4++) {
        const int cw = (int)q8[t4];
#pragma unroll
        for (int e = 0; e < E; e++) {
            acc[e] = __dp2a_lo(pk[LMAX + e - 2 - 4 * t4], cw, acc[e]);
            acc[e] = __dp2a_hi(pk[LMAX + e - 4 - 4 * t4], cw, acc[e]);
        }
    }
    return lpc_abs_sum<int, true>(xe, acc, shift, interior, lo, hi);
}

template <int LMAX, int FLAVOUR>
__device__ void lpc_errors_rolled(Blk &b, const int32_t *x, int L, const uint32_t *q8)
{
    Shared &sh = b.sh;
    const int n = b.n;
    constexpr int WSTEP = LMAX <= 12 ? 4 : 8; /* body widths: multiples of WSTEP */
    constexpr int Q8W = LMAX / 4;              /* packed coefficient words per order */
    u64 lane_tot = 0; /* lane (m-1) % 32 of every warp keeps the warp's running total of order m */
    for (int tile = 0; tile < b.ntiles; tile++) {
        const int i0 = (tile * b.B + b.tid) * E;
        const bool interior = i0 >= LMAX && i0 + E <= n;
        int h[LMAX + E];
        if (i0 < n) {
            ld_run<LMAX + E>(x, (uint32_t)(GUARD + i0 - LMAX), h);
        } else {
#pragma unroll
            for (int j = 0; j < LMAX + E; j++) h[j] = 0;
        }
        int pk[LMAX + E - 1], xe[E];
        if (FLAVOUR == 0) {
#pragma unroll
            for (int k = 0; k < LMAX + E - 1; k++) pk[k] = (int)__byte_perm((uint32_t)h[k + 1], (uint32_t)h[k], 0x5410);
#pragma unroll
            for (int e = 0; e < E; e++) xe[e] = h[LMAX + e];
        }
        const int hi = n - i0 < E ? (n - i0 < 0 ? 0 : n - i0) : E;
#pragma unroll 1
        for (int m = 1; m <= L; m++) {
            if (b.qncoef[m - 1] == 0) continue; /* block-uniform */
            const int shift = b.qshift[m - 1];
            const int lo = m - i0 > 0 ? m - i0 : 0; /* warm-up samples carry no residual */
            u64 tot = 0;
            const int wq = (m + WSTEP - 1) / WSTEP;
            if (FLAVOUR == 0) {
                const uint32_t *q = q8 + (m - 1) * Q8W;
                switch (wq) {
                case 1: tot = lpc_body_dp2a<LMAX, WSTEP>(pk, xe, q, shift, interior, lo, hi); break;
                case 2: tot = lpc_body_dp2a<LMAX, 2 * WSTEP>(pk, xe, q, shift, interior, lo, hi); break;
                case 3: tot = lpc_body_dp2a<LMAX, 3 * WSTEP>(pk, xe, q, shift, interior, lo, hi); break;
                default:
                    if (LMAX > 12) tot = lpc_body_dp2a<LMAX, LMAX>(pk, xe, q, shift, interior, lo, hi);
                    break;
                }
            } else {
                const int32_t *q = b.qc + (m - 1) * b.ldL;
                const bool small = FLAVOUR == 1 || ((u64)b.qabs[m - 1] << (b.p.bps - 1)) <= (1ull << 30);
                if (small) {
                    switch (wq) {
                    case 1: tot = lpc_body<LMAX, WSTEP, int>(h, q, shift, interior, lo, hi); break;
                    case 2: tot = lpc_body<LMAX, 2 * WSTEP, int>(h, q, shift, interior, lo, hi); break;
                    case 3: tot = lpc_body<LMAX, 3 * WSTEP, int>(h, q, shift, interior, lo, hi); break;
                    default:
                        if (LMAX > 12) tot = lpc_body<LMAX, LMAX, int>(h, q, shift, interior, lo, hi);
                        break;
                    }
                } else if (FLAVOUR == 2) {
                    switch (wq) {
                    case 1: tot = lpc_body<LMAX, WSTEP, i64>(h, q, shift, interior, lo, hi); break;
                    case 2: tot = lpc_body<LMAX, 2 * WSTEP, i64>(h, q, shift, interior, lo, hi); break;
                    case 3: tot = lpc_body<LMAX, 3 * WSTEP, i64>(h, q, shift, interior, lo, hi); break;
                    default:
                        if (LMAX > 12) tot = lpc_body<LMAX, LMAX, i64>(h, q, shift, interior, lo, hi);
                        break;
                    }
                }
            }
            /* a thread's sum is below 2^36: two reductions of 20-bit / 16-bit pieces are exact */
            const uint32_t p0 = (uint32_t)tot & 0xFFFFFu, p1 = (uint32_t)(tot >> 20);
            const u64 wt = (u64)__reduce_add_sync(FULL, p0) + ((u64)__reduce_add_sync(FULL, p1) << 20);
            if (b.lane == ((m - 1) & 31)) lane_tot += wt;
        }
    }
    b.red[b.warp * RED_STRIDE + b.lane] = lane_tot;
    __syncthreads();
    if (b.tid < L) {
        u64 t = 0;
        for (int w = 0; w < b.nwarps; w++) t += b.red[w * RED_STRIDE + b.tid];
        sh.lpc_err[b.tid] = t;
    }
    __syncthreads();
}

/* ------------------------------------------------------------------------------------- */
/* Winner's residual (encoder.py:537-548), zig-zag (utils.py:91-94) into res[], plus the
 * running sum cum[g] = sum of res[0 .. g*E) that the Rice search reads partition sums from */

template <int LMAX>
__device__ void residual_and_cum(Blk &b, const int32_t *x, const Decision &d, int c)
{
    /* In place: residual i takes the word of sample i (b.res == x + GUARD), so a frame needs no
     * second array.  A run only reads samples at or in front of itself, so the tiles are walked
     * from the back: when tile k is replaced (after a barrier: every thread holds its inputs in
     * registers), the samples tile k - 1 still needs lie in front of it.  The warm-up samples,
     * which the serialiser writes verbatim, are put aside first. */
    const int n = b.n;
    bool range_bad = false;
    if (b.tid < d.warm && b.tid < FLACB_MAX_LPC) b.sh.warm[c][b.tid] = x[swz((uint32_t)(GUARD + b.tid))];
    for (int tile = b.ntiles - 1; tile >= 0; tile--) {
        const int i0 = (tile * b.B + b.tid) * E;
        int a[E + 4], h[LMAX + E];
        if (i0 < n) {
            if (d.type == 0) ld_run<E + 4>(x, (uint32_t)(GUARD + i0 - 4), a);
            else ld_run<LMAX + E>(x, (uint32_t)(GUARD + i0 - LMAX), h);
        }
        __syncthreads();
        if (i0 < n) {
            int u[E];
            if (d.type == 0) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    int r[5];
                    fixed_residuals(a, e, r);
                    const int rr = d.order == 0 ? r[0] : d.order == 1 ? r[1] : d.order == 2 ? r[2]
                                   : d.order == 3 ? r[3] : r[4];
                    u[e] = (int)(((uint32_t)rr << 1) ^ (uint32_t)(rr >> 31));
                }
            } else {
                /* four samples at a time: half the 64-bit accumulators live at once */
#pragma unroll
                for (int e0 = 0; e0 < E; e0 += 4) {
                    i64 acc[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[e] = 0;
#pragma unroll
                    for (int j = 0; j < LMAX; j++) {
                        if (j < d.ncoef) {
                            const i64 cj = d.q[j];
#pragma unroll
                            for (int e = 0; e < 4; e++) acc[e] += (i64)h[LMAX + e0 + e - 1 - j] * cj;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const i64 r = (i64)h[LMAX + e0 + e] - (acc[e] >> d.shift);
                        const u64 z = ((u64)r << 1) ^ (u64)(r >> 63);
                        if ((z >> 32) != 0 && i0 + e0 + e >= d.warm && i0 + e0 + e < n) range_bad = true;
                        u[e0 + e] = (int)(uint32_t)z;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < E; e++)
                if (i0 + e < d.warm || i0 + e >= n) u[e] = 0;
            st_run<E>((int32_t *)b.res, (uint32_t)i0, u);
        }
    }
    /* running sums, front to back: every thread adds up its own runs again */
    u64 base = 0;
    for (int tile = 0; tile < b.ntiles; tile++) {
        const int g = tile * b.B + b.tid;
        const int i0 = g * E;
        u64 run = 0;
        if (i0 < n) {
            int u[E];
            ld_run<E>((const int32_t *)b.res, (uint32_t)i0, u);
#pragma unroll
            for (int e = 0; e < E; e++) run += (uint32_t)u[e];
        }
        u64 total;
        const u64 ex = block_excl_scan_u64(b, run, total);
        if (i0 <= n + E) b.cum[g] = base + ex;
        base += total;
    }
    if (range_bad) b.sh.range_flag = 1;
    if (b.tid == 0) {
        const int G = b.ntiles * b.B; /* run index just behind the last tile */
        if (G * E <= n + E) b.cum[G] = base;
    }
    __syncthreads();
    if (b.tid == 0 && b.sh.range_flag) set_status(b, FLACB_E_RANGE);
    __syncthreads();
}

/* sum of res[0 .. pos) */
__device__ __forceinline__ u64 cum_at(const Blk &b, int pos)
{
    const int g = pos / E;
    u64 s = b.cum[g];
    for (int j = g * E; j < pos; j++) s += b.res[swz((uint32_t)j)];
    return s;
}

__device__ __forceinline__ uint32_t fast_div(uint32_t i, uint32_t d, uint32_t magic)
{
    return d == 1 ? i : __umulhi(i, magic); /* exact for i, d < 2^16; magic = ceil(2^32 / d) */
}

/* ------------------------------------------------------------------------------------- */
/* encoder.py:655-760 rice_partitions: every candidate order is evaluated, parameter
 * k = floor(log2(sum / len)) per partition (find_rice_parameter), size estimate with the
 * reference's per-partition "+4" (estimate_rice_partition_size_and_parameter), first minimum.
 * Returns through d.porder / d.method and the exact number of residual bits in *res_bits.
 * keep_params: leave kall[] valid for the chosen order (phase 3).                           */

static __device__ void rice_search(Blk &b, Decision &d, uint32_t *res_bits)
{
    Shared &sh = b.sh;
    const EncParams &p = b.p;
    const int n = b.n;
    if (b.tid == 0) {
        int nc = 0;
        for (int o = p.rmin; o <= p.rmax; o++) { /* encoder.py:664-667 */
            if ((n % (1 << o)) == 0 && (n >> o) > d.warm) {
                sh.cand_o[nc] = (uint32_t)o;
                const uint32_t psz = (uint32_t)(n >> o);
                sh.cand_magic[nc] = psz == 1 ? 0u : 0xFFFFFFFFu / psz + 1u;
                nc++;
            }
        }
        if (nc == 0) {
            /* encoder.py:669 assert: no admissible partition order.  Status recorded; the
             * frame is still laid out with order 0 so that the stream stays well formed. */
            set_status(b, FLACB_E_ASSERT);
            sh.cand_o[0] = 0;
            sh.cand_magic[0] = n == 1 ? 0u : 0xFFFFFFFFu / (uint32_t)n + 1u;
            nc = 1;
        }
        sh.ncand = nc;
        sh.errkey = 0xFFFFFFFFu;
    }
    if (b.tid < MAX_CAND) sh.sizeA[b.tid] = 0, sh.sizeB[b.tid] = 0, sh.any5[b.tid] = 0, sh.qsum[b.tid] = 0;
    __syncthreads();
    const int ncand = sh.ncand;
    /* parameters of every (order, partition) */
    for (int ci = 0; ci < ncand; ci++) {
        const int o = (int)sh.cand_o[ci];
        const int psz = n >> o;
        uint32_t a_loc = 0, b_loc = 0, f5 = 0;
        for (int pi = b.tid; pi < (1 << o); pi += b.B) {
            const int start = pi * psz < d.warm ? d.warm : pi * psz;
            const int end = (pi + 1) * psz;
            const uint32_t len = (uint32_t)(end - start);
            const u64 sum = cum_at(b, end) - cum_at(b, start);
            int k;
            if (sum < (u64)len) {
                /* mean < 1: math domain error (sum == 0, encoder.py:753) or a negative shift
                 * count (encoder.py:758).  Lenient policy: parameter 0. */
                atomicMin(&sh.errkey, ((uint32_t)ci << 20) | ((uint32_t)pi << 4) | FLACB_E_VALUE);
                k = 0;
            } else {
                k = (63 - __clzll(sum)) - (31 - __clz(len));
                if (((u64)len << k) > sum) k--;
            }
            b.kall[(1u << o) + (uint32_t)pi] = (uint8_t)k;
            a_loc += 4u + (k > 14 ? 5u : 4u); /* encoder.py:718-725 */
            b_loc += len * (uint32_t)(1 + k);
            if (k > 14) f5 = 1;
        }
        if (a_loc) {
            atomicAdd(&sh.sizeA[ci], a_loc);
            atomicAdd(&sh.sizeB[ci], b_loc);
            if (f5) atomicOr(&sh.any5[ci], 1u);
        }
    }
    __syncthreads();
    /* sum over samples of (u >> k) for every candidate order (rice_size, encoder.py:756-760);
     * u >> k < 2^29 for bps <= 24, so a thread's tile sum fits 32 bits.  The candidate orders are
     * consecutive (both admission tests are monotone in the order) and their partitions nest, so
     * a run that lies in one partition of the finest order lies in one partition of every order,
     * whose index is the finest one shifted: one pass over the residual serves up to RG orders. */
    u64 lane_q = 0;
    constexpr int RG = 6;
    const int o_last = (int)sh.cand_o[ncand - 1];
    const uint32_t psz_f = (uint32_t)n >> o_last, magic_f = sh.cand_magic[ncand - 1];
#pragma unroll 1
    for (int g0 = 0; g0 < ncand; g0 += RG) {
        u64 acc[RG];
#pragma unroll
        for (int j = 0; j < RG; j++) acc[j] = 0;
#pragma unroll 1
        for (int tile = 0; tile < b.ntiles; tile++) {
            const int i0 = (tile * b.B + b.tid) * E;
            if (i0 < n) {
                int ui[E];
                ld_run<E>((const int32_t *)b.res, (uint32_t)i0, ui);
                uint32_t u[E];
#pragma unroll
                for (int e = 0; e < E; e++) u[e] = (uint32_t)ui[e];
                const uint32_t p0 = fast_div((uint32_t)i0, psz_f, magic_f);
                if ((uint32_t)(i0 + E) <= (p0 + 1) * psz_f) { /* the run lies in one partition (the usual case) */
#pragma unroll
                    for (int j = 0; j < RG; j++) {
                        if (g0 + j < ncand) {
                            const int o = (int)sh.cand_o[g0 + j];
                            const uint32_t k0 = b.kall[(1u << o) + (p0 >> (o_last - o))];
                            uint32_t t32 = 0;
#pragma unroll
                            for (int e = 0; e < E; e++) t32 += u[e] >> k0;
                            acc[j] += t32;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < RG; j++) {
                        if (g0 + j < ncand) {
                            const uint32_t o = sh.cand_o[g0 + j];
                            const uint32_t psz = (uint32_t)n >> o, magic = sh.cand_magic[g0 + j];
                            const uint8_t *kk = b.kall + (1u << o);
                            uint32_t t32 = 0;
#pragma unroll
                            for (int e = 0; e < E; e++)
                                if (i0 + e < n) t32 += u[e] >> kk[fast_div((uint32_t)(i0 + e), psz, magic)];
                            acc[j] += t32;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < RG; j++) {
            if (g0 + j < ncand) { /* block-uniform */
                const u64 t = warp_sum_u64(acc[j]);
                if (b.lane == g0 + j) lane_q = t;
            }
        }
    }
    if (b.lane < MAX_CAND) b.red[b.warp * RED_STRIDE + b.lane] = lane_q;
    __syncthreads();
    if (b.tid < ncand) {
        u64 t = 0;
        for (int w = 0; w < b.nwarps; w++) t += b.red[w * RED_STRIDE + b.tid];
        sh.qsum[b.tid] = t;
    }
    __syncthreads();
    if (b.tid == 0) {
        int best = 0;
        u64 best_size = 0;
        for (int ci = 0; ci < ncand; ci++) { /* first minimum, encoder.py:685-695 */
            const u64 size = (u64)sh.sizeA[ci] + (u64)sh.sizeB[ci] + sh.qsum[ci];
            if (ci == 0 || size < best_size) best = ci, best_size = size;
        }
        d.porder = (int)sh.cand_o[best];
        d.method = sh.any5[best] ? 5 : 4; /* encoder.py:648-650 */
        /* method (2) + partition order (4) + one parameter per partition + the codes */
        sh.sizeA[0] = 6u + ((uint32_t)d.method << d.porder) + sh.sizeB[best] + (uint32_t)sh.qsum[best];
        if (sh.errkey != 0xFFFFFFFFu) set_status(b, (int)(sh.errkey & 15u));
    }
    __syncthreads();
    *res_bits = sh.sizeA[0];
    __syncthreads();
}

/* ------------------------------------------------------------------------------------- */
/* bit packer (binary.py:144-216 Put: MSB first)                                          */

/* serial put by one thread into the staging window */
__device__ __forceinline__ void put_serial(Blk &b, uint32_t v, int nb)
{
    Shared &sh = b.sh;
    const uint32_t off = sh.pos & 31u;
    const uint32_t w = (sh.pos >> 5) - sh.wbase;
    const u64 xx = (u64)v << (64 - (int)off - nb);
    b.stg[w] |= (uint32_t)(xx >> 32);
    if ((uint32_t)xx) b.stg[w + 1] |= (uint32_t)xx;
    sh.pos += (uint32_t)nb;
}

__device__ __forceinline__ uint32_t gf16_mul(uint32_t a, uint32_t bb)
{
    uint32_t r = 0; /* polynomial product modulo 0x18005 (common.py:13) */
#pragma unroll
    for (int i = 15; i >= 0; i--) {
        r <<= 1;
        if (r & 0x10000u) r ^= 0x18005u;
        if ((bb >> i) & 1u) r ^= a;
    }
    return r;
}

/* phase 2, deferred: byte offset of this frame = sum of the sizes of all frames in front of it.
 * Decoupled look-back, one warp, 32 predecessors per probe: sizes are summed backwards until a
 * frame that already knows its own inclusive prefix is met. */
static __device__ __noinline__ void resolve_offset(Shared *shp, int tid)
{
    /* out of line and on plain arguments: called from every flush site, and a reference to the
     * per-thread block state would force that state into local memory */
    Shared &sh = *shp;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t f = sh.frame;
    const u64 slot_stride = sh.o_slot_stride;
    u64 *lb_state = sh.o_lb_state;
    if (slot_stride) {
        if (tid == 0) sh.frame_off = (u64)f * slot_stride;
    } else if (warp == 0) {
        const u64 mine = (u64)sh.frame_bytes;
        u64 excl = 0;
        bool failed = false;
        if (f > 0) {
            int64_t j = (int64_t)f - 1;
            uint32_t spins = 0;
            while (true) {
                const int64_t idx = j - lane;
                u64 s = LB_PREFIX; /* before frame 0: an empty prefix */
                if (idx >= 0) s = *(volatile u64 *)&lb_state[idx];
                const uint32_t flag = (uint32_t)(s >> 62);
                const unsigned has_prefix = __ballot_sync(FULL, flag == 2);
                const unsigned invalid = __ballot_sync(FULL, flag == 0);
                const int first_prefix = has_prefix ? __ffs((int)has_prefix) - 1 : 32;
                const unsigned need = first_prefix == 32 ? FULL : ((2u << first_prefix) - 1u);
                if (invalid & need) { /* a predecessor has not published yet */
                    /* predecessors are resident (ticket order) and will publish; the bound only catches a
                     * dead predecessor, after minutes, not a time-sliced or profiled GPU */
                    if (++spins > (1u << 30)) { failed = true; break; }
                    __nanosleep(spins < 64 ? 64 : 256);
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
        }
    }
    __syncthreads();
    if (tid == 0) {
        const u64 out_cap = sh.o_cap;
        const u64 lim = slot_stride ? slot_stride : out_cap - (sh.frame_off < out_cap ? sh.frame_off : out_cap);
        sh.writable = (u64)sh.frame_bytes <= lim && (slot_stride == 0 || sh.frame_off + sh.frame_bytes <= out_cap) &&
                      sh.status != FLACB_E_INTERNAL;
        if (!sh.writable && sh.status == 0) sh.status = FLACB_E_NOSPACE;
        sh.have_off = 1;
    }
    __syncthreads();
}

/* Flush the first `nbytes` bytes of the staging window to the frame in HBM and fold them into
 * the frame CRC-16 (crc.py:25-30): every thread takes a 64-byte slice, and because the CRC has
 * zero init and no final xor, crc(A || B) = crc(A) * x^(8|B|) + crc(B), so a slice's CRC times
 * x^(8 * bytes behind it) can simply be xor-ed into the frame's.  Then the window is rebased at
 * word `wend` (the partial word, if any, moves to the front). */
static __device__ __noinline__ void flush_window_impl(Shared *shp, uint32_t *stg, uint32_t stg_words, int tid, int B,
                                              uint32_t nbytes, uint32_t wend)
{
    Shared &sh = *shp;
    if (!sh.have_off) resolve_offset(shp, tid); /* block-uniform */
    const uint32_t byte0 = sh.wbase * 4u; /* frame-relative */
    if (sh.writable && nbytes) {
        /* staging words hold the stream MSB first; memory wants those bytes in order.  Aligned
         * 32-bit stores for the body, single bytes in front of / behind it. */
        uint8_t *dst = sh.o_out + sh.frame_off + byte0;
        uint32_t head = (4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u;
        if (head > nbytes) head = nbytes;
        const uint32_t nq = (nbytes - head) >> 2;
        const uint32_t tail0 = head + nq * 4u;
        for (uint32_t idx = (uint32_t)tid; idx < head + (nbytes - tail0); idx += (uint32_t)B) {
            const uint32_t i = idx < head ? idx : tail0 + (idx - head);
            dst[i] = (uint8_t)(stg[i >> 2] >> (24u - 8u * (i & 3u)));
        }
        uint32_t *dstw = (uint32_t *)(dst + head);
        const uint32_t r = head & 3u; /* byte offset of every aligned word inside the staging words */
        for (uint32_t q = (uint32_t)tid; q < nq; q += (uint32_t)B) {
            const uint32_t j = (head + 4u * q) >> 2;
            const uint32_t s0 = __byte_perm(stg[j], 0, 0x0123), s1 = __byte_perm(stg[j + 1], 0, 0x0123);
            dstw[q] = r == 0 ? s0 : __funnelshift_r(s0, s1, 8u * r);
        }
    }
    const uint32_t nslices = (nbytes + CRC_SLICE - 1) / CRC_SLICE;
    for (uint32_t sl = (uint32_t)tid; sl < nslices; sl += (uint32_t)B) {
        const uint32_t lo = sl * CRC_SLICE;
        const uint32_t hi = lo + CRC_SLICE < nbytes ? lo + CRC_SLICE : nbytes;
        uint32_t c = 0;
        for (uint32_t idx = lo; idx < hi; idx++) {
            const uint32_t byte = (stg[idx >> 2] >> (24u - 8u * (idx & 3u))) & 0xFFu;
            c = ((c << 8) ^ sh.crc16_tab[(c >> 8) ^ byte]) & 0xFFFFu;
        }
        const uint32_t behind = sh.payload_bytes - (byte0 + hi);
        const uint32_t w = gf16_mul(c_xpow_lo[behind & 255u], c_xpow_hi[behind >> 8]);
        atomicXor(&sh.crc, gf16_mul(c, w));
    }
    const uint32_t carry_idx = wend - sh.wbase;
    const uint32_t carry = carry_idx <= stg_words ? stg[carry_idx] : 0u;
    /* nothing beyond the word the bit cursor is in has been written */
    uint32_t used = (sh.pos >> 5) - sh.wbase;
    if (used < (nbytes + 3u) / 4u) used = (nbytes + 3u) / 4u; /* a tile wider than the window is flushed mid-way */
    used += 2u;
    if (used > stg_words + 4u) used = stg_words + 4u;
    __syncthreads();
    for (uint32_t idx = (uint32_t)tid; idx < used; idx += (uint32_t)B) stg[idx] = idx == 0 ? carry : 0u;
    if (tid == 0) sh.wbase = wend;
    __syncthreads();
}

__device__ __forceinline__ void flush_window(Blk &b, uint32_t nbytes, uint32_t wend)
{
    flush_window_impl(&b.sh, b.stg, b.stg_words, b.tid, b.B, nbytes, wend);
}

/* or `nb` bits (nb <= 32) into the staging window at frame bit position bp, MSB first
 * (binary.py:144-216).  The window is zero before a frame's bits arrive, so fields only ever set
 * bits: a shared-memory atomic or per touched word, no ownership protocol between the threads
 * whose codes share a word.  CHECK: only words inside the window [wbase, wbase + words). */
template <bool CHECK>
__device__ __forceinline__ void or_bits(uint32_t *stg, uint32_t wbase, uint32_t words, uint32_t bp, uint32_t v, int nb)
{
    const uint32_t wi = (bp >> 5) - wbase;
    const u64 x = (u64)v << (64 - (int)(bp & 31u) - nb);
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    if (!CHECK || wi < words) atomicOr(&stg[wi], hi);
    if (lo && (!CHECK || wi + 1u < words)) atomicOr(&stg[wi + 1u], lo);
}

/* encoder.py:765-806 put_residual: method, partition order, then per partition the parameter
 * and one Rice code per sample (u >> k zeros, a one, k low bits).  A block-wide scan of the
 * threads' bit counts gives every thread the position of its E codes; the zeros cost nothing
 * (the window is pre-zeroed), the rest is or-ed in. */
static __device__ void pack_residual(Blk &b, const Decision &d)
{
    Shared &sh = b.sh;
    const int n = b.n;
    const uint32_t psz = (uint32_t)n >> d.porder;
    const uint32_t magic = psz == 1 ? 0u : 0xFFFFFFFFu / psz + 1u;
    const uint8_t *kk = b.kall + (1u << d.porder);
    const uint32_t method_code = d.method == 4 ? 0u : 1u;
    const uint32_t hv = (method_code << 4) | (uint32_t)d.porder; /* encoder.py:766-767 */
    /* parameter of sample i's partition and the prefix bits in front of its code: the parameter
     * field at the start of a partition, method + partition order in front of the first one */
    auto param_of = [&](int i, uint32_t &k, uint32_t &pre) {
        const uint32_t pi = fast_div((uint32_t)i, psz, magic);
        const uint32_t pstart = pi * psz < (uint32_t)d.warm ? (uint32_t)d.warm : pi * psz;
        k = kk[pi];
        pre = (uint32_t)i == pstart ? (uint32_t)d.method + (i == d.warm ? 6u : 0u) : 0u;
    };
    for (int tile = 0; tile < b.ntiles; tile++) {
        const int i0 = (tile * b.B + b.tid) * E;
        uint32_t u[E];
        uint32_t nbits = 0;
        {
            int ui[E];
            if (i0 < n) {
                ld_run<E>((const int32_t *)b.res, (uint32_t)i0, ui);
            } else {
#pragma unroll
                for (int e = 0; e < E; e++) ui[e] = 0;
            }
#pragma unroll
            for (int e = 0; e < E; e++) u[e] = (uint32_t)ui[e];
        }
        /* the usual run: inside the block, behind the warm-up, in one partition; only its first
         * sample can open the partition.  Such a thread carries one parameter over the scan; the
         * few others look theirs up again per sample. */
        const uint32_t p0 = fast_div((uint32_t)i0, psz, magic);
        const bool usual = i0 >= d.warm && i0 + E <= n && (uint32_t)(i0 + E) <= (p0 + 1) * psz;
        uint32_t k0 = 0, pre0 = 0;
        if (usual) {
            param_of(i0, k0, pre0);
#pragma unroll
            for (int e = 0; e < E; e++) nbits += (u[e] >> k0) + 1u + k0;
            nbits += pre0;
        } else {
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int i = i0 + e;
                if (i >= d.warm && i < n) {
                    uint32_t k, pre;
                    param_of(i, k, pre);
                    nbits += pre + (u[e] >> k) + 1u + k;
                }
            }
        }
        uint32_t tile_bits;
        const uint32_t ex = block_excl_scan_u32(b, nbits, tile_bits);
        if (tile_bits == 0) continue; /* uniform */
        const uint32_t tile_end = sh.pos + tile_bits;
        if ((tile_end >> 5) >= sh.wbase + b.stg_words) flush_window(b, ((sh.pos >> 5) - sh.wbase) * 4u, sh.pos >> 5);
        else __syncthreads();
        const uint32_t p_start = sh.pos + ex;
        for (;;) {
            const uint32_t wbase = sh.wbase;
            /* uniform: the whole tile lies inside the window, no range checks needed */
            const bool fits = (sh.pos >> 5) >= wbase && (tile_end >> 5) + 1u < wbase + b.stg_words;
            auto put = [&](uint32_t bp, uint32_t v, int nb) {
                if (fits) or_bits<false>(b.stg, wbase, b.stg_words, bp, v, nb);
                else or_bits<true>(b.stg, wbase, b.stg_words, bp, v, nb);
            };
            if (nbits) {
                uint32_t cur = p_start;
                if (usual) {
                    if (pre0) {
                        if (pre0 > (uint32_t)d.method) put(cur, hv, 6), cur += 6;
                        put(cur, k0, d.method); /* encoder.py:791 */
                        cur += (uint32_t)d.method;
                    }
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        cur += u[e] >> k0; /* q zeros: the window is pre-zeroed (encoder.py:802) */
                        put(cur, (1u << k0) | (u[e] & ((1u << k0) - 1u)), (int)k0 + 1);
                        cur += k0 + 1u;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const int i = i0 + e;
                        if (i >= d.warm && i < n) {
                            uint32_t k, pre;
                            param_of(i, k, pre);
                            if (pre) {
                                if (pre > (uint32_t)d.method) put(cur, hv, 6), cur += 6;
                                put(cur, k, d.method);
                                cur += (uint32_t)d.method;
                            }
                            cur += u[e] >> k;
                            put(cur, (1u << k) | (u[e] & ((1u << k) - 1u)), (int)k + 1);
                            cur += k + 1u;
                        }
                    }
                }
            }
            __syncthreads();
            if ((tile_end >> 5) < sh.wbase + b.stg_words) break;
            flush_window(b, b.stg_words * 4u, sh.wbase + b.stg_words);
        }
        if (b.tid == 0) sh.pos = tile_end;
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------- */
/* encoder.py:194-234 put_frame_header (+ coded_number.py:7-39, crc.py:17-22)              */

static __device__ int build_frame_header(const EncParams &p, int bs, u64 number, uint8_t *h)
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
    const uint32_t chn = (p.flags & FLACB_SPEC_CHANNELS) ? (uint32_t)(p.channels - 1) : 1u; /* encoder.py:95 */
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
    uint32_t crc = 0; /* crc.py:17-22, polynomial 0x107 */
    for (int i = 0; i < len; i++) {
        crc ^= h[i];
        for (int bit = 0; bit < 8; bit++) crc = (crc & 0x80u) ? ((crc << 1) ^ 0x07u) & 0xFFu : (crc << 1) & 0xFFu;
    }
    h[len++] = (uint8_t)crc;
    return len;
}

/* ------------------------------------------------------------------------------------- */
/* the kernel                                                                             */

/* Front end of a channel: fixed-predictor errors and, for LPC configurations, the
 * autocorrelation (everything the Levinson recursion needs). */
template <int LMAX, bool LPC>
__device__ void analyse_front(Blk &b, const int32_t *x)
{
    const EncParams &p = b.p;
    const int L = p.L;
    FLACB_MARK(b, 0); /* everything before: load, waits */
    fixed_errors(b, x);
    FLACB_MARK(b, 1);
    if constexpr (LPC) {
        const int bad_win = (b.sh.frame == p.n_frames - 1 && p.tail_size != p.block_size) ? p.win_tail_bad : p.win_full_bad;
        if (b.tid == 0) *b.lev_stat = bad_win ? FLACB_E_ZERODIV : 0; /* encoder.py:437 */
        if (!bad_win) autocorrelation<4, (LMAX <= 12 ? 13 : 8)>(b, x, L);
        FLACB_MARK(b, 2);
    }
}

/* encoder.py:387-404, :135-157: error of every LPC order, best order, fixed-vs-LPC choice.
 * FLAVOUR < 0: no LPC analysis in this kernel (fixed predictors only, p.L == 0). */
template <int LMAX, int FLAVOUR>
__device__ void analyse_choose(Blk &b, const int32_t *x, int c, const uint32_t *q8)
{
    Shared &sh = b.sh;
    const EncParams &p = b.p;
    const int n = b.n;
    const int L = p.L;
    Decision &d = sh.dec[c];
    flacb_subframe_info *inf = p.info ? p.info + ((u64)sh.frame * (u64)p.channels + (u64)c) : nullptr;
    bool lpc_ok = false;
    if (L > 0) {
        if (*b.lev_stat == 0) {
            FLACB_MARK(b, 4);
            if constexpr (FLAVOUR >= 0) lpc_errors_rolled<LMAX, FLAVOUR>(b, x, L, q8);
            FLACB_MARK(b, 5);
            lpc_ok = true;
        } else if (b.tid == 0) set_status(b, *b.lev_stat);
        __syncthreads();
    }
    if (b.tid == 0) {
        /* encoder.py:352 first minimum over the fixed orders (order 0 only when n <= 4) */
        int fo = 0;
        if (n > 4)
            for (int o = 1; o < 5; o++)
                if (b.fixed_err[o] < b.fixed_err[fo]) fo = o;
        const u64 ftotal = b.fixed_err[fo];
        int lo = 0;
        u64 ltotal = 0;
        if (lpc_ok) {
            for (int m = 1; m <= L; m++) { /* encoder.py:404 first minimum */
                /* negative-shift quirk: no coefficients -> residual is the signal itself */
                const u64 t = b.qncoef[m - 1] ? sh.lpc_err[m - 1] : b.fixed_err[0];
                if (m == 1 || t < ltotal) lo = m, ltotal = t;
            }
        }
        bool use_lpc = false;
        if (lpc_ok) { /* encoder.py:135-157 */
            if (ltotal < ftotal) use_lpc = true;
            else if (!(ftotal < ltotal)) set_status(b, FLACB_E_ASSERT);
        }
        if (use_lpc && (lo >= n || b.qncoef[lo - 1] == 0)) {
            /* only reachable on tails not longer than the order: the reference then fails in
             * rice_partitions (encoder.py:669) */
            set_status(b, FLACB_E_ASSERT);
            use_lpc = false;
        }
        d.type = use_lpc ? 1 : 0;
        d.order = use_lpc ? lo : fo;
        d.warm = d.order;
        d.shift = use_lpc ? b.qshift[lo - 1] : 0;
        d.ncoef = use_lpc ? lo : 0;
        if (use_lpc)
            for (int j = 0; j < lo; j++) d.q[j] = b.qc[(lo - 1) * b.ldL + j];
        if (use_lpc && p.P - 1 == 15) set_status(b, FLACB_E_ASSERT); /* encoder.py:619 */
        if (inf) {
            inf->type = d.type;
            inf->order = d.order;
            inf->shift = d.shift;
            inf->precision = use_lpc ? p.P : 0;
            inf->fixed_order = fo;
            inf->lpc_order = lpc_ok ? lo : 0;
            for (int o = 0; o < 5; o++) inf->fixed_err[o] = (n <= 4 && o > 0) ? -1 : (int64_t)b.fixed_err[o];
            for (int m = 0; m < FLACB_MAX_LPC; m++) {
                const bool have = lpc_ok && m < L;
                inf->lpc_err[m] = have ? (int64_t)(b.qncoef[m] ? sh.lpc_err[m] : b.fixed_err[0]) : 0;
                inf->lpc_shift[m] = have ? b.qshift[m] : 0;
                inf->lpc_ncoef[m] = have ? b.qncoef[m] : 0;
                for (int j = 0; j < FLACB_MAX_LPC; j++) inf->lpc_q[m][j] = (have && j < b.qncoef[m]) ? b.qc[m * b.ldL + j] : 0;
            }
            for (int l = 0; l <= FLACB_MAX_LPC; l++) inf->autoc[l] = (L > 0 && l <= L) ? b.autoc[l] : 0.0;
        }
    }
    __syncthreads();
    FLACB_MARK(b, 6); /* decision */
}

/* winner's residual, Rice partition search, exact subframe size (encoder.py:632-760) */
template <int LMAX>
__device__ void analyse_size(Blk &b, const int32_t *x, int c)
{
    Shared &sh = b.sh;
    const EncParams &p = b.p;
    Decision &d = sh.dec[c];
    flacb_subframe_info *inf = p.info ? p.info + ((u64)sh.frame * (u64)p.channels + (u64)c) : nullptr;
    residual_and_cum<LMAX>(b, x, d, c);
    FLACB_MARK(b, 7);
    uint32_t res_bits;
    rice_search(b, d, &res_bits);
    FLACB_MARK(b, 8);
    if (b.tid == 0) {
        /* encoder.py:553-627: header byte, warm-up, [precision, shift, coefficients], residual */
        uint32_t bits = 8u + (uint32_t)d.warm * (uint32_t)p.bps + res_bits;
        if (d.type == 1) bits += 4u + 5u + (uint32_t)d.ncoef * (uint32_t)p.P;
        d.bits = bits;
        if (inf) {
            inf->porder = d.porder;
            inf->method = d.method;
            inf->bits = bits;
            inf->nparams = 1 << d.porder;
            for (int i = 0; i < FLACB_INFO_PARAMS; i++) inf->params[i] = i < (1 << d.porder) ? b.kall[(1u << d.porder) + i] : 0;
        }
    }
    __syncthreads();
}

/* point the LPC front-end views of the block state at one record (shared memory or HBM) */
__device__ __forceinline__ void bind_lpc_record(Blk &b, unsigned char *q, const SmemLayout &lay, int L)
{
    b.autoc = (double *)(q + lay.o_autoc);
    b.lev_prod = (double *)(q + lay.o_levprod);
    b.ldc = (double *)(q + lay.o_ldc);
    b.qc = (int32_t *)(q + lay.o_qc);
    const int Lq = L > 0 ? L : 1;
    b.qshift = (int32_t *)(q + lay.o_qmeta);
    b.qncoef = b.qshift + Lq;
    b.qstat = b.qncoef + Lq;
    b.qabs = (uint32_t *)(b.qstat + Lq);
    b.lev_stat = (int32_t *)(b.qabs + Lq);
    b.fixed_err = (u64 *)(q + lay.o_fixed);
}

/* Levinson recursion and quantisation of one subframe by one warp (lane m-1 quantises order m),
 * then the first failing order (encoder.py:378 runs the orders in turn) */
static __device__ void lpc_coefficients_warp(Blk &b, int L, int P)
{
    if (*b.lev_stat == 0) levinson(b, L);
    __syncwarp();
    if (*b.lev_stat == 0)
        for (int m = b.lane + 1; m <= L; m += 32) quantize_order(b, m, P);
    __syncwarp();
    if (b.lane == 0) {
        int st = *b.lev_stat;
        for (int m = 0; m < L && st == 0; m++) st = b.qstat[m];
        *b.lev_stat = st;
    }
    __syncwarp();
}

/* ---- staged pipeline, stage A: fixed-predictor errors and autocorrelation of every subframe of
 * a frame -> the subframe's record in HBM.  Small shared-memory footprint (planes, windowed
 * ends), no packer state: more frames per SM than the full kernel. */
template <int LMAX, int TB, int MINB, int FIXB>
__global__ void __launch_bounds__(TB, MINB) analyse_kernel(const EncParams p)
{
    FLACB_DYN_SMEM(smem);
    __shared__ Shared sh;
    FLACB_SIM_POISON(sh);
    Blk b(p, sh);
    b.tid = (int)threadIdx.x;
    b.B = FIXB ? FIXB : (int)blockDim.x; /* FIXB: the block size as a compile-time constant (the usual 288) */
    b.lane = b.tid & 31;
    b.warp = b.tid >> 5;
    b.nwarps = b.B >> 5;
    const uint32_t f = blockIdx.x;
    pdl_launch_dependents(); /* stage B's CTAs may come in behind this stage's last ones */
    if (b.tid == 0) {
        sh.frame = f;
        sh.wmax_hi = 0;
#ifdef FLACB_PHASE_TIMING
        sh.t_last = clock64();
#endif
    }
    __syncthreads();
    b.n = (f == p.n_frames - 1) ? p.tail_size : p.block_size;
    b.ntiles = (b.n + b.B * E - 1) / (b.B * E);
    const bool tail_frame = f == p.n_frames - 1 && p.tail_size != p.block_size;
    b.win = tail_frame ? p.win_tail : p.win_full;
    b.flat_lo = tail_frame ? p.flat_tail_lo : p.flat_full_lo;
    b.flat_hi = tail_frame ? p.flat_tail_hi : p.flat_full_hi;
    const SmemLayout &lay = p.lay;
    b.xs = (int32_t *)(smem + lay.xs);
    b.xwords = lay.xwords;
    b.res = (uint32_t *)(smem + lay.res);
    b.cum = (u64 *)(smem + lay.cum);
    b.kall = (uint8_t *)(smem + lay.kall);
    b.stg = (uint32_t *)(smem + lay.stg);
    b.fb = (uint32_t *)(smem + lay.fb);
    b.pst = (uint32_t *)(smem + lay.pst);
    b.red = (u64 *)(smem + lay.red);
    b.ldL = lpc_row_stride(p.L);
    b.stg_words = lay.stg_words;
    b.wx_bytes = lay.kall - lay.res;
    if (lay.xs_all) load_planes(b, -1);
    for (int c = 0; c < p.channels; c++) {
        if (!lay.xs_all) load_planes(b, c);
        bind_lpc_record(b, p.pre + ((u64)f * (u64)p.channels + (u64)c) * (u64)p.pre_stride, lay, p.L);
        __syncthreads();
        analyse_front<LMAX, true>(b, lay.xs_all ? b.xs + (uint32_t)c * b.xwords : b.xs);
    }
}

/* ---- stage C1: error of every LPC order on every sample, best order, fixed-vs-LPC choice
 * (encoder.py:387-404, :135-157) -> one SubDecision per subframe in HBM.  Same small footprint
 * as stage A plus the subframes' coefficient records; its code is the unrolled all-orders
 * filter and little else, so the instruction cache holds it. */
template <int LMAX, int TB, int MINB, int FLAVOUR, int FIXB>
__global__ void __launch_bounds__(TB, MINB) choose_kernel(const EncParams p)
{
    FLACB_DYN_SMEM(smem);
    __shared__ Shared sh;
    FLACB_SIM_POISON(sh);
    Blk b(p, sh);
    b.tid = (int)threadIdx.x;
    b.B = FIXB ? FIXB : (int)blockDim.x;
    b.lane = b.tid & 31;
    b.warp = b.tid >> 5;
    b.nwarps = b.B >> 5;
    const uint32_t f = blockIdx.x;
    if (b.tid == 0) {
        sh.frame = f;
#ifdef FLACB_PHASE_TIMING
        sh.t_last = clock64();
#endif
    }
    if (b.tid < MAX_CH) sh.chan_status[b.tid] = 0;
    if (b.tid == 0) sh.wmax_hi = 0;
    b.n = (f == p.n_frames - 1) ? p.tail_size : p.block_size;
    b.ntiles = (b.n + b.B * E - 1) / (b.B * E);
    const SmemLayout &lay = p.lay;
    b.xs = (int32_t *)(smem + lay.xs);
    b.xwords = lay.xwords;
    b.red = (u64 *)(smem + lay.red);
    b.ldL = lpc_row_stride(p.L);
    pdl_launch_dependents();
    __syncthreads(); /* sh.frame */
    /* the frame's PCM does not depend on the stages in front: it is split while they finish */
    if (lay.xs_all) load_planes(b, -1);
    pdl_wait();
    /* the records of this frame's subframes (autocorrelation only for the introspection record) */
    for (int c = 0; c < p.channels; c++) {
        const uint32_t *rec = (const uint32_t *)(p.pre + ((u64)f * (u64)p.channels + (u64)c) * (u64)p.pre_stride);
        uint32_t *loc = (uint32_t *)(smem + lay.lpc + (uint32_t)c * lay.lpc_stride);
        for (uint32_t i = (uint32_t)b.tid; i < lay.o_levprod / 4; i += (uint32_t)b.B) loc[i] = rec[i];
        for (uint32_t i = lay.o_qc / 4 + (uint32_t)b.tid; i < lay.lpc_stride / 4; i += (uint32_t)b.B) loc[i] = rec[i];
    }
    __syncthreads();
    /* coefficient rows zero padded to the row stride (the filter bodies have fixed widths); for
     * the dp2a flavour also packed as bytes, four to a word, in the (unused) Levinson matrix */
    for (int idx = b.tid; idx < p.channels * p.L * b.ldL; idx += b.B) {
        const int c = idx / (p.L * b.ldL), r = idx - c * (p.L * b.ldL);
        const int m1 = r / b.ldL, j = r - m1 * b.ldL;
        unsigned char *q = smem + lay.lpc + (uint32_t)c * lay.lpc_stride;
        const int nco = ((const int32_t *)(q + lay.o_qmeta))[(p.L > 0 ? p.L : 1) + m1]; /* qncoef[m1] */
        if (j >= nco) ((int32_t *)(q + lay.o_qc))[r] = 0;
    }
    __syncthreads();
    if (FLAVOUR == 0) {
        for (int idx = b.tid; idx < p.channels * p.L * (LMAX / 4); idx += b.B) {
            const int c = idx / (p.L * (LMAX / 4)), r = idx - c * (p.L * (LMAX / 4));
            const int m1 = r / (LMAX / 4), t4 = r - m1 * (LMAX / 4);
            unsigned char *q = smem + lay.lpc + (uint32_t)c * lay.lpc_stride;
            const int32_t *row = (const int32_t *)(q + lay.o_qc) + m1 * b.ldL;
            uint32_t w = 0;
            for (int k = 0; k < 4; k++)
                if (4 * t4 + k < b.ldL) w |= ((uint32_t)row[4 * t4 + k] & 0xFFu) << (8 * k);
            ((uint32_t *)(q + lay.o_ldc))[r] = w;
        }
    }
    __syncthreads();
    for (int c = 0; c < p.channels; c++) {
        if (!lay.xs_all) load_planes(b, c);
        bind_lpc_record(b, smem + lay.lpc + (uint32_t)c * lay.lpc_stride, lay, p.L);
        if (b.tid == 0) sh.cur_ch = c;
        __syncthreads();
        analyse_choose<LMAX, FLAVOUR>(b, lay.xs_all ? b.xs + (uint32_t)c * b.xwords : b.xs, c, (const uint32_t *)b.ldc);
        SubDecision *out = p.dec + ((u64)f * (u64)p.channels + (u64)c);
        const uint32_t *src = (const uint32_t *)&sh.dec[c];
        for (uint32_t i = (uint32_t)b.tid; i < sizeof(Decision) / 4; i += (uint32_t)b.B) ((uint32_t *)&out->d)[i] = src[i];
        if (b.tid == 0) out->status = sh.chan_status[c];
    }
}

/* ---- stage B: Levinson-Durbin and coefficient quantisation, one warp per subframe.  The
 * recursion is a chain of dependent fp64 operations: in the one-kernel design a frame's other
 * warps waited for it (a fifth of the frame's time); here every subframe of the launch runs its
 * chain at once and nobody waits. */
#ifndef FLACB_TU_KERNELS_ONLY /* the non-template kernels are defined once, in flacb200.cu */
__global__ void __launch_bounds__(128) lpc_coefficients_kernel(const EncParams p, uint32_t n_sub)
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
    b.ldL = lpc_row_stride(p.L);
    const SmemLayout &lay = p.lay;
    const uint32_t s = blockIdx.x * (uint32_t)b.nwarps + (uint32_t)b.warp;
    pdl_launch_dependents();
    pdl_wait(); /* the autocorrelation of stage A */
    if (s >= n_sub) return;
    uint32_t *rec = (uint32_t *)(p.pre + (u64)s * (u64)p.pre_stride);
    uint32_t *loc = (uint32_t *)(smem + (uint32_t)b.warp * p.pre_stride);
    /* in: autocorrelation and status; out: everything behind the Levinson matrix */
    for (uint32_t i = (uint32_t)b.lane; i < lay.o_levprod / 4; i += 32) loc[i] = rec[i];
    bind_lpc_record(b, (unsigned char *)loc, lay, p.L);
    if (b.lane == 0) *b.lev_stat = *(int32_t *)((unsigned char *)rec + ((unsigned char *)b.lev_stat - (unsigned char *)loc));
    __syncwarp();
    lpc_coefficients_warp(b, p.L, p.P);
    for (uint32_t i = lay.o_qc / 4 + (uint32_t)b.lane; i < lay.o_fixed / 4; i += 32) rec[i] = loc[i];
}
#endif

/* The serialising kernel: one CTA per frame, winner's residual -> Rice partitions -> final bytes
 * at the frame's final place in the stream.
 * CHOSEN = true : stage C2 of the LPC pipeline (analyse_kernel -> lpc_coefficients_kernel ->
 *                 choose_kernel -> this): the subframes' choices are read from HBM.
 * CHOSEN = false: fixed predictors only (max_lpc_order == 0): the fixed-predictor errors and the
 *                 choice are computed here as well; nothing else is launched.
 * TB / MINB: launch bounds.  <576, 1> for frames that fill an SM's shared memory on their own,
 * <288, 3> (72 registers) where three frames fit an SM.  FIXB: the actual block size when it is
 * known at compile time (288, the block-4608 configurations), else 0. */
template <int LMAX, int TB, int MINB, bool CHOSEN, int FIXB>
__global__ void __launch_bounds__(TB, MINB) encode_frames_kernel(const EncParams p)
{
    FLACB_DYN_SMEM(smem);
    __shared__ Shared sh;
    FLACB_SIM_POISON(sh);
    Blk b(p, sh);
    b.tid = (int)threadIdx.x;
    b.B = FIXB ? FIXB : (int)blockDim.x;
    b.lane = b.tid & 31;
    b.warp = b.tid >> 5;
    b.nwarps = b.B >> 5;

    if (b.tid == 0) {
        /* frames are taken in launch order so that every predecessor of a frame is already
         * running (or done) when the frame looks back for its offset */
        sh.frame = p.slot_stride ? blockIdx.x : atomicAdd(p.ticket, 1u);
        sh.status = 0;
        sh.range_flag = 0;
#ifdef FLACB_PHASE_TIMING
        sh.t_last = clock64();
#endif
        sh.o_out = p.out, sh.o_cap = p.out_cap, sh.o_slot_stride = p.slot_stride;
        sh.o_lb_state = p.lb_state, sh.o_frame_off = p.frame_off, sh.o_n_frames = p.n_frames;
        sh.crc = 0;
        sh.pos = 0;
        sh.wbase = 0;
    }
    for (int i = b.tid; i < 256; i += b.B) sh.crc16_tab[i] = c_crc16_tab[i];
    __syncthreads();
    const uint32_t f = sh.frame;
    b.n = (f == p.n_frames - 1) ? p.tail_size : p.block_size;
    b.ntiles = (b.n + b.B * E - 1) / (b.B * E);
    const bool tail_frame = f == p.n_frames - 1 && p.tail_size != p.block_size;
    b.win = tail_frame ? p.win_tail : p.win_full;
    b.flat_lo = tail_frame ? p.flat_tail_lo : p.flat_full_lo;
    b.flat_hi = tail_frame ? p.flat_tail_hi : p.flat_full_hi;

    const SmemLayout &lay = p.lay;
    b.xs = (int32_t *)(smem + lay.xs);
    b.xwords = lay.xwords;
    auto select_channel = [&](int c) { /* per-channel areas: residual, running sums, Rice parameters, LPC front end */
        const uint32_t o = lay.xs_all ? (uint32_t)c * lay.chan_stride : 0u;
        b.res = (uint32_t *)(smem + lay.res) + (lay.xs_all ? (uint32_t)c * lay.xwords : 0u); /* the plane itself */
        b.cum = (u64 *)(smem + lay.cum + o);
        b.kall = (uint8_t *)(smem + lay.kall + o);
        bind_lpc_record(b, smem + lay.lpc + (uint32_t)c * lay.lpc_stride, lay, p.L);
        if (b.tid == 0) sh.cur_ch = c;
    };
    b.stg = (uint32_t *)(smem + lay.stg);
    b.fb = (uint32_t *)(smem + lay.fb);
    b.pst = (uint32_t *)(smem + lay.pst);
    b.red = (u64 *)(smem + lay.red);
    b.ldL = lpc_row_stride(p.L);
    b.stg_words = lay.stg_words;
    b.wx_bytes = lay.kall - lay.res;
    if (b.tid < MAX_CH) sh.chan_status[b.tid] = 0;
    if (b.tid == 0) sh.wmax_hi = 0;

    /* ---- phase 0 + 1 ---- */
    if (lay.xs_all) load_planes(b, -1);
    if constexpr (CHOSEN) {
        /* LPC configurations: the choice of every subframe was made by choose_kernel */
        pdl_wait();
        for (uint32_t i = (uint32_t)b.tid; i < (uint32_t)p.channels * (sizeof(Decision) / 4); i += (uint32_t)b.B) {
            const uint32_t c = i / (uint32_t)(sizeof(Decision) / 4), k = i - c * (uint32_t)(sizeof(Decision) / 4);
            ((uint32_t *)&sh.dec[c])[k] = ((const uint32_t *)&p.dec[(u64)f * (u64)p.channels + c].d)[k];
        }
        __syncthreads(); /* chan_status was zeroed above */
        if (b.tid < p.channels) sh.chan_status[b.tid] = p.dec[(u64)f * (u64)p.channels + (u64)b.tid].status;
        for (int c = 0; c < p.channels; c++) {
            if (!lay.xs_all && (p.channels > 1 || c == 0)) load_planes(b, c);
            select_channel(c);
            __syncthreads();
            analyse_size<LMAX>(b, lay.xs_all ? b.xs + (uint32_t)c * b.xwords : b.xs, c);
        }
    } else {
        /* fixed predictors only (p.L == 0): the whole path in this kernel */
        for (int c = 0; c < p.channels; c++) {
            const int32_t *x = lay.xs_all ? b.xs + (uint32_t)c * b.xwords : b.xs;
            if (!lay.xs_all) load_planes(b, c);
            select_channel(c);
            __syncthreads();
            analyse_front<LMAX, false>(b, x);
            analyse_choose<LMAX, -1>(b, x, c, nullptr);
            analyse_size<LMAX>(b, x, c);
        }
    }
    if (b.tid == 0) { /* the frame's status: first failing channel */
        for (int c = p.channels - 1; c >= 0; c--)
            if (sh.chan_status[c]) sh.status = sh.chan_status[c];
    }

    /* ---- phase 2: frame size -> offset ---- */
    __shared__ uint8_t hdr[16];
    __shared__ int hdr_len;
    if (b.tid == 0) {
        hdr_len = build_frame_header(p, b.n, p.first_frame_number + f, hdr);
        u64 bits = (u64)hdr_len * 8u;
        for (int c = 0; c < p.channels; c++) bits += sh.dec[c].bits;
        sh.payload_bytes = (uint32_t)((bits + 7) >> 3); /* encoder.py:159-160 zero padding */
        sh.frame_bytes = sh.payload_bytes + 2u;        /* + CRC-16 */
        p.frame_bytes[f] = sh.frame_bytes;
    }
    __syncthreads();
    /* The size is published at once; the frame's own offset is only looked up when the first
     * bytes leave shared memory (flush_window -> resolve_offset), by which time the frames in
     * front have long published theirs: nobody waits for a slower predecessor's analysis. */
    if (b.tid == 0) {
        sh.have_off = 0;
        if (!p.slot_stride) atomicExch(&p.lb_state[f], f == 0 ? (LB_PREFIX | (u64)sh.frame_bytes) : (LB_AGG | (u64)sh.frame_bytes));
    }
    for (int idx = b.tid; idx < (int)b.stg_words + 4; idx += b.B) b.stg[idx] = 0;
    __syncthreads();

    /* ---- phase 3: serialise ---- */
    if (b.tid == 0)
        for (int i = 0; i < hdr_len; i++) put_serial(b, hdr[i], 8);
    __syncthreads();
    for (int c = 0; c < p.channels; c++) {
        const Decision &d = sh.dec[c];
        const int32_t *x = lay.xs_all ? b.xs + (uint32_t)c * b.xwords : b.xs;
        select_channel(c);
        if (!lay.xs_all) {
            /* only one channel's residual fits: recompute it and the chosen order's parameters */
            if (!lay.xs_all) load_planes(b, c);
            residual_and_cum<LMAX>(b, x, d, c);
            const int psz = b.n >> d.porder;
            for (int pi = b.tid; pi < (1 << d.porder); pi += b.B) {
                const int start = pi * psz < d.warm ? d.warm : pi * psz;
                const int end = (pi + 1) * psz;
                const uint32_t len = (uint32_t)(end - start);
                const u64 sum = cum_at(b, end) - cum_at(b, start);
                int k = 0;
                if (sum >= (u64)len && len > 0) {
                    k = (63 - __clzll(sum)) - (31 - __clz(len));
                    if (((u64)len << k) > sum) k--;
                }
                b.kall[(1u << d.porder) + (uint32_t)pi] = (uint8_t)k;
            }
        }
        /* make sure the serial part fits the window: header byte, warm-up, LPC parameters */
        if (((sh.pos + 2048u) >> 5) >= sh.wbase + b.stg_words) flush_window(b, ((sh.pos >> 5) - sh.wbase) * 4u, sh.pos >> 5);
        else __syncthreads();
        if (b.tid == 0) {
            const uint32_t mask = p.bps >= 32 ? 0xFFFFFFFFu : ((1u << p.bps) - 1u);
            put_serial(b, d.type ? (0x20u | (uint32_t)(d.order - 1)) << 1 : (0x08u | (uint32_t)d.order) << 1, 8); /* encoder.py:553-569 */
            for (int i = 0; i < d.warm; i++) put_serial(b, (uint32_t)sh.warm[c][i] & mask, p.bps);
            if (d.type) { /* encoder.py:619-625 */
                put_serial(b, (uint32_t)(p.P - 1) & 15u, 4);
                put_serial(b, (uint32_t)d.shift & 31u, 5);
                const uint32_t pm = (1u << p.P) - 1u;
                for (int i = 0; i < d.ncoef; i++) put_serial(b, (uint32_t)d.q[i] & pm, p.P);
            }
        }
        __syncthreads();
        FLACB_MARK(b, 9); /* phase 2, serial header parts, phase-3 recompute */
        pack_residual(b, d);
        FLACB_MARK(b, 10);
    }
    /* encoder.py:159-165: pad to a byte, CRC-16 over the whole frame, append big endian */
    flush_window(b, sh.payload_bytes - sh.wbase * 4u, (sh.payload_bytes + 3u) >> 2);
    if (b.tid == 0) {
        if (sh.writable) {
            uint8_t *dst = p.out + sh.frame_off + sh.payload_bytes;
            dst[0] = (uint8_t)(sh.crc >> 8);
            dst[1] = (uint8_t)sh.crc;
        }
        FLACB_MARK(b, 11); /* final flush, CRC-16 */
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

/* ------------------------------------------------------------------------------------- */
/* slot mode (debugging aid): exclusive scan of frame sizes and compaction                 */

#ifndef FLACB_TU_KERNELS_ONLY
__global__ void scan_frame_bytes_kernel(const uint32_t *frame_bytes, u64 *frame_off, uint32_t n)
{
    /* single CTA, n is small (frames of one call) */
    __shared__ u64 carry;
    __shared__ u64 warp_tot[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const u64 v = i < n ? frame_bytes[i] : 0;
        u64 inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const u64 t = __shfl_up_sync(FULL, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        u64 wb = 0, tot = 0;
        for (int w = 0; w < nw; w++) {
            if (w < warp) wb += warp_tot[w];
            tot += warp_tot[w];
        }
        if (i < n) frame_off[i] = carry + wb + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) frame_off[n] = carry;
}

__global__ void compact_slots_kernel(const uint8_t *slots, u64 slot_stride, const uint32_t *frame_bytes,
                                     const u64 *frame_off, uint8_t *out, u64 out_cap, int32_t *frame_status)
{
    const uint32_t f = blockIdx.x;
    const uint32_t nb = frame_bytes[f];
    const u64 off = frame_off[f];
    if (off + nb > out_cap) {
        if (threadIdx.x == 0 && frame_status[f] == 0) frame_status[f] = FLACB_E_NOSPACE;
        return;
    }
    const uint8_t *src = slots + (u64)f * slot_stride;
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) out[off + i] = src[i];
}
#endif

#endif /* !FLACB_HOST_ONLY */

} /* namespace flacb */
