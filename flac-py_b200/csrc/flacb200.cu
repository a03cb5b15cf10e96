/*
 * flacb200.cu -- libflacb200.so: C ABI (include/flacb200.h) over the sm_100a kernels of
 * flacb200_kernels.cuh.  Host side only does plumbing: validation with the reference's error
 * behaviour, buffers, copies, launches, status collection.  There is no CPU implementation of
 * any stage in here: without a CUDA device every compute entry point fails with FLACB_E_CUDA.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <functional>
#include <string>
#include <vector>

#include "flacb200_host.h"
#include "flacb200_launch.h"
#include "flacb200_serialise.cuh"
#include "flacb200_decode.cuh"

using namespace flacb;

static thread_local std::string g_cuda_error;
static std::atomic<uint64_t> g_launches{0};

/* Clears a call's counters, statuses and look-back words.  A kernel and not cudaMemsetAsync: on the host path the
 * memset queued up on a copy engine behind another chunk's input or output copy (60 us between the end of a chunk's
 * input copy and its first kernel on the call's timeline). */
__global__ void zero_words_kernel(unsigned long long *p, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0ull;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char buf_[512];                                                                    \
            snprintf(buf_, sizeof(buf_), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                     cudaGetErrorString(e_));                                                  \
            g_cuda_error = buf_;                                                               \
            return FLACB_E_CUDA;                                                               \
        }                                                                                      \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return FLACB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        CUDA_TRY(cudaMalloc(&p, want));
        cap = want;
        return FLACB_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return FLACB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        CUDA_TRY(cudaMallocHost(&p, want));
        cap = want;
        return FLACB_OK;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

#ifndef FLACB_SLOTS
#define FLACB_SLOTS 4
#endif
#define FLACB_AHEAD 2 /* chunks enqueued behind the one whose output copy is being started */

/* one encode chunk in flight: its stream, events and buffers */
struct EncSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_c0 = nullptr, ev_c1 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr;
    cudaEvent_t ev_s[3] = {nullptr, nullptr, nullptr}; /* between the stages of the LPC pipeline */
    int staged = 0; /* 0: one kernel (its time under stage_ms[3]); 1: stages timed one by one; 2: stages overlapped, not timed */
    DevBuf pcm, out, slots, frame_bytes, misc, info, pre, dec;
    HostBuf h_meta;
    bool d2h_pending = false;
};

struct flacb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr; /* stream == slot[0].stream */
    /* every bulk copy of the chunked host path goes through one stream per direction: a stream's
     * copies are bound to one copy engine, and with the copies on the chunks' own streams a
     * chunk's input copy queued up behind another chunk's output copy (measured: no overlap) */
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_c0 = nullptr, ev_c1 = nullptr;
    int smem_optin = 0, sm_count = 0;
    size_t static_smem12 = 0, static_smem32 = 0;
    size_t static2_fixed = 0, static2_12 = 0, static2_32 = 0; /* serialise_kernel */
    DevBuf crc4, crc_w; /* slice-by-4 tables and x^(8 i) weights of the word-wise CRC-16 */
    size_t crc_w_n = 0;
    EncSlot slot[FLACB_SLOTS];
    DevBuf win_full, win_tail;
    int win_full_n = -1, win_tail_n = -1;
    bool win_full_ok = true, win_tail_ok = true;
    int flat_full[2] = {0, 0}, flat_tail[2] = {0, 0};
    /* decode */
    DevBuf d_stream, d_pcm, d_cand, d_frames, d_dmisc, d_recs;
    HostBuf h_dmeta;
    cudaEvent_t ev_d0 = nullptr, ev_d1 = nullptr; /* decode: output copy */
    /* a decode whose output did not fit the caller's buffer: flacb_decode_output finishes it */
    struct {
        bool valid = false;
        u64 n_cand = 0, total = 0, nsf = 0, n_frames = 0;
        int fch = 0, status = 0;
        float k_ms = 0.f, h2d_ms = 0.f;
    } pend_dec;
};

extern "C" int flacb_version(void) { return FLACB_VERSION; }

extern "C" int flacb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" const char *flacb_strerror(int status) { return status_text(status); }
extern "C" const char *flacb_last_cuda_error(void) { return g_cuda_error.c_str(); }
extern "C" uint64_t flacb_kernel_launches(void) { return g_launches.load(); }

extern "C" uint64_t flacb_encode_bound(uint64_t n_sample_frames, int channels, int bits, int block_size, int rice_max)
{
    if (block_size < 1) return 0;
    const uint64_t n_frames = (n_sample_frames + (uint64_t)block_size - 1) / (uint64_t)block_size;
    return n_frames * frame_bound(block_size, channels, bits, rice_max) + 16;
}

extern "C" int flacb_stream_prologue(uint32_t sample_rate, int channels, int bits, uint64_t total_samples,
                                     int block_size, uint8_t out[42])
{
    stream_prologue(sample_rate, channels, bits, total_samples, block_size, out);
    return FLACB_OK;
}

static int create_impl(flacb_ctx *c, int device);

extern "C" int flacb_create(int device, flacb_ctx **out)
{
    if (!out) return FLACB_E_ASSERT;
    *out = nullptr;
    int n = 0;
    CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) {
        g_cuda_error = "no such CUDA device";
        return FLACB_E_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    flacb_ctx *c = new flacb_ctx();
    c->device = device;
    const int rc = create_impl(c, device);
    if (rc != FLACB_OK) {
        const std::string why = g_cuda_error; /* flacb_destroy may overwrite it */
        flacb_destroy(c); /* streams, events and buffers made so far */
        g_cuda_error = why;
        return rc;
    }
    *out = c;
    return FLACB_OK;
}

static int create_impl(flacb_ctx *c, int device)
{
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    cudaEvent_t *evs[] = {&c->ev_k0, &c->ev_k1, &c->ev_t0, &c->ev_t1, &c->ev_c0, &c->ev_c1, &c->ev_d0, &c->ev_d1};
    for (cudaEvent_t *e : evs) CUDA_TRY(cudaEventCreate(e));
    for (int i = 0; i < FLACB_SLOTS; i++) {
        EncSlot &sl = c->slot[i];
        if (i == 0) sl.stream = c->stream;
        else CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        cudaEvent_t *se[] = {&sl.ev_k0, &sl.ev_k1, &sl.ev_c0, &sl.ev_c1, &sl.ev_d0, &sl.ev_d1, &sl.ev_s[0], &sl.ev_s[1], &sl.ev_s[2]};
        for (cudaEvent_t *e : se) CUDA_TRY(cudaEventCreate(e));
    }
    CUDA_TRY(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    c->static_smem12 = static_smem_12();
    c->static_smem32 = static_smem_32();
    if (static_smem_fixed() > c->static_smem12) c->static_smem12 = static_smem_fixed();
    /* CRC-16 tables (crc.py:4-14) and the x^n weights of the parallel CRC, once per kernel
     * translation unit */
    uint16_t tab[256], lo[256];
    std::vector<uint16_t> hi(XPOW_HI);
    crc16_table(tab);
    crc16_xpow_tables(lo, hi.data(), XPOW_HI);
    CUDA_TRY(init_fixed(tab, lo, hi.data()));
    CUDA_TRY(init_12(tab, lo, hi.data()));
    CUDA_TRY(init_32(tab, lo, hi.data()));
    c->static2_fixed = static_smem2_fixed();
    c->static2_12 = static_smem2_12();
    c->static2_32 = static_smem2_32();
    {
        uint16_t t4[4][256];
        crc16_slice4(t4);
        int rc4 = c->crc4.ensure(sizeof(t4));
        if (rc4) return rc4;
        CUDA_TRY(cudaMemcpy(c->crc4.p, t4, sizeof(t4), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaDeviceSynchronize());
    return FLACB_OK;
}

extern "C" void flacb_destroy(flacb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&c->win_full, &c->win_tail, &c->crc4, &c->crc_w, &c->d_stream, &c->d_pcm, &c->d_cand, &c->d_frames, &c->d_dmisc, &c->d_recs};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < FLACB_SLOTS; i++) {
        EncSlot &sl = c->slot[i];
        DevBuf *sb[] = {&sl.pcm, &sl.out, &sl.slots, &sl.frame_bytes, &sl.misc, &sl.info, &sl.pre, &sl.dec};
        for (DevBuf *b : sb) b->release();
        sl.h_meta.release();
        cudaEvent_t se[] = {sl.ev_k0, sl.ev_k1, sl.ev_c0, sl.ev_c1, sl.ev_d0, sl.ev_d1, sl.ev_s[0], sl.ev_s[1], sl.ev_s[2]};
        for (cudaEvent_t e : se)
            if (e) cudaEventDestroy(e);
        if (i > 0 && sl.stream) cudaStreamDestroy(sl.stream);
    }
    c->h_dmeta.release();
    cudaEvent_t evs[] = {c->ev_k0, c->ev_k1, c->ev_t0, c->ev_t1, c->ev_c0, c->ev_c1, c->ev_d0, c->ev_d1};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    delete c;
}

extern "C" int flacb_host_alloc(void **p, size_t bytes)
{
    if (!p) return FLACB_E_ASSERT;
    CUDA_TRY(cudaMallocHost(p, bytes ? bytes : 1));
    return FLACB_OK;
}
extern "C" void flacb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
extern "C" int flacb_device_alloc(flacb_ctx *c, void **p, size_t bytes)
{
    if (!c || !p) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMalloc(p, bytes ? bytes : 1));
    return FLACB_OK;
}
extern "C" void flacb_device_free(flacb_ctx *c, void *p)
{
    if (!c || !p) return;
    cudaSetDevice(c->device);
    cudaFree(p);
}
extern "C" int flacb_memcpy_h2d(flacb_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (!c) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FLACB_OK;
}
extern "C" int flacb_memcpy_d2h(flacb_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (!c) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FLACB_OK;
}
extern "C" int flacb_timer_begin(flacb_ctx *c)
{
    if (!c) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->ev_t0, c->stream));
    return FLACB_OK;
}
extern "C" int flacb_timer_end(flacb_ctx *c, float *ms)
{
    if (!c || !ms) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->ev_t1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev_t1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return FLACB_OK;
}
extern "C" int flacb_sync(flacb_ctx *c)
{
    if (!c) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    for (int i = 0; i < FLACB_SLOTS; i++) CUDA_TRY(cudaStreamSynchronize(c->slot[i].stream));
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(c->h2d_stream));
    CUDA_TRY(cudaStreamSynchronize(c->d2h_stream));
    return FLACB_OK;
}

/* ------------------------------------------------------------------------------------- */

static int upload_window(flacb_ctx *c, int n, DevBuf &buf, int &cached_n, bool &ok, int &flat_lo, int &flat_hi)
{
    if (cached_n == n) return FLACB_OK;
    std::vector<double> w;
    ok = tukey_window(n, w, 64);
    window_flat_range(w, n, flat_lo, flat_hi);
    int rc = buf.ensure(w.size() * sizeof(double));
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(buf.p, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream)); /* w is a stack-lifetime vector */
    cached_n = n;
    return FLACB_OK;
}

/* One launch over frames [0, n_frames) of device-resident PCM into device-resident output, on
 * one slot (stream + buffers).  On return the kernels are enqueued on the slot's stream and the
 * copy of the meta data (frame offsets, statuses, counters) to pinned host memory behind them. */
struct ChunkMeta {
    u64 *h_frame_off;  /* n_frames + 1 */
    int32_t *h_status; /* n_frames */
    u64 *h_stats;      /* 8 */
    int threads, smem;
};

struct EncPlan {
    LaunchShape shape;  /* the serialising kernel (encode_frames_kernel) */
    bool lmax12;
    size_t static_smem;
    /* LPC configurations: analyse_kernel -> lpc_coefficients_kernel -> choose_kernel ->
     * encode_frames_kernel<.., true>; fixed predictors only: encode_frames_kernel<.., false> */
    bool staged;
    LaunchShape front;  /* shape of analyse_kernel */
    LaunchShape choose; /* shape of choose_kernel */
    /* second-generation serialiser (flacb200_serialise.cuh) when the configuration is inside its limits */
    bool ser2;
    int ser_variant, ser_threads, ser_ctas;
    SerLayout slay;
};

/* serialise_kernel: pick samples per thread / launch bounds (ser_variant) for the most resident
 * threads per SM.  FLACB200_SERIALISE=1: the first-generation kernel; =2,<variant>[,<frames per SM>]:
 * that variant (tuning / tests). */
static void plan_serialise2(flacb_ctx *c, const EncodeArgs &a, EncPlan &plan)
{
    plan.ser2 = false;
    int forced = -1, forced_ctas = 0;
    if (const char *e = getenv("FLACB200_SERIALISE")) {
        if (atoi(e) == 1) return;
        sscanf(e, "2,%d,%d", &forced, &forced_ctas);
    }
    if ((a.flags & FLACB_SLOT_COMPACT) || a.rice_max > V2_MAX_RICE) return;
    const size_t st = a.max_lpc_order == 0 ? c->static2_fixed : a.max_lpc_order <= 12 ? c->static2_12 : c->static2_32;
    int best_score = 0;
    for (int v = 0; v < SER_VARIANTS; v++) {
        if (forced >= 0 && v != forced) continue;
        const SerVariant sv = ser_variant(v);
        const int S = sv.S;
        int B = ((a.block_size + S - 1) / S + 31) / 32 * 32;
        if (B > sv.max_threads) continue;
        int ctas = 65536 / (sv.regs * B);
        if (ctas > sv.ctas && sv.ctas > 1) ctas = sv.ctas; /* measured: more frames per SM than three did not pay */
        if (ctas > 8) ctas = 8;
        if (forced_ctas > 0 && ctas > forced_ctas) ctas = forced_ctas;
        for (; ctas >= 1; ctas--) {
            size_t per_cta = (size_t)(228 * 1024) / (size_t)ctas;
            if (per_cta > (size_t)c->smem_optin + 1024) per_cta = (size_t)c->smem_optin + 1024;
            if (per_cta <= st + 1024 + 4096) continue;
            const SerLayout lay = make_ser_layout(a.block_size, a.channels, a.bits, a.rice_max, B, S, (uint32_t)(per_cta - st - 1024));
            if (lay.stg_words == 0) continue;
            const int score = ctas * B;
            if (score > best_score) {
                best_score = score;
                plan.ser2 = true;
                plan.ser_variant = v, plan.ser_threads = B, plan.ser_ctas = ctas, plan.slay = lay;
            }
            break;
        }
    }
}

/* "threads,all_planes,variant" tuning knobs */
static void shape_from_env(const char *name, LaunchShape &sh, size_t static_smem, size_t smem_optin,
                           const std::function<size_t(int, int)> &layout_total)
{
    const char *e = getenv(name);
    if (!e) return;
    int t = 0, ap = 1, v = 0;
    if (sscanf(e, "%d,%d,%d", &t, &ap, &v) == 3 && t >= 32 && t % 32 == 0 && t <= (v ? 288 : 576) &&
        layout_total(t, ap) + static_smem <= smem_optin)
        sh.threads = t, sh.all_planes = ap, sh.variant = v;
}

static int plan_encode(flacb_ctx *c, const EncodeArgs &a, EncPlan &plan)
{
    plan.lmax12 = a.max_lpc_order <= 12;
    plan.static_smem = plan.lmax12 ? c->static_smem12 : c->static_smem32;
    if ((size_t)c->smem_optin <= plan.static_smem + 1024) return FLACB_E_UNSUPPORTED;
    if (frame_bound(a.block_size, a.channels, a.bits, a.rice_max) >= (u64)XPOW_HI * 256) return FLACB_E_UNSUPPORTED;
    /* the serialising kernel carries no LPC front end in shared memory */
    auto layout_total = [&](int threads, int all_planes) {
        return (size_t)make_layout(a.block_size, a.channels, a.bits, a.rice_max, 0, threads, all_planes).total;
    };
    plan.shape = pick_shape(a.block_size, E, plan.static_smem, (size_t)c->smem_optin, 228 * 1024, layout_total);
    if (plan.shape.threads == 0) return FLACB_E_UNSUPPORTED; /* the block does not fit shared memory */
    shape_from_env("FLACB200_SHAPE", plan.shape, plan.static_smem, (size_t)c->smem_optin, layout_total);
    plan.staged = a.max_lpc_order > 0;
    plan_serialise2(c, a, plan);
    if (plan.staged) {
        /* stages A and C1: as many resident frames as the registers allow (three at 72), every
         * plane resident when that does not cost a frame */
        auto small_shape = [&](const std::function<size_t(int, int)> &total, LaunchShape &out, bool wide_ok) {
            /* blocks of more than four 288-thread tiles (block 16384 of config 3): the analysis ran
             * 10 % faster with one 512-thread frame per SM than with 256-thread frames (measured) */
            const bool wide = wide_ok && a.block_size > 4 * 288 * E;
            LaunchShape f = {pick_threads(a.block_size, E, wide ? 576 : 288), 0, 0, wide ? 0 : 1};
            for (int ap = 1; ap >= 0; ap--) {
                const size_t dyn = total(f.threads, ap);
                if (dyn + plan.static_smem > (size_t)c->smem_optin) continue;
                int ctas = (int)((228 * 1024) / (dyn + plan.static_smem + 1024));
                const int by_regs = 65536 / ((wide ? 96 : 72) * f.threads);
                if (ctas > by_regs) ctas = by_regs;
                if (ctas > f.ctas_per_sm) f.all_planes = ap, f.ctas_per_sm = ctas;
            }
            out = f;
            return f.ctas_per_sm > 0;
        };
        auto front_total = [&](int threads, int all_planes) {
            return (size_t)make_layout_front(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, threads, all_planes).total;
        };
        auto choose_total = [&](int threads, int all_planes) {
            return (size_t)make_layout_choose(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, threads, all_planes).total;
        };
        if (!small_shape(front_total, plan.front, true) || !small_shape(choose_total, plan.choose, false)) return FLACB_E_UNSUPPORTED;
        /* the all-orders filter needs no more than 56 registers: four 288-thread frames per SM
         * when their planes and coefficient records fit */
        if (plan.choose.variant == 1 && plan.choose.threads == 288 &&
            4 * (choose_total(288, plan.choose.all_planes) + plan.static_smem + 1024) <= 228 * 1024)
            plan.choose.variant = 2, plan.choose.ctas_per_sm = 4;
        shape_from_env("FLACB200_SHAPE_A", plan.front, plan.static_smem, (size_t)c->smem_optin, front_total);
        shape_from_env("FLACB200_SHAPE_C1", plan.choose, plan.static_smem, (size_t)c->smem_optin, choose_total);
    }
    return FLACB_OK;
}

static int launch_encode(flacb_ctx *c, EncSlot &sl, const EncPlan &plan, const uint8_t *d_pcm, const EncodeArgs &a,
                         int full_tail, uint8_t *d_out, u64 out_cap, flacb_subframe_info *d_info, ChunkMeta &meta)
{
    /* full_tail: the stream's last frame is not in this chunk, every frame here is a full block */
    const u64 n_frames = (a.n_sample_frames + (u64)a.block_size - 1) / (u64)a.block_size;
    const int tail = full_tail ? a.block_size : (int)(a.n_sample_frames - (n_frames - 1) * (u64)a.block_size);
    const LaunchShape &shape = plan.shape;
    const int B = shape.threads;
    const SmemLayout lay = make_layout(a.block_size, a.channels, a.bits, a.rice_max, 0, B, shape.all_planes);
    const SmemLayout lay_rec = make_layout(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, B, 0); /* record offsets */

    int rc; /* the window tables were uploaded by the caller before the first chunk */
    /* One device block for everything a call clears and reads back, so that a call makes one memset and
     * one copy to the host instead of two and three (per-call driver work is what separates the step time
     * from the sum of the stages, more so with eight ranks on one host):
     *   [frame_off u64 n+1][stats + ticket 512 B][status i32 n] | [look-back words u64 n]   <- cleared
     *   `-------------------- copied to the host ------------'                                         */
    const size_t o_foff = 0, o_misc = (n_frames + 1) * 8, o_status = o_misc + 512;
    const size_t meta_bytes = (o_status + n_frames * 4 + 15) & ~(size_t)15, o_lb = meta_bytes;
    if ((rc = sl.frame_bytes.ensure(n_frames * 4))) return rc;
    if ((rc = sl.misc.ensure(o_lb + n_frames * 8))) return rc;
    if ((rc = sl.h_meta.ensure(meta_bytes + 16))) return rc;
    uint8_t *dmeta = (uint8_t *)sl.misc.p;

    EncParams p;
    memset(&p, 0, sizeof(p));
    p.pcm = d_pcm;
    p.n_sample_frames = a.n_sample_frames;
    p.channels = a.channels;
    p.bps = a.bits;
    p.cbytes = (a.flags & FLACB_PCM_INT32) ? 4 : (a.bits + 7) / 8;
    p.block_size = a.block_size;
    p.L = a.max_lpc_order;
    p.P = a.qlp_precision;
    p.rmin = a.rice_min;
    p.rmax = a.rice_max;
    p.first_frame_number = a.first_frame_number;
    p.flags = a.flags;
    p.n_frames = (uint32_t)n_frames;
    p.tail_size = tail;
    const bool own_tail = tail != a.block_size; /* only the stream's last frame can be short */
    p.win_full = (const double *)c->win_full.p;
    p.win_tail = own_tail ? (const double *)c->win_tail.p : p.win_full;
    p.win_full_bad = !c->win_full_ok;
    p.win_tail_bad = own_tail ? !c->win_tail_ok : p.win_full_bad;
    p.flat_full_lo = c->flat_full[0], p.flat_full_hi = c->flat_full[1];
    p.flat_tail_lo = own_tail ? c->flat_tail[0] : p.flat_full_lo;
    p.flat_tail_hi = own_tail ? c->flat_tail[1] : p.flat_full_hi;
    p.frame_bytes = (uint32_t *)sl.frame_bytes.p;
    p.frame_off = (u64 *)(dmeta + o_foff);
    p.frame_status = (int32_t *)(dmeta + o_status);
    p.lb_state = (u64 *)(dmeta + o_lb);
    p.ticket = (uint32_t *)(dmeta + o_misc);
    p.stats = (u64 *)(dmeta + o_misc + 64);
    p.info = d_info;
    p.all_planes = shape.all_planes;
    p.lay = lay;
    if (a.flags & FLACB_SLOT_COMPACT) {
        p.slot_stride = (frame_bound(a.block_size, a.channels, a.bits, a.rice_max) + 15) & ~15ull;
        if ((rc = sl.slots.ensure(p.slot_stride * n_frames))) return rc;
        p.out = (uint8_t *)sl.slots.p;
        p.out_cap = p.slot_stride * n_frames;
    } else {
        p.slot_stride = 0;
        p.out = d_out;
        p.out_cap = out_cap;
    }

    const u64 n_sub = n_frames * (u64)a.channels;
    if (plan.staged) {
        p.pre_stride = (lay_rec.lpc_stride + 15u) & ~15u;
        if ((rc = sl.pre.ensure(n_sub * (u64)p.pre_stride))) return rc;
        p.pre = (uint8_t *)sl.pre.p;
        if ((rc = sl.dec.ensure(n_sub * sizeof(SubDecision)))) return rc;
        p.dec = (SubDecision *)sl.dec.p;
    }
    /* how choose_kernel multiplies: sample pairs against coefficient bytes (16-bit PCM, at most
     * 8-bit coefficients), 32-bit sums proven exact for every order, or decided per order */
    /* ... or, for samples wider than 16 bits with coefficients of at most 8 bits, as byte planes on the
     * tensor cores (flavour 3, flacb200_choose_mma.cuh) */
    int flavour = 2;
    if (a.bits <= 16 && a.qlp_precision <= 8) flavour = 0;
    else if (a.qlp_precision <= 8 && a.bits <= 24 && a.max_lpc_order > 16) flavour = 3; /* measured: below that the scalar filter is faster */
    else if (((u64)a.max_lpc_order << (a.qlp_precision - 1 + a.bits - 1)) <= (1ull << 30)) flavour = 1;
    if (const char *e = getenv("FLACB200_FLAVOUR")) { /* tuning / test knob: another, equally exact flavour */
        const int v = atoi(e);
        if (v == 3 && a.qlp_precision <= 8 && a.bits <= 24) flavour = 3;
        else if (v == 2 || (v == 1 && ((u64)a.max_lpc_order << (a.qlp_precision - 1 + a.bits - 1)) <= (1ull << 30)) ||
                 (v == 0 && a.bits <= 16 && a.qlp_precision <= 8))
            flavour = v;
    }
    const MmaLayout mlay = make_layout_choose_mma(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, 256);
    if (flavour == 3 && (size_t)mlay.total + 16384 > (size_t)c->smem_optin) flavour = 2; /* the planes of a block do not fit */
    /* stage events in between (FLACB_STAGE_TIMES) keep the stages apart; without them the stages behind the
     * first are launched with programmatic stream serialisation (FLACB200_NO_PDL=1: plain launches) */
    const bool stage_times = (a.flags & FLACB_STAGE_TIMES) != 0;
    const bool pdl = !stage_times && plan.staged && !getenv("FLACB200_NO_PDL");
    auto launch = [&](int kind, int variant, int threads, size_t smem, const EncParams &kp) {
        g_launches++;
        const unsigned grid = (unsigned)n_frames;
        const int fl = flavour | ((pdl && kind != K_ANALYSE) ? 0x100 : 0);
        if (kind == K_FIXED || kind == K_FIXED2) return launch_fixed(kind, variant, fl, grid, threads, smem, sl.stream, kp);
        return plan.lmax12 ? launch_12(kind, variant, fl, grid, threads, smem, sl.stream, kp)
                           : launch_32(kind, variant, fl, grid, threads, smem, sl.stream, kp);
    };

    /* FLACB200_DEBUG_SYNC=1 (development aid): wait behind every stage, so that a fault names its kernel */
    const bool dbg_sync = getenv("FLACB200_DEBUG_SYNC") != nullptr;
    auto stage_done = [&](const char *what) -> int {
        if (!dbg_sync) return FLACB_OK;
        cudaError_t e = cudaStreamSynchronize(sl.stream);
        if (e != cudaSuccess) {
            char buf[256];
            snprintf(buf, sizeof(buf), "stage %s: %s", what, cudaGetErrorString(e));
            g_cuda_error = buf;
            return FLACB_E_CUDA;
        }
        return FLACB_OK;
    };
    {   /* counters, statuses, look-back words (o_misc and the length are multiples of 8) */
        const uint64_t words = (o_lb + n_frames * 8 - o_misc) / 8;
        if (getenv("FLACB200_MEMSET")) CUDA_TRY(cudaMemsetAsync(dmeta + o_misc, 0, words * 8, sl.stream));
        else zero_words_kernel<<<(unsigned)((words + 255) / 256), 256, 0, sl.stream>>>((unsigned long long *)(dmeta + o_misc), words);
    }
    CUDA_TRY(cudaEventRecord(sl.ev_k0, sl.stream));
    sl.staged = !plan.staged ? 0 : stage_times ? 1 : 2;
    if (plan.staged) {
        EncParams pa = p;
        pa.lay = make_layout_front(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, plan.front.threads,
                                   plan.front.all_planes);
        pa.all_planes = plan.front.all_planes;
        CUDA_TRY(launch(K_ANALYSE, plan.front.variant, plan.front.threads, pa.lay.total, pa));
        if ((rc = stage_done("A (analyse_kernel)"))) return rc;
        if (stage_times) CUDA_TRY(cudaEventRecord(sl.ev_s[0], sl.stream));
        const int wpb = 4; /* subframes (warps) per CTA of stage B */
        const size_t smem_b = (size_t)wpb * p.pre_stride;
        EncParams pb = p;
        pb.lay = lay_rec;
        CUDA_TRY(cudaFuncSetAttribute(lpc_coefficients_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        if (pdl) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)((n_sub + wpb - 1) / wpb), 1, 1);
            cfg.blockDim = dim3(wpb * 32, 1, 1);
            cfg.dynamicSmemBytes = smem_b;
            cfg.stream = sl.stream;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr.val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, lpc_coefficients_kernel, pb, (uint32_t)n_sub));
        } else {
            lpc_coefficients_kernel<<<(unsigned)((n_sub + wpb - 1) / wpb), wpb * 32, smem_b, sl.stream>>>(pb, (uint32_t)n_sub);
            CUDA_TRY(cudaGetLastError());
        }
        g_launches++;
        if ((rc = stage_done("B (lpc_coefficients_kernel)"))) return rc;
        if (stage_times) CUDA_TRY(cudaEventRecord(sl.ev_s[1], sl.stream));
        EncParams pc = p;
        pc.lay = make_layout_choose(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, plan.choose.threads,
                                    plan.choose.all_planes);
        pc.all_planes = plan.choose.all_planes;
        if (flavour == 3) {
            EncParams pm = p;
            pm.lay = lay_rec;
            g_launches++;
            CUDA_TRY(plan.lmax12 ? launch_mma_12(pdl, (unsigned)n_frames, 256, mlay.total, sl.stream, pm, mlay)
                                 : launch_mma_32(pdl, (unsigned)n_frames, 256, mlay.total, sl.stream, pm, mlay));
        } else
            CUDA_TRY(launch(K_CHOOSE, plan.choose.variant, plan.choose.threads, pc.lay.total, pc));
        if ((rc = stage_done(flavour == 3 ? "C1 (choose_mma_kernel)" : "C1 (choose_kernel)"))) return rc;
        if (stage_times) CUDA_TRY(cudaEventRecord(sl.ev_s[2], sl.stream));
    }
    if (plan.ser2) {
        p.slay = plan.slay;
        p.crc4 = (const uint16_t *)c->crc4.p;
        p.crc_w = (const uint16_t *)c->crc_w.p;
        CUDA_TRY(launch(plan.staged ? K_SERIALISE2 : K_FIXED2, plan.ser_variant, plan.ser_threads, plan.slay.total, p));
    } else {
        CUDA_TRY(launch(plan.staged ? K_SERIALISE : K_FIXED, shape.variant, B, lay.total, p));
    }
    if (a.flags & FLACB_SLOT_COMPACT) {
        scan_frame_bytes_kernel<<<1, 256, 0, sl.stream>>>(p.frame_bytes, p.frame_off, p.n_frames);
        CUDA_TRY(cudaGetLastError());
        compact_slots_kernel<<<(unsigned)n_frames, 256, 0, sl.stream>>>((const uint8_t *)sl.slots.p, p.slot_stride, p.frame_bytes,
                                                                        p.frame_off, d_out, out_cap, p.frame_status);
        CUDA_TRY(cudaGetLastError());
        g_launches += 2;
    }
    if ((rc = stage_done("C2 (serialising kernel)"))) return rc;
    CUDA_TRY(cudaEventRecord(sl.ev_k1, sl.stream));

    uint8_t *h = (uint8_t *)sl.h_meta.p;
    meta.h_frame_off = (u64 *)(h + o_foff);
    meta.h_stats = (u64 *)(h + o_misc + 64);
    meta.h_status = (int32_t *)(h + o_status);
    meta.threads = plan.ser2 ? plan.ser_threads : B;
    meta.smem = plan.ser2 ? (int)plan.slay.total : (int)lay.total;
    CUDA_TRY(cudaMemcpyAsync(h, dmeta, meta_bytes, cudaMemcpyDeviceToHost, sl.stream));
    return FLACB_OK;
}

/* first frame of the chunk, in stream order, on which the reference raises (strict mode) */
static int first_bad(const ChunkMeta &meta, u64 n_frames, uint32_t flags, int64_t &bad, u64 &valid)
{
    valid = meta.h_frame_off[n_frames];
    bad = -1;
    for (u64 f = 0; f < n_frames; f++) {
        const int s = meta.h_status[f];
        const bool ood = s >= FLACB_E_ASSERT && s <= FLACB_E_ZERODIV;
        if (s != 0 && !(ood && (flags & FLACB_LENIENT))) {
            bad = (int64_t)f;
            valid = meta.h_frame_off[f];
            return s;
        }
    }
    return FLACB_OK;
}

static int encode_pcm_impl(flacb_ctx *c, const void *pcm, uint64_t n_sample_frames, int channels, int bits,
                           int block_size, int max_lpc_order, int qlp_precision, int rice_min, int rice_max,
                           uint64_t first_frame_number, uint32_t flags, uint8_t *out, uint64_t out_cap,
                           uint64_t *out_len, uint64_t *frame_offsets, flacb_subframe_info *info,
                           flacb_stats *stats);

extern "C" int flacb_encode_pcm(flacb_ctx *c, const void *pcm, uint64_t n_sample_frames, int channels, int bits,
                                int block_size, int max_lpc_order, int qlp_precision, int rice_min, int rice_max,
                                uint64_t first_frame_number, uint32_t flags, uint8_t *out, uint64_t out_cap,
                                uint64_t *out_len, uint64_t *frame_offsets, flacb_subframe_info *info,
                                flacb_stats *stats)
{
    const int rc = encode_pcm_impl(c, pcm, n_sample_frames, channels, bits, block_size, max_lpc_order, qlp_precision, rice_min,
                                   rice_max, first_frame_number, flags, out, out_cap, out_len, frame_offsets, info, stats);
    if (c && rc != FLACB_OK) {
        /* An early exit (a failed allocation, launch or copy, or the first frame the reference raises on)
         * can leave kernels and copies of other chunks in flight that still write into the caller's
         * buffers: wait for all of them before the buffers go back to the caller, and forget them. */
        cudaSetDevice(c->device);
        for (int i = 0; i < FLACB_SLOTS; i++) {
            cudaStreamSynchronize(c->slot[i].stream);
            c->slot[i].d2h_pending = false;
        }
        cudaStreamSynchronize(c->h2d_stream);
        cudaStreamSynchronize(c->d2h_stream);
        if (rc == FLACB_E_CUDA) cudaGetLastError(); /* the text is in flacb_last_cuda_error() */
    }
    return rc;
}

static int encode_pcm_impl(flacb_ctx *c, const void *pcm, uint64_t n_sample_frames, int channels, int bits,
                           int block_size, int max_lpc_order, int qlp_precision, int rice_min, int rice_max,
                           uint64_t first_frame_number, uint32_t flags, uint8_t *out, uint64_t out_cap,
                           uint64_t *out_len, uint64_t *frame_offsets, flacb_subframe_info *info,
                           flacb_stats *stats)
{
    if (out_len) *out_len = 0;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->first_bad_frame = -1;
    }
    if (!c) return FLACB_E_ASSERT;
    EncodeArgs a = {n_sample_frames, channels, bits, block_size, max_lpc_order, qlp_precision,
                    rice_min, rice_max, first_frame_number, flags};
    int rc = validate_encode_args(a);
    if (rc) return rc;
    const u64 n_frames = (n_sample_frames + (u64)block_size - 1) / (u64)block_size;
    if (n_frames == 0) {
        if (frame_offsets) frame_offsets[0] = 0;
        return FLACB_OK;
    }
    if ((!pcm) || (!out && out_cap)) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    const int cbytes = (flags & FLACB_PCM_INT32) ? 4 : (bits + 7) / 8;
    const u64 frame_pcm = (u64)block_size * (u64)channels * (u64)cbytes;
    const u64 pcm_bytes = n_sample_frames * (u64)channels * (u64)cbytes;
    EncPlan plan;
    if ((rc = plan_encode(c, a, plan))) return rc;
    {   /* Tukey windows of the full block and of the stream's last, shorter block */
        const int real_tail = (int)(n_sample_frames - (n_frames - 1) * (u64)block_size);
        if ((rc = upload_window(c, block_size, c->win_full, c->win_full_n, c->win_full_ok, c->flat_full[0], c->flat_full[1]))) return rc;
        if (real_tail != block_size &&
            (rc = upload_window(c, real_tail, c->win_tail, c->win_tail_n, c->win_tail_ok, c->flat_tail[0], c->flat_tail[1])))
            return rc;
    }
    if (plan.ser2) { /* weights of the word-wise CRC-16: x^(8 i) for every distance to the end of a frame */
        const size_t need = (size_t)frame_bound(block_size, channels, bits, rice_max) + 64;
        if (c->crc_w_n < need) {
            std::vector<uint16_t> w(need + need / 2);
            crc16_weights(w.data(), w.size());
            if ((rc = flacb_sync(c))) return rc; /* nothing in flight may still read the old table */
            if ((rc = c->crc_w.ensure(w.size() * 2))) return rc;
            CUDA_TRY(cudaMemcpy(c->crc_w.p, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
            c->crc_w_n = w.size();
        }
    }
    const bool host_in = !(flags & FLACB_INPUT_ON_DEVICE), host_out = !(flags & FLACB_OUTPUT_ON_DEVICE);

    /* Host buffers on both sides: cut the frame range into chunks and keep several in flight, so
     * that the H2D copy of one chunk, the kernels of another and the D2H copy of a third overlap.
     * Chunks are whole waves of resident CTAs.  Everything else: one chunk. */
    std::vector<uint64_t> chunk_f0; /* first frame of every chunk, then n_frames */
    if (host_in && host_out && !info && !(flags & FLACB_SLOT_COMPACT)) {
        const char *e_want = getenv("FLACB200_CHUNK_FRAMES"), *e_edge = getenv("FLACB200_EDGE_CHUNK_FRAMES"); /* tuning / test knobs */
        const u64 wave = (u64)c->sm_count * (u64)(plan.ser2 ? plan.ser_ctas : plan.shape.ctas_per_sm);
        long long e_first = 0, e_last = 0; /* "first[,last]" */
        if (e_edge) sscanf(e_edge, "%lld,%lld", &e_first, &e_last);
        else if (plan.staged) {
            /* LPC configurations: a third of a wave first (the kernels start sooner behind a shorter copy) and
             * two thirds last (four stages of kernels and the output copy run with nothing to overlap):
             * config 2 end to end 2.73 -> 2.57 ms.  Fixed predictors only: one wave at both ends measured best. */
            e_first = (long long)(wave / 3 ? wave / 3 : 1);
            e_last = (long long)(2 * wave / 3 ? 2 * wave / 3 : 1);
        }
        chunk_f0 = chunk_plan(n_frames, wave, frame_pcm, e_want ? atoll(e_want) : 0, e_first, e_last);
    } else {
        chunk_f0.push_back(0);
        chunk_f0.push_back(n_frames);
    }
    const u64 n_chunks = chunk_f0.size() - 1;
    float h2d_ms = 0.f, d2h_ms = 0.f, k_ms = 0.f, stage_ms[4] = {0.f, 0.f, 0.f, 0.f};
    u64 host_off = 0; /* bytes produced so far */
    int64_t bad_frame = -1;
    u64 ood = 0, hist[6] = {0, 0, 0, 0, 0, 0};
    int threads = 0, smem = 0;
    int status = FLACB_OK;

    const bool timeline = getenv("FLACB200_TIMELINE") != nullptr; /* development aid: when did each chunk's copies and kernels run */
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    struct Pending {
        ChunkMeta meta;
        u64 f0, nfr;
        bool active;
    } pend[FLACB_SLOTS];
    for (int i = 0; i < FLACB_SLOTS; i++) pend[i].active = false;

    /* retire chunk in slot s: wait for its kernel, read sizes and statuses, start its D2H */
    auto retire = [&](int s) -> int {
        EncSlot &sl = c->slot[s];
        Pending &pd = pend[s];
        if (!pd.active) return FLACB_OK;
        pd.active = false;
        CUDA_TRY(cudaStreamSynchronize(sl.stream));
        if (timeline) fprintf(stderr, "host: chunk f0=%llu kernels seen done at %.3f\n", (unsigned long long)pd.f0, host_ms());
        /* first things first: the chunk's size is on the host now, so its output copy can start; the
         * bookkeeping (a dozen event queries, the offsets) follows while it runs -- for the last chunk of a call
         * nothing else hides it */
        const bool discard = status != FLACB_OK; /* an earlier chunk already failed */
        int64_t bad = -1;
        u64 valid = 0;
        int st = FLACB_OK;
        bool nospace = false;
        if (!discard) {
            st = first_bad(pd.meta, pd.nfr, flags, bad, valid);
            if (host_out) {
                if (host_off + valid > out_cap) nospace = true;
                else if (valid) {
                    const double h0 = host_ms();
                    /* the chunk's kernels are done (the host has just waited for them) */
                    CUDA_TRY(cudaEventRecord(sl.ev_d0, c->d2h_stream));
                    CUDA_TRY(cudaMemcpyAsync(out + host_off, sl.out.p, valid, cudaMemcpyDeviceToHost, c->d2h_stream));
                    CUDA_TRY(cudaEventRecord(sl.ev_d1, c->d2h_stream));
                    if (timeline) fprintf(stderr, "host: d2h of chunk f0=%llu enqueued %.3f-%.3f\n", (unsigned long long)pd.f0, h0, host_ms());
                    sl.d2h_pending = true;
                }
            }
        }
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, sl.ev_k0, sl.ev_k1));
        k_ms += ms;
        if (timeline) {
            float a0 = 0, a1 = 0, b0 = 0, b1 = 0;
            if (host_in) cudaEventElapsedTime(&a0, c->ev_c0, sl.ev_c0), cudaEventElapsedTime(&a1, c->ev_c0, sl.ev_c1);
            cudaEventElapsedTime(&b0, c->ev_c0, sl.ev_k0), cudaEventElapsedTime(&b1, c->ev_c0, sl.ev_k1);
            fprintf(stderr, "chunk f0=%llu slot %d: h2d %.3f-%.3f kernels %.3f-%.3f\n", (unsigned long long)pd.f0, s, a0, a1, b0, b1);
        }
        if (sl.staged == 1) {
            cudaEvent_t ev[5] = {sl.ev_k0, sl.ev_s[0], sl.ev_s[1], sl.ev_s[2], sl.ev_k1};
            for (int k = 0; k < 4; k++) {
                CUDA_TRY(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
                stage_ms[k] += ms;
            }
        } else if (sl.staged == 0) {
            stage_ms[3] += ms;
        }
        if (host_in) {
            CUDA_TRY(cudaEventElapsedTime(&ms, sl.ev_c0, sl.ev_c1));
            h2d_ms += ms;
        }
        if (discard) return FLACB_OK;
        if (frame_offsets)
            for (u64 f = 0; f <= pd.nfr; f++) frame_offsets[pd.f0 + f] = host_off + pd.meta.h_frame_off[f];
        for (int o = 0; o < 6; o++) hist[o] += pd.meta.h_stats[o];
#ifdef FLACB_PHASE_TIMING
        {
            u64 tot = 0;
            for (int k = 0; k < 11; k++) tot += pd.meta.h_stats[8 + k];
            fprintf(stderr, "phase cycles/frame:");
            for (int k = 0; k < 11; k++) fprintf(stderr, " %d:%.0f", k, (double)pd.meta.h_stats[8 + k] / (double)pd.nfr);
            fprintf(stderr, "  (total %.0f; look-back retries/frame %.2f)\n", (double)tot / (double)pd.nfr,
                    (double)pd.meta.h_stats[8 + 11] / (double)pd.nfr);
        }
#endif
        ood += pd.meta.h_stats[6];
        threads = pd.meta.threads, smem = pd.meta.smem;
        if (nospace) {
            status = st != FLACB_OK ? st : FLACB_E_NOSPACE;
            bad_frame = (int64_t)pd.f0 + (bad >= 0 ? bad : 0);
            return FLACB_OK;
        }
        host_off += valid;
        if (st != FLACB_OK) {
            status = st;
            bad_frame = (int64_t)pd.f0 + bad;
        }
        return FLACB_OK;
    };
    auto finish_d2h = [&](int s) -> int {
        EncSlot &sl = c->slot[s];
        if (!sl.d2h_pending) return FLACB_OK;
        sl.d2h_pending = false;
        CUDA_TRY(cudaEventSynchronize(sl.ev_d1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, sl.ev_d0, sl.ev_d1));
        d2h_ms += ms;
        if (timeline) {
            float d0 = 0, d1 = 0;
            cudaEventElapsedTime(&d0, c->ev_c0, sl.ev_d0), cudaEventElapsedTime(&d1, c->ev_c0, sl.ev_d1);
            fprintf(stderr, "slot %d: d2h %.3f-%.3f\n", s, d0, d1);
        }
        return FLACB_OK;
    };

    flacb_subframe_info *d_info = nullptr;
    if (info) {
        if ((rc = c->slot[0].info.ensure(n_frames * (u64)channels * sizeof(flacb_subframe_info)))) return rc;
        d_info = (flacb_subframe_info *)c->slot[0].info.p;
    }
    if (timeline) CUDA_TRY(cudaEventRecord(c->ev_c0, c->slot[0].stream));
    int ahead = FLACB_AHEAD;
    if (const char *e = getenv("FLACB200_AHEAD")) { /* tuning knob */
        const int v = atoi(e);
        if (v >= 1 && v < FLACB_SLOTS) ahead = v;
    }
    /* The input copies are what the call waits for (the link is slower than the kernels), so the copy of chunk
     * ci + 1 is queued as soon as its slot's input buffer is free -- before the host goes to wait for an older
     * chunk's kernels -- and not when the loop gets to it: the copy engine never runs dry. */
    std::vector<char> h2d_issued(n_chunks, 0);
    auto chunk_span = [&](u64 cj, u64 &f0, u64 &nfr, u64 &s0, u64 &ns) {
        f0 = chunk_f0[cj];
        nfr = chunk_f0[cj + 1] - f0;
        s0 = f0 * (u64)block_size;
        ns = (f0 + nfr == n_frames) ? n_sample_frames - s0 : nfr * (u64)block_size;
    };
    auto issue_h2d = [&](u64 cj) -> int {
        if (!host_in || cj >= n_chunks || h2d_issued[cj]) return FLACB_OK;
        const int sj = (int)(cj % FLACB_SLOTS);
        if (pend[sj].active) return FLACB_OK; /* the kernels of the slot's chunk may still read its input: later */
        EncSlot &sl = c->slot[sj];
        u64 f0, nfr, s0, ns;
        chunk_span(cj, f0, nfr, s0, ns);
        const u64 nb = ns * (u64)channels * (u64)cbytes;
        int rc2;
        if ((rc2 = sl.pcm.ensure(nb))) return rc2;
        if (timeline) fprintf(stderr, "host: h2d of chunk %llu enqueued at %.3f\n", (unsigned long long)cj, host_ms());
        CUDA_TRY(cudaEventRecord(sl.ev_c0, c->h2d_stream));
        CUDA_TRY(cudaMemcpyAsync(sl.pcm.p, (const uint8_t *)pcm + s0 * (u64)channels * (u64)cbytes, nb, cudaMemcpyHostToDevice, c->h2d_stream));
        CUDA_TRY(cudaEventRecord(sl.ev_c1, c->h2d_stream));
        h2d_issued[cj] = 1;
        return FLACB_OK;
    };
    for (u64 ci = 0; ci < n_chunks; ci++) {
        const int s = (int)(ci % FLACB_SLOTS);
        EncSlot &sl = c->slot[s];
        /* the slot's previous chunk must be retired (and its output copied) before reuse;
         * chunks are retired in order */
        if ((rc = retire(s))) return rc;
        if ((rc = finish_d2h(s))) return rc;
        if (status != FLACB_OK) break;
        u64 f0, nfr, s0, ns;
        chunk_span(ci, f0, nfr, s0, ns);
        EncodeArgs ca = a;
        ca.n_sample_frames = ns;
        ca.first_frame_number = first_frame_number + f0;
        const uint8_t *d_pcm = (const uint8_t *)pcm + s0 * (u64)channels * (u64)cbytes;
        if (host_in) {
            if ((rc = issue_h2d(ci))) return rc; /* unless it went out an iteration ago */
            CUDA_TRY(cudaStreamWaitEvent(sl.stream, sl.ev_c1, 0));
            d_pcm = (const uint8_t *)sl.pcm.p;
        }
        uint8_t *d_out = out;
        u64 d_cap = out_cap;
        if (host_out) {
            d_cap = flacb_encode_bound(ns, channels, bits, block_size, rice_max);
            if ((rc = sl.out.ensure(d_cap))) return rc;
            d_out = (uint8_t *)sl.out.p;
        }
        if ((rc = launch_encode(c, sl, plan, d_pcm, ca, f0 + nfr != n_frames, d_out, d_cap, d_info, pend[s].meta))) return rc;
        pend[s].f0 = f0;
        pend[s].nfr = nfr;
        pend[s].active = true;
        if ((rc = issue_h2d(ci + 1))) return rc;
        /* Start the output copy of the chunk `ahead` behind this one now (it needs the chunk's
         * size on the host, hence a wait for its kernels) instead of when its slot is needed
         * again: by then the copy is done, and while the host waits here the copy engine and
         * the SMs still have the chunks in between to work on. */
        if (ci >= (u64)ahead && (rc = retire((int)((ci - (u64)ahead) % FLACB_SLOTS)))) return rc;
    }
    /* drain in chunk order */
    for (u64 k = 0; k < FLACB_SLOTS; k++) {
        const int s = (int)((n_chunks + k) % FLACB_SLOTS);
        if ((rc = retire(s))) return rc;
    }
    for (int s = 0; s < FLACB_SLOTS; s++)
        if ((rc = finish_d2h(s))) return rc;
    if (host_in && status != FLACB_OK) CUDA_TRY(cudaStreamSynchronize(c->h2d_stream)); /* a copy queued ahead still reads the caller's buffer */
    if (info && status != FLACB_E_CUDA) {
        CUDA_TRY(cudaMemcpyAsync(info, d_info, n_frames * (u64)channels * sizeof(flacb_subframe_info), cudaMemcpyDeviceToHost,
                                 c->slot[0].stream));
        CUDA_TRY(cudaStreamSynchronize(c->slot[0].stream));
    }
    if (out_len) *out_len = host_off;
    if (stats) {
        stats->frames = n_frames;
        stats->subframes = n_frames * (u64)channels;
        for (int o = 0; o < 5; o++) stats->subframes_fixed[o] = hist[o];
        stats->subframes_lpc = hist[5];
        stats->out_of_domain = ood;
        stats->bytes_in = pcm_bytes;
        stats->bytes_out = host_off;
        stats->first_bad_frame = bad_frame;
        stats->kernel_ms = k_ms;
        for (int k = 0; k < 4; k++) stats->stage_ms[k] = stage_ms[k];
        stats->h2d_ms = h2d_ms;
        stats->d2h_ms = d2h_ms;
        stats->threads_per_frame = threads;
        stats->smem_bytes = smem;
    }
    return status;
}

/* ------------------------------------------------------------------------------------- */

/* Last step of a decode: inter-channel decorrelation + interleave (decoder.py:431-451) into the
 * caller's buffer.  FLACB_E_NOSPACE leaves the decoded frames pending in the context, with the
 * size that is needed in *n_sample_frames and *frame_channels: flacb_decode_output finishes. */
static int decode_emit(flacb_ctx *c, uint32_t flags, int32_t *pcm_out, uint64_t cap_samples, uint64_t *n_sample_frames,
                       int *frame_channels, flacb_stats *stats)
{
    const u64 total = c->pend_dec.total, nsf = c->pend_dec.nsf;
    const int fch = c->pend_dec.fch;
    if (n_sample_frames) *n_sample_frames = nsf;
    if (frame_channels) *frame_channels = fch;
    if (total > cap_samples) return FLACB_E_NOSPACE;
    if (total && !pcm_out) return FLACB_E_ASSERT;
    int rc;
    float d2h_ms = 0.f, k2_ms = 0.f;
    int32_t *d_out = pcm_out;
    if (!(flags & FLACB_OUTPUT_ON_DEVICE)) {
        if ((rc = c->d_pcm.ensure((total + 16) * 4))) return rc;
        d_out = (int32_t *)c->d_pcm.p;
    }
    CUDA_TRY(cudaEventRecord(c->ev_k0, c->stream));
    if (total) {
        interleave_kernel<<<(unsigned)c->pend_dec.n_cand, 256, 0, c->stream>>>((const Candidate *)c->d_cand.p, (const int32_t *)c->d_frames.p,
                                                                               d_out, total, fch);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(c->ev_k1, c->stream));
    if (!(flags & FLACB_OUTPUT_ON_DEVICE) && total) {
        /* same stream: the copy starts when the kernel is done, and its events time the copy alone */
        CUDA_TRY(cudaEventRecord(c->ev_d0, c->stream));
        CUDA_TRY(cudaMemcpyAsync(pcm_out, d_out, total * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev_d1, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaEventElapsedTime(&d2h_ms, c->ev_d0, c->ev_d1));
    } else {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    CUDA_TRY(cudaEventElapsedTime(&k2_ms, c->ev_k0, c->ev_k1));
    c->pend_dec.valid = false;
    if (stats) {
        stats->frames = c->pend_dec.n_frames;
        stats->subframes = c->pend_dec.n_frames * (u64)(fch > 0 ? fch : 0);
        stats->bytes_out = total * 4;
        stats->first_bad_frame = c->pend_dec.status ? (int64_t)c->pend_dec.n_frames : -1;
        stats->kernel_ms = c->pend_dec.k_ms + k2_ms;
        stats->h2d_ms = c->pend_dec.h2d_ms;
        stats->d2h_ms = d2h_ms;
    }
    return c->pend_dec.status;
}

/* flacb_scan_frames: where the frames of the stream are, instead of their samples */
struct ScanOut {
    uint64_t *offsets;      /* cap + 1 */
    uint32_t *block_sizes;  /* cap, may be null */
    uint64_t cap;
    uint64_t n_frames;
};

static int decode_frames_impl(flacb_ctx *c, const uint8_t *frames, uint64_t len, int streaminfo_bits,
                              uint32_t flags, int32_t *pcm_out, uint64_t cap_samples, uint64_t *n_sample_frames,
                              int *frame_channels, flacb_stats *stats, int check_crc8, ScanOut *scan = nullptr)
{
    if (n_sample_frames) *n_sample_frames = 0;
    if (frame_channels) *frame_channels = 0;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->first_bad_frame = -1;
    }
    if (!c || (!frames && len)) return FLACB_E_ASSERT;
    c->pend_dec.valid = false;
    if (len < 2) return FLACB_OK; /* decoder.py:104-108: EOFError ends the frame loop */
    if (len >= (1ull << 40)) return FLACB_E_UNSUPPORTED;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc;
    float h2d_ms = 0.f, k_ms = 0.f;
    const uint8_t *d_s = frames;
    if (!(flags & FLACB_INPUT_ON_DEVICE)) {
        if ((rc = c->d_stream.ensure(len))) return rc;
        CUDA_TRY(cudaEventRecord(c->ev_c0, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->d_stream.p, frames, len, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev_c1, c->stream));
        d_s = (const uint8_t *)c->d_stream.p;
    }
    /* D1: sync candidates in stream order */
    const u64 n_chunks = (len + SYNC_CHUNK - 1) / SYNC_CHUNK;
    if (n_chunks > 0x7FFFFFFFull) return FLACB_E_UNSUPPORTED;
    /* d_dmisc: [counts u32 n_chunks][offsets u64 n_chunks + 1][result] */
    const size_t off_counts = 0;
    const size_t off_offsets = (n_chunks * 4 + 15) & ~(size_t)15;
    const size_t off_result = off_offsets + (n_chunks + 1) * 8;
    if ((rc = c->d_dmisc.ensure(off_result + sizeof(DecodeResult) + 64))) return rc;
    uint8_t *dm = (uint8_t *)c->d_dmisc.p;
    uint32_t *d_counts = (uint32_t *)(dm + off_counts);
    u64 *d_offsets = (u64 *)(dm + off_offsets);
    DecodeResult *d_res = (DecodeResult *)(dm + off_result);
    if ((rc = c->h_dmeta.ensure(256))) return rc;
    u64 *h_u64 = (u64 *)c->h_dmeta.p;

    CUDA_TRY(cudaEventRecord(c->ev_k0, c->stream));
    find_sync_kernel<<<(unsigned)n_chunks, SYNC_THREADS, 0, c->stream>>>(d_s, len, d_counts, d_offsets, nullptr, 0);
    scan_frame_bytes_kernel<<<1, 256, 0, c->stream>>>(d_counts, d_offsets, (uint32_t)n_chunks);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_u64, d_offsets + n_chunks, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const u64 n_cand = h_u64[0];
    if (n_cand == 0) return len >= 2 ? FLACB_E_ASSERT : FLACB_OK; /* decoder.py:137: no sync at the start */
    if (n_cand > 0x7FFFFFF0ull) return FLACB_E_UNSUPPORTED;
    /* d_cand: [Candidate n][sizes u32 n][tmp_off u64 n + 1][jump u32 2n][reach u32 n] */
    const size_t o_sizes = n_cand * sizeof(Candidate);
    const size_t o_tmpoff = (o_sizes + n_cand * 4 + 15) & ~(size_t)15;
    const size_t o_jump = o_tmpoff + (n_cand + 1) * 8;
    const size_t o_reach = o_jump + 2 * n_cand * 4;
    if ((rc = c->d_cand.ensure(o_reach + n_cand * 4 + 64))) return rc;
    uint8_t *dc = (uint8_t *)c->d_cand.p;
    Candidate *d_cand = (Candidate *)dc;
    uint32_t *d_sizes = (uint32_t *)(dc + o_sizes);
    u64 *d_tmpoff = (u64 *)(dc + o_tmpoff);
    uint32_t *d_jump = (uint32_t *)(dc + o_jump);
    uint32_t *d_reach = (uint32_t *)(dc + o_reach);
    find_sync_kernel<<<(unsigned)n_chunks, SYNC_THREADS, 0, c->stream>>>(d_s, len, d_counts, d_offsets, d_cand, 1);
    /* D2: headers, then the scratch layout */
    parse_headers_kernel<<<(unsigned)((n_cand + 127) / 128), 128, 0, c->stream>>>(d_s, len, streaminfo_bits, d_cand,
                                                                                  (uint32_t)n_cand, d_sizes, check_crc8);
    scan_frame_bytes_kernel<<<1, 256, 0, c->stream>>>(d_sizes, d_tmpoff, (uint32_t)n_cand);
    g_launches += 3;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_u64, d_tmpoff + n_cand, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const u64 tmp_samples = h_u64[0];
    if ((rc = c->d_frames.ensure((tmp_samples + 16) * 4))) return rc;
    int32_t *d_tmp = (int32_t *)c->d_frames.p;
    /* D3 + D4 */
    /* Frames per warp.  A lane's walk through its frame is a chain of dependent instructions: few frames per warp
     * mean many warps to hide it behind, until the schedulers are full (about a dozen warps per SM: config 2's
     * 5743 frames among 6919 candidates are fastest at 4 lanes, config 3's 21094 frames at 16). */
    int lanes = 2;
    while (lanes < 32 && n_cand > (u64)lanes * (u64)c->sm_count * 12u) lanes *= 2;
    if (const char *e = getenv("FLACB200_DECODE_LANES")) { /* tuning knob */
        const int v = atoi(e);
        if (v >= 1 && v <= 32) lanes = v;
    }
    /* one record per possible subframe of every candidate: what restore_kernel has to do */
    const u64 n_slots = n_cand * (u64)SUBREC_SLOTS;
    if ((rc = c->d_recs.ensure(n_slots * sizeof(SubRec)))) return rc;
    SubRec *d_recs = (SubRec *)c->d_recs.p;
    CUDA_TRY(cudaMemsetAsync(d_recs, 0, n_slots * sizeof(SubRec), c->stream));
    const u64 n_warps = (n_cand + (u64)lanes - 1) / (u64)lanes;
    decode_frames_kernel<<<(unsigned)((n_warps + DECODE_WARPS - 1) / DECODE_WARPS), 32 * DECODE_WARPS, 0, c->stream>>>(
        d_s, len, d_cand, (uint32_t)n_cand, d_tmpoff, d_tmp, d_recs, check_crc8, lanes);
    restore_kernel<<<(unsigned)((n_slots + 63) / 64), 64, 0, c->stream>>>(d_recs, (uint32_t)n_cand, d_tmp);
    g_launches++;
    chain_kernel<<<1, 1024, 0, c->stream>>>(d_cand, (uint32_t)n_cand, len, d_jump, d_reach, d_res);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
    DecodeResult *h_res = (DecodeResult *)(h_u64 + 8);
    CUDA_TRY(cudaMemcpyAsync(h_res, d_res, sizeof(DecodeResult), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const u64 nsf = h_res->n_sample_frames;
    const int fch = h_res->frame_channels;
    rc = h_res->status;
    if (rc == STATUS_CRC8_MISMATCH || rc == STATUS_DEFERRED) return STATUS_CRC8_MISMATCH; /* the chain reached a frame with a bad header CRC: caller retries unfiltered */
    if (scan) {
        /* the frames the reference visits, in stream order: the candidates on the chain */
        std::vector<Candidate> h_cand(n_cand);
        CUDA_TRY(cudaMemcpy(h_cand.data(), d_cand, n_cand * sizeof(Candidate), cudaMemcpyDeviceToHost));
        u64 k = 0, end = 0;
        for (u64 i = 0; i < n_cand; i++) {
            if (!h_cand[i].on_chain) continue;
            if (k < scan->cap) {
                scan->offsets[k] = h_cand[i].start;
                if (scan->block_sizes) scan->block_sizes[k] = (uint32_t)h_cand[i].bs;
            }
            end = h_cand[i].end;
            k++;
        }
        scan->n_frames = k;
        if (n_sample_frames) *n_sample_frames = nsf;
        if (frame_channels) *frame_channels = fch;
        if (k > scan->cap) return FLACB_E_NOSPACE;
        scan->offsets[k] = end;
        return rc;
    }
    CUDA_TRY(cudaEventRecord(c->ev_k1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&k_ms, c->ev_k0, c->ev_k1));
    if (!(flags & FLACB_INPUT_ON_DEVICE)) CUDA_TRY(cudaEventElapsedTime(&h2d_ms, c->ev_c0, c->ev_c1));
    c->pend_dec.valid = true;
    c->pend_dec.n_cand = n_cand;
    c->pend_dec.total = nsf * (u64)(fch > 0 ? fch : 1);
    c->pend_dec.nsf = nsf;
    c->pend_dec.n_frames = h_res->n_frames;
    c->pend_dec.fch = fch;
    c->pend_dec.status = rc;
    c->pend_dec.k_ms = k_ms;
    c->pend_dec.h2d_ms = h2d_ms;
    if (stats) {
        stats->bytes_in = len;
        stats->out_of_domain = n_cand - h_res->n_frames; /* sync candidates that were not frames */
    }
    return decode_emit(c, flags, pcm_out, cap_samples, n_sample_frames, frame_channels, stats);
}

extern "C" int flacb_decode_frames(flacb_ctx *c, const uint8_t *frames, uint64_t len, int streaminfo_bits,
                                   uint32_t flags, int32_t *pcm_out, uint64_t cap_samples, uint64_t *n_sample_frames,
                                   int *frame_channels, flacb_stats *stats)
{
    /* fast path: sync candidates must carry a valid header CRC-8.  The reference does not verify
     * CRCs (decoder.py:175), so if the frame chain runs into a frame whose CRC-8 is wrong the
     * decode is repeated with every sync pattern as a candidate, which is exactly the reference's
     * behaviour (at the price of parsing the false candidates too). */
    int rc = decode_frames_impl(c, frames, len, streaminfo_bits, flags, pcm_out, cap_samples, n_sample_frames,
                                frame_channels, stats, 1);
    if (rc == STATUS_CRC8_MISMATCH)
        rc = decode_frames_impl(c, frames, len, streaminfo_bits, flags, pcm_out, cap_samples, n_sample_frames,
                                frame_channels, stats, 0);
    return rc;
}

extern "C" int flacb_scan_frames(flacb_ctx *c, const uint8_t *frames, uint64_t len, int streaminfo_bits, uint32_t flags,
                                 uint64_t *offsets, uint32_t *block_sizes, uint64_t cap_frames, uint64_t *n_frames,
                                 uint64_t *n_sample_frames, int *frame_channels)
{
    if (n_frames) *n_frames = 0;
    if (!c || !offsets) return FLACB_E_ASSERT;
    ScanOut so = {offsets, block_sizes, cap_frames, 0};
    offsets[0] = 0;
    int rc = decode_frames_impl(c, frames, len, streaminfo_bits, flags, nullptr, 0, n_sample_frames, frame_channels, nullptr, 1, &so);
    if (rc == STATUS_CRC8_MISMATCH)
        rc = decode_frames_impl(c, frames, len, streaminfo_bits, flags, nullptr, 0, n_sample_frames, frame_channels, nullptr, 0, &so);
    if (n_frames) *n_frames = so.n_frames;
    c->pend_dec.valid = false;
    return rc;
}

extern "C" int flacb_decode_output(flacb_ctx *c, uint32_t flags, int32_t *pcm_out, uint64_t cap_samples,
                                   uint64_t *n_sample_frames, int *frame_channels, flacb_stats *stats)
{
    if (!c || !c->pend_dec.valid) return FLACB_E_ASSERT;
    CUDA_TRY(cudaSetDevice(c->device));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->first_bad_frame = -1;
    }
    return decode_emit(c, flags, pcm_out, cap_samples, n_sample_frames, frame_channels, stats);
}
