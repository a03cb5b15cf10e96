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
#include <string>
#include <vector>

#include "flacb200_host.h"
#include "flacb200_kernels.cuh"
#include "flacb200_decode.cuh"

using namespace flacb;

static thread_local std::string g_cuda_error;
static std::atomic<uint64_t> g_launches{0};

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

struct flacb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_c0 = nullptr, ev_c1 = nullptr;
    int smem_optin = 0, sm_count = 0;
    size_t static_smem12 = 0, static_smem32 = 0;
    DevBuf pcm, out, slots, frame_bytes, frame_off, status, lb, misc, info, win_full, win_tail;
    HostBuf h_meta;
    int win_full_n = -1, win_tail_n = -1;
    bool win_full_ok = true, win_tail_ok = true;
    int flat_full[2] = {0, 0}, flat_tail[2] = {0, 0};
    /* decode */
    DevBuf d_stream, d_pcm, d_cand, d_frames, d_dmisc;
    HostBuf h_dmeta;
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
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    cudaEvent_t *evs[] = {&c->ev_k0, &c->ev_k1, &c->ev_t0, &c->ev_t1, &c->ev_c0, &c->ev_c1};
    for (cudaEvent_t *e : evs) CUDA_TRY(cudaEventCreate(e));
    CUDA_TRY(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, encode_frames_kernel<12, 576, 1>));
    c->static_smem12 = fa.sharedSizeBytes;
    CUDA_TRY(cudaFuncGetAttributes(&fa, encode_frames_kernel<32, 576, 1>));
    c->static_smem32 = fa.sharedSizeBytes;
    /* CRC-16 tables (crc.py:4-14) and the x^n weights of the parallel CRC */
    uint16_t tab[256], lo[256];
    std::vector<uint16_t> hi(XPOW_HI);
    crc16_table(tab);
    crc16_xpow_tables(lo, hi.data(), XPOW_HI);
    CUDA_TRY(cudaMemcpyToSymbol(c_crc16_tab, tab, sizeof(tab)));
    CUDA_TRY(cudaMemcpyToSymbol(c_xpow_lo, lo, sizeof(lo)));
    CUDA_TRY(cudaMemcpyToSymbol(c_xpow_hi, hi.data(), sizeof(uint16_t) * XPOW_HI));
    CUDA_TRY(cudaDeviceSynchronize());
    *out = c;
    return FLACB_OK;
}

extern "C" void flacb_destroy(flacb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&c->pcm, &c->out, &c->slots, &c->frame_bytes, &c->frame_off, &c->status, &c->lb, &c->misc,
                      &c->info, &c->win_full, &c->win_tail, &c->d_stream, &c->d_pcm, &c->d_cand, &c->d_frames,
                      &c->d_dmisc};
    for (DevBuf *b : bufs) b->release();
    c->h_meta.release();
    c->h_dmeta.release();
    cudaEvent_t evs[] = {c->ev_k0, c->ev_k1, c->ev_t0, c->ev_t1, c->ev_c0, c->ev_c1};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
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
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
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

/* One launch over frames [0, n_frames) of device-resident PCM into device-resident output.
 * On return the kernels are enqueued on c->stream and the meta data copy (frame offsets,
 * statuses, counters) to pinned host memory is enqueued behind them. */
struct ChunkMeta {
    u64 *h_frame_off;  /* n_frames + 1 */
    int32_t *h_status; /* n_frames */
    u64 *h_stats;      /* 8 */
    int threads, smem;
};

static int launch_encode(flacb_ctx *c, const uint8_t *d_pcm, const EncodeArgs &a, uint8_t *d_out, u64 out_cap,
                         flacb_subframe_info *d_info, ChunkMeta &meta)
{
    const u64 n_frames = (a.n_sample_frames + (u64)a.block_size - 1) / (u64)a.block_size;
    const int tail = (int)(a.n_sample_frames - (n_frames - 1) * (u64)a.block_size);
    const bool lmax12 = a.max_lpc_order <= 12;
    const size_t static_smem = lmax12 ? c->static_smem12 : c->static_smem32;
    if ((size_t)c->smem_optin <= static_smem + 1024) return FLACB_E_UNSUPPORTED;
    if (frame_bound(a.block_size, a.channels, a.bits, a.rice_max) >= (u64)XPOW_HI * 256) return FLACB_E_UNSUPPORTED;
    auto layout_total = [&](int threads, int all_planes) {
        return (size_t)make_layout(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, threads, all_planes).total;
    };
    LaunchShape shape = pick_shape(a.block_size, E, static_smem, (size_t)c->smem_optin, 228 * 1024, layout_total);
    if (shape.threads == 0) return FLACB_E_UNSUPPORTED; /* the block does not fit shared memory */
    if (const char *e = getenv("FLACB200_SHAPE")) { /* tuning knob: "threads,all_planes,variant" */
        int t = 0, ap = 1, v = 0;
        if (sscanf(e, "%d,%d,%d", &t, &ap, &v) == 3 && t >= 32 && t % 32 == 0 && t <= (v ? 288 : 576) &&
            layout_total(t, ap) + static_smem <= (size_t)c->smem_optin)
            shape.threads = t, shape.all_planes = ap, shape.variant = v;
    }
    const int B = shape.threads;
    const SmemLayout lay = make_layout(a.block_size, a.channels, a.bits, a.rice_max, a.max_lpc_order, B, shape.all_planes);

    int rc;
    if ((rc = upload_window(c, a.block_size, c->win_full, c->win_full_n, c->win_full_ok, c->flat_full[0], c->flat_full[1]))) return rc;
    if ((rc = upload_window(c, tail, c->win_tail, c->win_tail_n, c->win_tail_ok, c->flat_tail[0], c->flat_tail[1]))) return rc;
    if ((rc = c->frame_bytes.ensure(n_frames * 4))) return rc;
    if ((rc = c->frame_off.ensure((n_frames + 1) * 8))) return rc;
    if ((rc = c->status.ensure(n_frames * 4))) return rc;
    if ((rc = c->lb.ensure(n_frames * 8))) return rc;
    if ((rc = c->misc.ensure(128))) return rc;
    if ((rc = c->h_meta.ensure((n_frames + 1) * 8 + n_frames * 4 + 64 + 16))) return rc;

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
    p.win_full = (const double *)c->win_full.p;
    p.win_tail = (const double *)c->win_tail.p;
    p.win_full_bad = !c->win_full_ok;
    p.win_tail_bad = !c->win_tail_ok;
    p.flat_full_lo = c->flat_full[0], p.flat_full_hi = c->flat_full[1];
    p.flat_tail_lo = c->flat_tail[0], p.flat_tail_hi = c->flat_tail[1];
    p.frame_bytes = (uint32_t *)c->frame_bytes.p;
    p.frame_off = (u64 *)c->frame_off.p;
    p.frame_status = (int32_t *)c->status.p;
    p.lb_state = (u64 *)c->lb.p;
    p.ticket = (uint32_t *)c->misc.p;
    p.stats = (u64 *)((uint8_t *)c->misc.p + 64);
    p.info = d_info;
    p.all_planes = shape.all_planes;
    if (a.flags & FLACB_SLOT_COMPACT) {
        p.slot_stride = (frame_bound(a.block_size, a.channels, a.bits, a.rice_max) + 15) & ~15ull;
        if ((rc = c->slots.ensure(p.slot_stride * n_frames))) return rc;
        p.out = (uint8_t *)c->slots.p;
        p.out_cap = p.slot_stride * n_frames;
    } else {
        p.slot_stride = 0;
        p.out = d_out;
        p.out_cap = out_cap;
    }

    CUDA_TRY(cudaMemsetAsync(c->misc.p, 0, 128, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->lb.p, 0, n_frames * 8, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev_k0, c->stream));
#define FLACB_LAUNCH(LM, TB, MB)                                                                                   \
    do {                                                                                                           \
        CUDA_TRY(cudaFuncSetAttribute(encode_frames_kernel<LM, TB, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)lay.total));                                                            \
        encode_frames_kernel<LM, TB, MB><<<(unsigned)n_frames, B, lay.total, c->stream>>>(p);                     \
    } while (0)
    if (lmax12 && shape.variant) FLACB_LAUNCH(12, 288, 3);
    else if (lmax12) FLACB_LAUNCH(12, 576, 1);
    else if (shape.variant) FLACB_LAUNCH(32, 288, 3);
    else FLACB_LAUNCH(32, 576, 1);
#undef FLACB_LAUNCH
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    if (a.flags & FLACB_SLOT_COMPACT) {
        scan_frame_bytes_kernel<<<1, 256, 0, c->stream>>>(p.frame_bytes, p.frame_off, p.n_frames);
        CUDA_TRY(cudaGetLastError());
        compact_slots_kernel<<<(unsigned)n_frames, 256, 0, c->stream>>>((const uint8_t *)c->slots.p, p.slot_stride, p.frame_bytes,
                                                                         p.frame_off, d_out, out_cap, p.frame_status);
        CUDA_TRY(cudaGetLastError());
        g_launches += 2;
    }
    CUDA_TRY(cudaEventRecord(c->ev_k1, c->stream));

    uint8_t *h = (uint8_t *)c->h_meta.p;
    meta.h_frame_off = (u64 *)h;
    meta.h_stats = (u64 *)(h + (n_frames + 1) * 8);
    meta.h_status = (int32_t *)(h + (n_frames + 1) * 8 + 64);
    meta.threads = B;
    meta.smem = (int)lay.total;
    CUDA_TRY(cudaMemcpyAsync(meta.h_frame_off, c->frame_off.p, (n_frames + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(meta.h_stats, p.stats, 64, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(meta.h_status, c->status.p, n_frames * 4, cudaMemcpyDeviceToHost, c->stream));
    return FLACB_OK;
}

extern "C" int flacb_encode_pcm(flacb_ctx *c, const void *pcm, uint64_t n_sample_frames, int channels, int bits,
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
    const u64 pcm_bytes = n_sample_frames * (u64)channels * (u64)cbytes;
    float h2d_ms = 0.f, d2h_ms = 0.f, k_ms = 0.f;

    const uint8_t *d_pcm = (const uint8_t *)pcm;
    if (!(flags & FLACB_INPUT_ON_DEVICE)) {
        if ((rc = c->pcm.ensure(pcm_bytes))) return rc;
        CUDA_TRY(cudaEventRecord(c->ev_c0, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->pcm.p, pcm, pcm_bytes, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev_c1, c->stream));
        d_pcm = (const uint8_t *)c->pcm.p;
    }
    uint8_t *d_out = out;
    u64 d_cap = out_cap;
    if (!(flags & FLACB_OUTPUT_ON_DEVICE)) {
        d_cap = flacb_encode_bound(n_sample_frames, channels, bits, block_size, rice_max);
        if ((rc = c->out.ensure(d_cap))) return rc;
        d_out = (uint8_t *)c->out.p;
    }
    flacb_subframe_info *d_info = nullptr;
    if (info) {
        if ((rc = c->info.ensure(n_frames * (u64)channels * sizeof(flacb_subframe_info)))) return rc;
        d_info = (flacb_subframe_info *)c->info.p;
    }
    ChunkMeta meta;
    if ((rc = launch_encode(c, d_pcm, a, d_out, d_cap, d_info, meta))) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&k_ms, c->ev_k0, c->ev_k1));
    if (!(flags & FLACB_INPUT_ON_DEVICE)) CUDA_TRY(cudaEventElapsedTime(&h2d_ms, c->ev_c0, c->ev_c1));

    /* status of the first frame, in stream order, on which the reference raises */
    u64 valid = meta.h_frame_off[n_frames];
    int64_t bad = -1;
    for (u64 f = 0; f < n_frames; f++) {
        const int s = meta.h_status[f];
        const bool ood = s >= FLACB_E_ASSERT && s <= FLACB_E_ZERODIV;
        if (s != 0 && !(ood && (flags & FLACB_LENIENT))) {
            rc = s;
            bad = (int64_t)f;
            valid = meta.h_frame_off[f];
            break;
        }
    }
    if (!(flags & FLACB_OUTPUT_ON_DEVICE)) {
        if (valid > out_cap) {
            if (rc == FLACB_OK) rc = FLACB_E_NOSPACE;
            valid = 0;
        }
        if (valid) {
            CUDA_TRY(cudaEventRecord(c->ev_c0, c->stream));
            CUDA_TRY(cudaMemcpyAsync(out, d_out, valid, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaEventRecord(c->ev_c1, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            CUDA_TRY(cudaEventElapsedTime(&d2h_ms, c->ev_c0, c->ev_c1));
        }
    }
    if (info) {
        CUDA_TRY(cudaMemcpyAsync(info, d_info, n_frames * (u64)channels * sizeof(flacb_subframe_info), cudaMemcpyDeviceToHost,
                                 c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (out_len) *out_len = valid;
    if (frame_offsets) memcpy(frame_offsets, meta.h_frame_off, (n_frames + 1) * 8);
    if (stats) {
        stats->frames = n_frames;
        stats->subframes = n_frames * (u64)channels;
        for (int o = 0; o < 5; o++) stats->subframes_fixed[o] = meta.h_stats[o];
        stats->subframes_lpc = meta.h_stats[5];
        stats->out_of_domain = meta.h_stats[6];
        stats->bytes_in = pcm_bytes;
        stats->bytes_out = valid;
        stats->first_bad_frame = bad;
        stats->kernel_ms = k_ms;
        stats->h2d_ms = h2d_ms;
        stats->d2h_ms = d2h_ms;
        stats->threads_per_frame = meta.threads;
        stats->smem_bytes = meta.smem;
    }
    return rc;
}

/* ------------------------------------------------------------------------------------- */

static int decode_frames_impl(flacb_ctx *c, const uint8_t *frames, uint64_t len, int streaminfo_bits,
                              uint32_t flags, int32_t *pcm_out, uint64_t cap_samples, uint64_t *n_sample_frames,
                              int *frame_channels, flacb_stats *stats, int check_crc8)
{
    if (n_sample_frames) *n_sample_frames = 0;
    if (frame_channels) *frame_channels = 0;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->first_bad_frame = -1;
    }
    if (!c || (!frames && len)) return FLACB_E_ASSERT;
    if (len < 2) return FLACB_OK; /* decoder.py:104-108: EOFError ends the frame loop */
    if (len >= (1ull << 40)) return FLACB_E_UNSUPPORTED;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc;
    float h2d_ms = 0.f, d2h_ms = 0.f, k_ms = 0.f;
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
    decode_frames_kernel<<<(unsigned)((n_cand + DECODE_WARPS - 1) / DECODE_WARPS), 32 * DECODE_WARPS, 0, c->stream>>>(d_s, len, d_cand, (uint32_t)n_cand,
                                                                               d_tmpoff, d_tmp, check_crc8);
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
    const u64 total = nsf * (u64)(fch > 0 ? fch : 1);
    if (total > cap_samples) {
        CUDA_TRY(cudaEventRecord(c->ev_k1, c->stream));
        return FLACB_E_NOSPACE;
    }
    int32_t *d_out = pcm_out;
    if (!(flags & FLACB_OUTPUT_ON_DEVICE)) {
        if ((rc = c->d_pcm.ensure((total + 16) * 4))) return rc;
        d_out = (int32_t *)c->d_pcm.p;
        rc = h_res->status;
    }
    if (total) {
        interleave_kernel<<<(unsigned)n_cand, 256, 0, c->stream>>>(d_cand, d_tmp, d_out, total, fch);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(c->ev_k1, c->stream));
    if (!(flags & FLACB_OUTPUT_ON_DEVICE) && total) {
        CUDA_TRY(cudaEventRecord(c->ev_c0, c->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaMemcpyAsync(pcm_out, d_out, total * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        CUDA_TRY(cudaEventRecord(c->ev_c1, c->copy_stream));
        CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
        CUDA_TRY(cudaEventElapsedTime(&d2h_ms, c->ev_c0, c->ev_c1));
    } else {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    CUDA_TRY(cudaEventElapsedTime(&k_ms, c->ev_k0, c->ev_k1));
    if (n_sample_frames) *n_sample_frames = nsf;
    if (frame_channels) *frame_channels = fch;
    if (stats) {
        stats->frames = h_res->n_frames;
        stats->subframes = h_res->n_frames * (u64)(fch > 0 ? fch : 0);
        stats->bytes_in = len;
        stats->bytes_out = total * 4;
        stats->first_bad_frame = rc ? (int64_t)h_res->n_frames : -1;
        stats->kernel_ms = k_ms;
        stats->h2d_ms = h2d_ms;
        stats->d2h_ms = d2h_ms;
        stats->out_of_domain = n_cand - h_res->n_frames; /* sync candidates that were not frames */
    }
    return rc;
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
