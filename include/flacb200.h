/*
 * flacb200.h -- C ABI of libflacb200.so: the B200 (sm_100a) FLAC encode / decode hot path
 * that sits behind the Python entry points of turlando/flac-py.
 *
 * The reference has no FFI of its own (it is pure Python); the boundary it exposes is
 *   flac.encoder.encode(sample_rate, sample_size, channels, frames, samples, parameters)
 *       -> Iterator[bytes]                                   (reference flac/encoder.py:48-165)
 *   flac.decoder.decode(buffer) -> (rate, bits, channels, total, Iterator[list[int]])
 *                                                            (reference flac/decoder.py:31-63)
 * The functions below are what the drop-in package (flac-py_b200/flac/) binds with ctypes to
 * replace the bodies of those two functions; INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes, no exceptions, no ownership transfer, every call
 * returns a status code (0 = ok).  A context belongs to one device and is not thread-safe;
 * contexts are independent (one per GPU, driven from one host thread / process each).
 * Status codes carry the Python exception class the reference raises in the same situation.
 */
#ifndef FLACB200_H
#define FLACB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLACB_VERSION 0x000100

enum {
    FLACB_OK = 0,
    FLACB_E_ASSERT = 1,      /* AssertionError  (encoder.py:40-43, :56-57, :157, :508, :619, :669) */
    FLACB_E_VALUE = 2,       /* ValueError      (encoder.py:255, :753, :758; coded_number.py:38)   */
    FLACB_E_ZERODIV = 3,     /* ZeroDivisionError (encoder.py:437, :469)                           */
    FLACB_E_EOF = 4,         /* EOFError        (binary.py:40)                                     */
    FLACB_E_NOSPACE = 5,     /* caller's output buffer too small (not a reference error)           */
    FLACB_E_OVERFLOW = 6,    /* OverflowError   (float pow / floor(inf))                           */
    FLACB_E_KEY = 7,         /* KeyError        (decoder.py:220 sample-rate table)                 */
    FLACB_E_NOTIMPL = 8,     /* NotImplementedError (encoder.py:772)                               */
    FLACB_E_CUDA = 9,        /* CUDA runtime error; flacb_last_cuda_error() has the text           */
    FLACB_E_UNSUPPORTED = 10,/* valid for the reference but outside what the kernels handle        */
    FLACB_E_RANGE = 11,      /* a residual does not fit 32 bits (never on in-range PCM)            */
    FLACB_E_INTERNAL = 12    /* look-back spin limit or similar: a bug, please report              */
};

/* flags for flacb_encode_* */
#define FLACB_SPEC_CHANNELS   0x01u /* frame header carries channels-1 instead of the reference's
                                       constant L_R nibble (encoder.py:95)                        */
#define FLACB_INPUT_ON_DEVICE 0x02u /* pcm is a device pointer on the context's device            */
#define FLACB_OUTPUT_ON_DEVICE 0x04u/* out is a device pointer (frame_offsets stays host memory)  */
#define FLACB_LENIENT         0x08u /* do not fail where the reference raises on out-of-domain
                                       input (silence, DC, ...): Rice parameter 0 when the mean is
                                       below 1, no LPC candidate when it cannot be computed, fixed
                                       wins ties.  Frames concerned are counted in the stats.     */
#define FLACB_SLOT_COMPACT    0x10u /* debugging: per-frame worst-case slots + compaction kernel
                                       instead of the single-pass look-back placement            */
#define FLACB_PCM_INT32       0x20u /* samples are int32 containers (whatever `bits` is); default
                                       is the WAV payload layout, ceil(bits/8) LE bytes/sample    */

#define FLACB_STAGE_TIMES     0x40u /* time every stage of an LPC encode on its own (flacb_stats.stage_ms);
                                       without it the stages are launched back to back with programmatic
                                       dependent launch, the start of one overlapping the end of the one in
                                       front, and stage_ms is left at zero                                   */

#define FLACB_MAX_LPC 32
#define FLACB_INFO_PARAMS 64

typedef struct flacb_ctx flacb_ctx;

typedef struct {
    uint64_t frames;            /* frames encoded / decoded by the call                      */
    uint64_t subframes;
    uint64_t subframes_fixed[5];/* winner histogram by fixed order                           */
    uint64_t subframes_lpc;     /* subframes where SubframeLPC won (encoder.py:147)          */
    uint64_t out_of_domain;     /* frames on which the reference would raise (lenient mode)  */
    uint64_t bytes_in, bytes_out;
    int64_t  first_bad_frame;   /* frame index (relative to the call) of the status, or -1   */
    float    kernel_ms;         /* CUDA-event time of the kernels of the call                */
    float    h2d_ms, d2h_ms;    /* copy time when host buffers are used (0 otherwise)        */
    int32_t  threads_per_frame; /* launch shape actually used (serialising kernel)           */
    int32_t  smem_bytes;
    float    stage_ms[4];       /* CUDA-event time of the stages of an LPC encode: analysis,
                                   coefficients, order choice, serialisation (fixed predictors
                                   only: everything under [3])                               */
} flacb_stats;

/* Per-subframe introspection (stage-level parity tests; optional, small inputs only). */
typedef struct {
    int32_t type;       /* 0 fixed, 1 lpc */
    int32_t order;      /* order written in the subframe header */
    int32_t shift;
    int32_t precision;
    int32_t porder;     /* Rice partition order (encoder.py:688) */
    int32_t method;     /* 4 or 5 (encoder.py:648-650) */
    int32_t fixed_order;
    int32_t lpc_order;  /* best LPC order, 1-based; 0 when LPC is disabled */
    int64_t fixed_err[5];
    int64_t lpc_err[FLACB_MAX_LPC];
    int32_t lpc_shift[FLACB_MAX_LPC];
    int32_t lpc_ncoef[FLACB_MAX_LPC];
    int32_t lpc_q[FLACB_MAX_LPC][FLACB_MAX_LPC];
    double  autoc[FLACB_MAX_LPC + 1];
    int64_t bits;       /* serialised size of the subframe */
    int32_t nparams;
    uint8_t params[FLACB_INFO_PARAMS];
} flacb_subframe_info;

int         flacb_version(void);
int         flacb_device_count(void);
const char *flacb_strerror(int status);
const char *flacb_last_cuda_error(void);

int  flacb_create(int device, flacb_ctx **out);
void flacb_destroy(flacb_ctx *ctx);

/* Pinned host memory for the end-to-end path. */
int  flacb_host_alloc(void **p, size_t bytes);
void flacb_host_free(void *p);
/* Device memory helpers (so a host language without a CUDA binding can keep data resident). */
int  flacb_device_alloc(flacb_ctx *ctx, void **p, size_t bytes);
void flacb_device_free(flacb_ctx *ctx, void *p);
int  flacb_memcpy_h2d(flacb_ctx *ctx, void *dst, const void *src, size_t bytes);
int  flacb_memcpy_d2h(flacb_ctx *ctx, void *dst, const void *src, size_t bytes);

/* Upper bound on the encoded size of n_sample_frames (frames only, no stream prologue). */
uint64_t flacb_encode_bound(uint64_t n_sample_frames, int channels, int bits, int block_size,
                            int rice_max);

/*
 * Encode the audio frames of `n_sample_frames` sample frames of interleaved little-endian PCM:
 * the body of the `for i, xs in enumerate(batch(samples, block_size))` loop of
 * flac/encoder.py:87-165 for every block, i.e. frame header (encoder.py:194-234), per channel
 * SubframeFixed analysis (encoder.py:331-359), SubframeLPC analysis (encoder.py:362-534), the
 * choice between them (encoder.py:135-157), Rice partition search (encoder.py:632-760),
 * serialisation (encoder.py:553-627, :765-806; binary.py:144-216), padding and CRC-16
 * (encoder.py:159-165; crc.py).  The stream prologue (encoder.py:61-83) is 42 bytes the host
 * writes itself (flacb_stream_prologue).
 *
 *   max_lpc_order  -l of the CLI (lpc_order = range(l + 1)); 0 = fixed predictors only (an
 *                  extension: the reference cannot express it)
 *   rice_min/max   rice_partition_order = range(rice_min, rice_max + 1)
 *   first_frame_number  coded number of the first frame (frame-range sharding across GPUs)
 *   out/out_cap    receives the frames back to back; *out_len bytes are valid
 *   frame_offsets  optional, HOST memory, n_frames + 1 byte offsets into out
 *   info           optional, HOST memory, n_frames * channels records
 *   stats          optional
 * Returns the status of the first frame (in stream order) on which the reference would raise,
 * FLACB_OK otherwise.
 */
int flacb_encode_pcm(flacb_ctx *ctx, const void *pcm, uint64_t n_sample_frames, int channels,
                     int bits, int block_size, int max_lpc_order, int qlp_precision,
                     int rice_min, int rice_max, uint64_t first_frame_number, uint32_t flags,
                     uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                     uint64_t *frame_offsets, flacb_subframe_info *info, flacb_stats *stats);

/* "fLaC" + metadata block header + STREAMINFO exactly as encoder.py:61-83, :170-189 write them. */
int flacb_stream_prologue(uint32_t sample_rate, int channels, int bits, uint64_t total_samples,
                          int block_size, uint8_t out[42]);

/*
 * Decode the audio frames of a stream (everything after the metadata): flac/decoder.py:100-130
 * get_frames/get_frame, :267-421 subframes and residuals, :431-498 decode_frame.  `frames`
 * points at the first frame header.  pcm_out receives interleaved int32 sample frames with the
 * channel count the frame headers carry (decoder.py:115-122), returned in *frame_channels.
 * A truncated last frame is dropped like decoder.py:104-108 does.  CRCs are not verified by the
 * reference (decoder.py:128, :175); here a CRC mismatch only rejects false sync candidates.
 */
int flacb_decode_frames(flacb_ctx *ctx, const uint8_t *frames, uint64_t len, int streaminfo_bits,
                        uint32_t flags, int32_t *pcm_out, uint64_t cap_samples,
                        uint64_t *n_sample_frames, int *frame_channels, flacb_stats *stats);

/*
 * When flacb_decode_frames returns FLACB_E_NOSPACE, *n_sample_frames and *frame_channels hold what
 * the stream decodes to and the decoded frames stay in the context: this call writes them to a
 * buffer of at least n_sample_frames * frame_channels samples without decoding again.  (The
 * reference never sizes anything in advance -- decoder.py:46-55 yields sample by sample -- and
 * STREAMINFO's sample count is not to be trusted for an allocation.)  Valid until the next
 * decode call on the context.
 */
int flacb_decode_output(flacb_ctx *ctx, uint32_t flags, int32_t *pcm_out, uint64_t cap_samples,
                        uint64_t *n_sample_frames, int *frame_channels, flacb_stats *stats);

/*
 * The frame index of a stream: byte offsets (relative to `frames`) of the frames the reference's
 * get_frames loop (decoder.py:100-108) visits, in stream order, and offsets[n] = the end of the
 * last one; optionally their block sizes.  The reference can only find frame k by parsing frames
 * 0 .. k-1 (frame lengths are not stored, STREAMINFO's frame sizes are zero, encoder.py:75-76);
 * here every sync candidate is parsed in parallel and the real frames are the chain that starts
 * at offset 0 -- the first half of flacb_decode_frames as a stage of its own, for seeking,
 * SEEKTABLE writing (flac.encoder.seektable_block) and sharding a decode.  FLACB_E_NOSPACE:
 * *n_frames frames exist, more than cap_frames.  The status is the one the decode would return.
 */
int flacb_scan_frames(flacb_ctx *ctx, const uint8_t *frames, uint64_t len, int streaminfo_bits,
                      uint32_t flags, uint64_t *offsets, uint32_t *block_sizes, uint64_t cap_frames,
                      uint64_t *n_frames, uint64_t *n_sample_frames, int *frame_channels);

/* CUDA-event timer on the context's stream, for the benchmark (kernel-only throughput). */
int flacb_timer_begin(flacb_ctx *ctx);
int flacb_timer_end(flacb_ctx *ctx, float *ms);
int flacb_sync(flacb_ctx *ctx);

/* how many kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t flacb_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif
