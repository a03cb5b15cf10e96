/*
 * flac_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, CPU-only restatement of turlando/flac-py's encode / decode hot path
 * (reference: flac/encoder.py, flac/decoder.py, flac/binary.py, flac/crc.py,
 * flac/coded_number.py, flac/utils.py).  It exists to CHECK the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  Nothing under flac-py_b200/ links or imports
 * it.  Pinned against the reference itself: tests/golden/ holds outputs generated
 * by importing the Python reference (tests/golden/make_golden.py).
 */
#ifndef FLAC_ORACLE_H
#define FLAC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes; the names follow the Python exception the reference raises */
enum {
    FO_OK = 0,
    FO_E_ASSERT = 1,   /* AssertionError */
    FO_E_VALUE = 2,    /* ValueError (math domain error, negative shift count, ...) */
    FO_E_ZERODIV = 3,  /* ZeroDivisionError */
    FO_E_EOF = 4,      /* EOFError */
    FO_E_NOSPACE = 5,  /* caller's output buffer too small (not a reference error) */
    FO_E_OVERFLOW = 6, /* OverflowError */
    FO_E_KEY = 7,      /* KeyError */
    FO_E_NOTIMPL = 8   /* NotImplementedError */
};

#define FO_FLAG_SPEC_CHANNELS 1u /* write channels-1 in the frame header instead of the
                                    reference's constant L_R nibble (encoder.py:95) */

#define FO_MAX_LPC 32
#define FO_INFO_PARAMS 512

/* Per-subframe introspection record (what the reference computed on the way). */
typedef struct {
    int32_t type;      /* 0 = fixed, 1 = lpc (the winner, encoder.py:138-157) */
    int32_t order;     /* winner's order as written in the subframe header */
    int32_t shift;     /* lpc only */
    int32_t precision; /* lpc only */
    int32_t porder;    /* Rice partition order chosen (encoder.py:688) */
    int32_t method;    /* 4 or 5 (encoder.py:648-650) */
    int32_t fixed_order; /* best fixed order (encoder.py:352) */
    int32_t lpc_order;   /* best lpc order, 1-based (encoder.py:404); 0 if lpc disabled */
    int64_t fixed_err[5];            /* encoder.py:350 */
    int64_t lpc_err[FO_MAX_LPC];     /* encoder.py:398 */
    int32_t lpc_shift[FO_MAX_LPC];   /* per order, encoder.py:384 */
    int32_t lpc_ncoef[FO_MAX_LPC];   /* 0 when the negative-shift branch returned [] */
    int32_t lpc_q[FO_MAX_LPC][FO_MAX_LPC];
    double autoc[FO_MAX_LPC + 1];    /* encoder.py:371 */
    int32_t qcoef[FO_MAX_LPC];       /* winner's coefficients */
    int64_t bits;                    /* size of the serialised subframe in bits */
    int32_t nparams;                 /* 1 << porder */
    uint8_t params[FO_INFO_PARAMS];  /* first FO_INFO_PARAMS Rice parameters */
} fo_subframe_info;

/* Encode the frames of `n_sample_frames` sample frames of interleaved int32 PCM
 * (encoder.py:87-165).  Emits frames only; the stream prologue is
 * fo_stream_prologue().  max_lpc_order == 0 disables the LPC candidate (the
 * "fixed predictors only" configuration; identical to the reference run with
 * encode_subframe_lpc stubbed so that fixed always wins).
 * frame_offsets, if non-NULL, receives n_frames + 1 byte offsets into `out`.
 * info, if non-NULL, receives n_frames * channels records.
 * On error returns the status, *err_frame = failing frame, *err_line = source line here. */
int fo_encode_frames(const int32_t *pcm, uint64_t n_sample_frames, int channels, int bps,
                     int block_size, int max_lpc_order, int qlp_precision,
                     int rice_min, int rice_max, uint64_t first_frame_number, uint32_t flags,
                     uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                     uint64_t *frame_offsets, fo_subframe_info *info,
                     int64_t *err_frame, int *err_line);

/* Same, fanned out over `nthreads` POSIX threads by contiguous frame range. */
int fo_encode_frames_mt(const int32_t *pcm, uint64_t n_sample_frames, int channels, int bps,
                        int block_size, int max_lpc_order, int qlp_precision,
                        int rice_min, int rice_max, uint64_t first_frame_number, uint32_t flags,
                        uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                        uint64_t *frame_offsets, int nthreads,
                        int64_t *err_frame, int *err_line);

/* "fLaC" + metadata block header + STREAMINFO (encoder.py:61-83, 170-189): 42 bytes. */
int fo_stream_prologue(uint32_t sample_rate, int channels, int bps, uint64_t total_samples,
                       int block_size, uint8_t out[42]);

/* Decode a whole stream (decoder.py:31-63).  pcm_out receives interleaved int32 sample
 * frames; *n_out = sample frames produced.  Channel count of the output is the one the
 * FRAME HEADERS carry (decoder.py:115-122), returned in *frame_channels. */
int fo_decode_stream(const uint8_t *stream, uint64_t len,
                     uint32_t *sample_rate, int *bps, int *channels, uint64_t *total_samples,
                     int32_t *pcm_out, uint64_t cap_samples, uint64_t *n_out,
                     int *frame_channels, int *err_line);

/* helpers exported for known-answer tests */
uint8_t fo_crc8(const uint8_t *p, uint64_t n);   /* crc.py:17-22, poly 0x107 */
uint16_t fo_crc16(const uint8_t *p, uint64_t n); /* crc.py:25-30, poly 0x18005 */
int fo_coded_number(uint64_t x, uint8_t out[7]); /* coded_number.py:7-39; returns length or -1 */
double fo_neumaier_sum(const double *x, uint64_t n); /* CPython >= 3.12 builtin sum() on floats */
void fo_tukey(int n, double *w);                 /* encoder.py:423-440; n in 4..7 is the caller's problem */

#ifdef __cplusplus
}
#endif
#endif
