/*
 * flac_oracle.c -- TEST INFRASTRUCTURE ONLY (see flac_oracle.h).
 *
 * CPU restatement of turlando/flac-py's encoder / decoder.  Every function cites
 * the reference file:line it follows.  Deliberately scalar and literal: the goal
 * is to reproduce the reference's results bit for bit, including its quirks
 * (SURVEY.md section 9), not to be fast.  Build with -O2 -ffp-contract=off
 * (no FMA contraction, no fast-math) -- see oracle/Makefile.
 *
 * Floating point that matters (encode_subframe_lpc):
 *   - builtin sum() over floats is Neumaier-compensated in CPython >= 3.12
 *     (Python/bltinmodule.c); restated in fo_neumaier_sum().
 *   - `x ** 2` is libm pow(|x|, 2.0) (Objects/floatobject.c float_pow); the
 *     exponent is read through a volatile so gcc cannot fold it to x*x.
 *   - math.cos / math.log2 are libm cos / log2; round() is rint() (half-even).
 */
#include "flac_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FAIL(code)                 \
    do {                           \
        if (err_line) *err_line = __LINE__; \
        return (code);             \
    } while (0)

/* ------------------------------------------------------------------------- */
/* crc.py:4-30                                                               */

static uint8_t crc8_table[256];
static uint16_t crc16_table[256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;

static void crc_init(void)
{
    /* crc.py:5-14 crc_table(n, generator) */
    for (int d = 0; d < 256; d++) {
        uint32_t v8 = (uint32_t)d; /* d << (8 - 8) */
        uint32_t v16 = (uint32_t)d << 8;
        for (int i = 0; i < 8; i++) {
            v8 <<= 1;
            if (v8 & 0x100u) v8 ^= 0x107u;
            v16 <<= 1;
            if (v16 & 0x10000u) v16 ^= 0x18005u;
        }
        crc8_table[d] = (uint8_t)v8;
        crc16_table[d] = (uint16_t)v16;
    }
}

uint8_t fo_crc8(const uint8_t *p, uint64_t n)
{
    pthread_once(&crc_once, crc_init);
    uint8_t crc = 0;
    for (uint64_t i = 0; i < n; i++) crc = crc8_table[p[i] ^ crc]; /* crc.py:21 */
    return crc;
}

uint16_t fo_crc16(const uint8_t *p, uint64_t n)
{
    pthread_once(&crc_once, crc_init);
    uint16_t crc = 0;
    for (uint64_t i = 0; i < n; i++) /* crc.py:29 */
        crc = (uint16_t)((crc << 8) ^ crc16_table[(crc >> 8) ^ p[i]]);
    return crc;
}

/* ------------------------------------------------------------------------- */
/* coded_number.py:7-39                                                      */

static int bit_length_u64(uint64_t x)
{
    int n = 0;
    while (x) {
        n++;
        x >>= 1;
    }
    return n;
}

int fo_coded_number(uint64_t x, uint8_t out[7])
{
    int n = bit_length_u64(x);
    int size;
    if (n > 36) return -1; /* coded_number.py:8 assert */
    if (n <= 7) size = 1;
    else if (n <= 11) size = 2;
    else if (n <= 16) size = 3;
    else if (n <= 21) size = 4;
    else if (n <= 26) size = 5;
    else if (n <= 31) size = 6;
    else return -2; /* coded_number.py:38 ValueError */
    if (size == 1) {
        out[0] = (uint8_t)x;
        return 1;
    }
    /* coded_number.py:16-21 */
    for (int i = 0; i < size; i++) {
        uint8_t b = (uint8_t)((x >> (6 * (size - 1 - i))) & 0x3F);
        if (i == 0) out[0] = (uint8_t)((((1u << size) - 1u) << (8 - size)) | b);
        else out[i] = (uint8_t)(0x80 | b);
    }
    return size;
}

/* ------------------------------------------------------------------------- */
/* binary.py:144-216  Put                                                    */

typedef struct {
    uint8_t *buf;
    uint64_t cap;
    uint64_t nbits;
    int overflow;
} put_t;

static void put_init(put_t *p, uint8_t *buf, uint64_t cap)
{
    p->buf = buf;
    p->cap = cap;
    p->nbits = 0;
    p->overflow = 0;
}

/* Put.uint(x, n): the n low bits of x, MSB first (binary.py:169-206).  x is masked,
 * which is how negative values become two's complement (binary.py:172, :30). */
static void put_uint(put_t *p, uint64_t x, int n)
{
    for (int i = n - 1; i >= 0; i--) {
        uint64_t byte = p->nbits >> 3;
        if (byte >= p->cap) {
            p->overflow = 1;
            return;
        }
        int bit = 7 - (int)(p->nbits & 7);
        if (bit == 7) p->buf[byte] = 0;
        if (i < 64 && ((x >> i) & 1)) p->buf[byte] |= (uint8_t)(1u << bit);
        p->nbits++;
    }
}

/* put.uint(0, q) for arbitrarily long q (encoder.py:802) */
static void put_zeros(put_t *p, uint64_t q)
{
    while (q > 0) {
        int c = q > 32 ? 32 : (int)q;
        put_uint(p, 0, c);
        q -= (uint64_t)c;
    }
}

/* ------------------------------------------------------------------------- */
/* CPython >= 3.12 builtin sum() over floats (Python/bltinmodule.c,          */
/* builtin_sum_impl, float fast path): the first item seeds the running sum, */
/* the rest are added with Neumaier's compensated scheme and the             */
/* compensation is folded in at the end.                                     */

double fo_neumaier_sum(const double *x, uint64_t n)
{
    if (n == 0) return 0.0;
    double f = 0.0 + x[0];
    double c = 0.0;
    for (uint64_t i = 1; i < n; i++) {
        double v = x[i];
        double t = f + v;
        if (fabs(f) >= fabs(v)) c += (f - t) + v;
        else c += (v - t) + f;
        f = t;
    }
    if (c != 0.0 && isfinite(c)) f += c;
    return f;
}

/* ------------------------------------------------------------------------- */
/* encoder.py:423-440  tukey(n, 0.5)                                         */

void fo_tukey(int n, double *w)
{
    int nr = (int)floor(0.5 / 2.0 * (double)n) - 1;
    for (int i = 0; i < n; i++) w[i] = 1.0;
    for (int i = 0; i <= nr; i++) {
        w[i] = 0.5 - 0.5 * cos(M_PI * (double)i / (double)nr);
        w[n - nr - 1 + i] = 0.5 - 0.5 * cos(M_PI * (double)(i + nr) / (double)nr);
    }
}

/* ------------------------------------------------------------------------- */
/* utils.py:91-94 zigzag_encode                                              */

static inline uint64_t zigzag(int64_t x)
{
    return (uint64_t)((x >> 63) ^ (int64_t)((uint64_t)x << 1));
}

/* ------------------------------------------------------------------------- */
/* encoder.py:537-548 prediction_residual                                    */

static void prediction_residual(const int64_t *x, int n, const int32_t *coef, int order,
                                int shift, int64_t *res /* n - order values */)
{
    for (int i = order; i < n; i++) {
        int64_t s = 0;
        for (int j = 0; j < order; j++) s += x[i - 1 - j] * (int64_t)coef[j];
        res[i - order] = x[i] - (s >> shift); /* arithmetic shift == Python >> */
    }
}

static int64_t sum_abs(const int64_t *r, int n)
{
    int64_t s = 0;
    for (int i = 0; i < n; i++) s += r[i] < 0 ? -r[i] : r[i];
    return s;
}

static const int32_t FIXED_COEF[5][4] = {
    /* common.py:15-21 */
    {0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};

/* ------------------------------------------------------------------------- */
/* per-thread scratch                                                        */

typedef struct {
    int n_alloc;
    int64_t *x;     /* samples of one channel */
    int64_t *res;   /* candidate residual */
    int64_t *best;  /* winning residual of a predictor family */
    int64_t *fixed_res;
    int64_t *lpc_res;
    uint64_t *zz;
    double *win;
    int win_n;
    double *wx;
    double *prod;
    uint8_t *params;     /* chosen parameters */
    uint8_t *params_tmp; /* candidate parameters */
} scratch_t;

static int scratch_init(scratch_t *s, int n, int max_parts)
{
    memset(s, 0, sizeof(*s));
    s->n_alloc = n;
    s->x = malloc(sizeof(int64_t) * (size_t)n);
    s->res = malloc(sizeof(int64_t) * (size_t)n);
    s->best = malloc(sizeof(int64_t) * (size_t)n);
    s->fixed_res = malloc(sizeof(int64_t) * (size_t)n);
    s->lpc_res = malloc(sizeof(int64_t) * (size_t)n);
    s->zz = malloc(sizeof(uint64_t) * (size_t)n);
    s->win = malloc(sizeof(double) * (size_t)n);
    s->wx = malloc(sizeof(double) * (size_t)n);
    s->prod = malloc(sizeof(double) * (size_t)n);
    s->params = malloc((size_t)max_parts);
    s->params_tmp = malloc((size_t)max_parts);
    s->win_n = -1;
    return s->x && s->res && s->best && s->fixed_res && s->lpc_res && s->zz && s->win && s->wx &&
           s->prod && s->params && s->params_tmp;
}

static void scratch_free(scratch_t *s)
{
    free(s->x);
    free(s->res);
    free(s->best);
    free(s->fixed_res);
    free(s->lpc_res);
    free(s->zz);
    free(s->win);
    free(s->wx);
    free(s->prod);
    free(s->params);
    free(s->params_tmp);
}

/* ------------------------------------------------------------------------- */
/* encoder.py:331-359 encode_subframe_fixed                                  */

static void subframe_fixed(scratch_t *s, int n, int *order_out, int64_t *res_out, int *res_len,
                           int64_t *total_out, int64_t err5[5])
{
    if (n <= 4) { /* encoder.py:334-338 */
        prediction_residual(s->x, n, FIXED_COEF[0], 0, 0, res_out);
        *order_out = 0;
        *res_len = n;
        *total_out = sum_abs(res_out, n);
        for (int o = 0; o < 5; o++) err5[o] = o == 0 ? *total_out : -1;
        return;
    }
    int best = 0;
    for (int o = 0; o < 5; o++) { /* encoder.py:341-350 */
        prediction_residual(s->x, n, FIXED_COEF[o], o, 0, s->res);
        err5[o] = sum_abs(s->res, n - o);
        if (o == 0 || err5[o] < err5[best]) { /* first minimum, encoder.py:352 */
            best = o;
            memcpy(res_out, s->res, sizeof(int64_t) * (size_t)(n - o));
        }
    }
    *order_out = best;
    *res_len = n - best;
    *total_out = err5[best];
}

/* ------------------------------------------------------------------------- */
/* encoder.py:453-479 levinson_durbin, run once to max order: the state after
 * iteration k is exactly levinson_durbin(xs[:k+2]).  coefs_out[k]
 * receives coefs[1..k+1] after iteration k.                                  */

static volatile double TWO = 2.0;

static int levinson_all(const double *xs, int order, double coefs_out[FO_MAX_LPC][FO_MAX_LPC],
                        int *err_line)
{
    double coefs[FO_MAX_LPC + 1];
    for (int i = 0; i <= order; i++) coefs[i] = 0.0;
    coefs[0] = 1.0;
    double error = xs[0];
    for (int k = 0; k < order; k++) {
        double lambda = 0.0;
        for (int j = 0; j <= k; j++) lambda -= coefs[j] * xs[k + 1 - j]; /* encoder.py:467 */
        if (error == 0.0) FAIL(FO_E_ZERODIV);                           /* encoder.py:469 */
        lambda /= error;
        for (int n = 0; n <= (k + 1) / 2; n++) { /* encoder.py:471-474 */
            double temp = coefs[k + 1 - n] + lambda * coefs[n];
            coefs[n] = coefs[n] + lambda * coefs[k + 1 - n];
            coefs[k + 1 - n] = temp;
        }
        double l2 = fabs(lambda) == 1.0 ? 1.0 : pow(fabs(lambda), TWO); /* float_pow */
        if (isinf(l2) && !isinf(lambda)) FAIL(FO_E_OVERFLOW);
        error *= 1.0 - l2; /* encoder.py:476 */
        for (int j = 0; j <= k; j++) coefs_out[k][j] = coefs[j + 1];
    }
    return FO_OK;
}

/* encoder.py:482-534 quantize_lpc_coefficients.  Returns the number of coefficients
 * actually produced (0 in the negative-shift branch, which never appends). */
static int quantize_coefs(const double *c, int order, int precision, int32_t *q, int *shift_out,
                          int *ncoef, int *err_line)
{
    double cmax = 0.0;
    int any_nan = 0;
    for (int i = 0; i < order; i++) {
        double a = fabs(c[i]);
        if (isnan(a)) any_nan = 1;
        if (a > cmax) cmax = a;
    }
    if (any_nan || !(cmax > 0.0)) FAIL(FO_E_ASSERT); /* encoder.py:496 */
    if (isinf(cmax)) FAIL(FO_E_OVERFLOW);             /* floor(inf) */
    int shift_max = (1 << (5 - 1)) - 1;
    int shift_min = -(1 << (5 - 1));
    int shift = precision - (int)floor(log2(cmax)) - 2; /* encoder.py:503 */
    if (shift > shift_max) shift = shift_max;
    else if (shift < shift_min) FAIL(FO_E_ASSERT); /* encoder.py:508 */
    int64_t qmax = ((int64_t)1 << (precision - 1)) - 1;
    int64_t qmin = -((int64_t)1 << (precision - 1));
    double error = 0.0;
    if (shift >= 0) { /* encoder.py:517-522 */
        for (int i = 0; i < order; i++) {
            error += c[i] * (double)((int64_t)1 << shift);
            double r = rint(error); /* round(): half to even */
            int64_t qi = r < (double)qmin ? qmin : (r > (double)qmax ? qmax : (int64_t)r);
            error -= (double)qi;
            q[i] = (int32_t)qi;
        }
        *ncoef = order;
        *shift_out = shift;
    } else { /* encoder.py:523-532: computes but never appends; returns ([], 0) */
        *ncoef = 0;
        *shift_out = 0;
    }
    return FO_OK;
}

/* ------------------------------------------------------------------------- */
/* encoder.py:362-420 encode_subframe_lpc                                    */

typedef struct {
    int order; /* 1-based order written in the header (index + 1) */
    int ncoef; /* len(coefficients): 0 in the negative-shift quirk */
    int shift;
    int32_t q[FO_MAX_LPC];
    int res_len;
    int64_t total;
} lpc_result_t;

static int subframe_lpc(scratch_t *s, int n, int L, int precision, lpc_result_t *out,
                        int64_t *res_out, fo_subframe_info *info, int *err_line)
{
    /* 1. window (encoder.py:368, :423-440) */
    if (s->win_n != n) {
        int nr = (int)floor(0.5 / 2.0 * (double)n) - 1;
        if (nr == 0) FAIL(FO_E_ZERODIV); /* pi * i / nr, encoder.py:437 */
        fo_tukey(n, s->win);
        s->win_n = n;
    }
    for (int i = 0; i < n; i++) s->wx[i] = (double)s->x[i] * s->win[i];

    /* 2. autocorrelation, lags 0..L (encoder.py:371, :443-450) */
    double autoc[FO_MAX_LPC + 1];
    for (int lag = 0; lag <= L; lag++) {
        int cnt = n - lag - 1; /* range(len - lag - 1): the last product is dropped */
        if (cnt <= 0) {
            autoc[lag] = 0.0; /* sum(()) == 0 */
            continue;
        }
        for (int j = 0; j < cnt; j++) s->prod[j] = s->wx[j] * s->wx[j + lag];
        autoc[lag] = fo_neumaier_sum(s->prod, (uint64_t)cnt);
    }
    if (info)
        for (int lag = 0; lag <= L; lag++) info->autoc[lag] = autoc[lag];

    /* 3. Levinson-Durbin per order (encoder.py:374-375) */
    static __thread double coefs[FO_MAX_LPC][FO_MAX_LPC];
    int rc = levinson_all(autoc, L, coefs, err_line);
    if (rc) return rc;

    /* 4-6. quantise, residual and total error per order (encoder.py:378-398) */
    int best = -1;
    int64_t best_total = 0;
    for (int m = 1; m <= L; m++) {
        int32_t q[FO_MAX_LPC];
        int shift, ncoef;
        rc = quantize_coefs(coefs[m - 1], m, precision, q, &shift, &ncoef, err_line);
        if (rc) return rc;
        int len;
        if (ncoef == 0) { /* prediction_residual(samples, [], 0): every sample, order 0 */
            for (int i = 0; i < n; i++) s->res[i] = s->x[i];
            len = n;
        } else {
            len = n - m;
            if (len < 0) len = 0;
            if (len > 0) prediction_residual(s->x, n, q, m, shift, s->res);
        }
        int64_t total = sum_abs(s->res, len);
        if (info) {
            info->lpc_err[m - 1] = total;
            info->lpc_shift[m - 1] = shift;
            info->lpc_ncoef[m - 1] = ncoef;
            for (int j = 0; j < ncoef; j++) info->lpc_q[m - 1][j] = q[j];
        }
        if (best < 0 || total < best_total) { /* first minimum, encoder.py:404 */
            best = m;
            best_total = total;
            out->order = m;
            out->ncoef = ncoef;
            out->shift = shift;
            memcpy(out->q, q, sizeof(int32_t) * (size_t)ncoef);
            out->res_len = len;
            out->total = total;
            memcpy(res_out, s->res, sizeof(int64_t) * (size_t)len);
        }
    }
    return FO_OK;
}

/* ------------------------------------------------------------------------- */
/* encoder.py:632-760 encode_residual / rice_partitions                      */

static int rice_search(scratch_t *s, const int64_t *res, int res_len, int bs, int pred_order,
                       int rice_min, int rice_max, int *porder_out, int *method_out,
                       int *err_line)
{
    for (int i = 0; i < res_len; i++) s->zz[i] = zigzag(res[i]); /* encoder.py:639 */

    int have = 0;
    uint64_t best_size = 0;
    for (int o = rice_min; o <= rice_max; o++) {
        /* encoder.py:664-667 */
        if (!(bs % (1 << o) == 0 && (bs >> o) > pred_order)) continue;
        /* encoder.py:698-711 split_samples_in_partitions */
        int p0 = (bs >> o) - pred_order;
        int ps = bs >> o;
        uint64_t total = 0;
        int pos = 0, np = 0;
        /* [samples[:p0], *group(samples[p0:], ps)]: partition 0 always exists */
        while (np == 0 || pos < res_len) {
            int len = np == 0 ? p0 : ps;
            if (len > res_len - pos) len = res_len - pos;
            /* encoder.py:730-753 find_rice_parameter */
            uint64_t sum = 0;
            for (int i = 0; i < len; i++) sum += s->zz[pos + i];
            if (len == 0) FAIL(FO_E_ZERODIV);
            double mean = (double)sum / (double)len;
            if (!(mean > 0.0)) FAIL(FO_E_VALUE); /* log2(0.0): math domain error */
            int k = (int)floor(log2(mean));
            if (k < 0) FAIL(FO_E_VALUE); /* 1 << negative, encoder.py:758 */
            /* encoder.py:714-727 */
            uint64_t size = 4 + (uint64_t)(k > 14 ? 5 : 4);
            for (int i = 0; i < len; i++)
                size += (k < 64 ? (s->zz[pos + i] >> k) : 0) + 1 + (uint64_t)k;
            total += size;
            if (np < (1 << 15)) s->params_tmp[np] = (uint8_t)(k > 255 ? 255 : k);
            np++;
            pos += len;
            if (len == 0) break;
        }
        if (!have || total < best_size) { /* first minimum, encoder.py:688 */
            have = 1;
            best_size = total;
            *porder_out = 0;
            { /* Residual.partition_order = log2i(len(partitions)), common.py:410-412 */
                int x = np, r = 0;
                while (x > 0) {
                    x >>= 1;
                    r++;
                }
                *porder_out = r - 1;
            }
            memcpy(s->params, s->params_tmp, (size_t)np);
            int m = 4;
            for (int i = 0; i < np; i++)
                if (s->params_tmp[i] > 14) m = 5; /* encoder.py:648-650 */
            *method_out = m;
        }
    }
    if (!have) FAIL(FO_E_ASSERT); /* encoder.py:669 */
    return FO_OK;
}

/* encoder.py:765-806 put_residual */
static void put_residual(put_t *put, scratch_t *s, int res_len, int bs, int pred_order, int porder,
                         int method)
{
    put_uint(put, method == 4 ? 0 : 1, 2); /* encoder.py:777-782 */
    put_uint(put, (uint64_t)porder, 4);
    int nparts = 1 << porder;
    int pos = 0;
    for (int p = 0; p < nparts; p++) {
        int len = p == 0 ? (bs >> porder) - pred_order : (bs >> porder);
        if (len > res_len - pos) len = res_len - pos;
        int k = s->params[p];
        put_uint(put, (uint64_t)k, method); /* encoder.py:791 */
        for (int i = 0; i < len; i++) {     /* encoder.py:798-806 */
            uint64_t u = s->zz[pos + i];
            put_zeros(put, u >> k);
            put_uint(put, 1, 1);
            put_uint(put, u, k); /* k LSBs, MSB first */
        }
        pos += len;
    }
}

/* ------------------------------------------------------------------------- */
/* encoder.py:194-234 put_frame_header                                       */

static int block_size_code(int bs, int *extra_bits)
{
    /* common.py:85-105, encoder.py:245-255 */
    static const int tab[][2] = {{192, 1},   {576, 2},   {1152, 3},  {2304, 4},  {4608, 5},
                                 {256, 8},   {512, 9},   {1024, 10}, {2048, 11}, {4096, 12},
                                 {8192, 13}, {16384, 14}, {32768, 15}};
    *extra_bits = 0;
    for (unsigned i = 0; i < sizeof(tab) / sizeof(tab[0]); i++)
        if (tab[i][0] == bs) return tab[i][1];
    int n = bit_length_u64((uint64_t)bs);
    if (n > 0 && n <= 8) {
        *extra_bits = 8;
        return 6;
    }
    if (n > 8 && n <= 16) {
        *extra_bits = 16;
        return 7;
    }
    return -1;
}

static int put_frame_header(put_t *put, int bs, int channels, uint64_t number, uint32_t flags,
                            int *err_line)
{
    int extra;
    int code = block_size_code(bs, &extra);
    if (code < 0) FAIL(FO_E_VALUE); /* encoder.py:255 */
    put_uint(put, 0x7FFC, 15);      /* common.py:10 */
    put_uint(put, 0, 1);            /* BlockingStrategy.Fixed */
    put_uint(put, (uint64_t)code, 4);
    put_uint(put, 0, 4); /* sample rate: from STREAMINFO (encoder.py:94) */
    /* encoder.py:95: Channels.L_R whatever the channel count */
    put_uint(put, (flags & FO_FLAG_SPEC_CHANNELS) ? (uint64_t)(channels - 1) : 1, 4);
    put_uint(put, 0, 3); /* sample size: from STREAMINFO (encoder.py:96) */
    put_uint(put, 0, 1);
    uint8_t cn[7];
    int len = fo_coded_number(number, cn);
    if (len == -1) FAIL(FO_E_ASSERT);
    if (len < 0) FAIL(FO_E_VALUE);
    for (int i = 0; i < len; i++) put_uint(put, cn[i], 8);
    if (extra) put_uint(put, (uint64_t)(bs - 1), extra);
    put_uint(put, fo_crc8(put->buf, put->nbits >> 3), 8); /* encoder.py:232 */
    return FO_OK;
}

/* ------------------------------------------------------------------------- */
/* encoder.py:87-165: one frame                                              */

static int encode_frame(scratch_t *s, const int32_t *pcm /* first sample frame of the block */,
                        int bs, int channels, int bps, int L, int precision, int rice_min,
                        int rice_max, uint64_t number, uint32_t flags, uint8_t *out,
                        uint64_t cap, uint64_t *len_out, fo_subframe_info *info, int *err_line)
{
    put_t put;
    put_init(&put, out, cap);
    int rc = put_frame_header(&put, bs, channels, number, flags, err_line);
    if (rc) return rc;

    for (int c = 0; c < channels; c++) {
        fo_subframe_info *inf = info ? &info[c] : NULL;
        if (inf) memset(inf, 0, sizeof(*inf));
        for (int i = 0; i < bs; i++) s->x[i] = pcm[(size_t)i * (size_t)channels + (size_t)c];

        int f_order, f_len;
        int64_t f_total, err5[5];
        subframe_fixed(s, bs, &f_order, s->fixed_res, &f_len, &f_total, err5);
        if (inf) {
            for (int o = 0; o < 5; o++) inf->fixed_err[o] = err5[o];
            inf->fixed_order = f_order;
        }

        int use_lpc = 0;
        lpc_result_t lpc;
        memset(&lpc, 0, sizeof(lpc));
        if (L > 0) {
            rc = subframe_lpc(s, bs, L, precision, &lpc, s->lpc_res, inf, err_line);
            if (rc) return rc;
            if (inf) inf->lpc_order = lpc.order;
            if (f_total < lpc.total) use_lpc = 0; /* encoder.py:138 */
            else if (lpc.total < f_total) use_lpc = 1;
            else FAIL(FO_E_ASSERT); /* encoder.py:157 */
        }

        uint64_t bits0 = put.nbits;
        int porder = 0, method = 4;
        if (!use_lpc) {
            /* encoder.py:553-569, 581-598 */
            put_uint(&put, 0, 1);
            put_uint(&put, (uint64_t)(0x08 | f_order), 6);
            put_uint(&put, 0, 1);
            rc = rice_search(s, s->fixed_res, f_len, bs, f_order, rice_min, rice_max, &porder,
                             &method, err_line);
            if (rc) return rc;
            for (int i = 0; i < f_order; i++) put_uint(&put, (uint64_t)s->x[i], bps);
            put_residual(&put, s, f_len, bs, f_order, porder, method);
        } else {
            /* SubframeLPC.order = len(warmup), warmup = samples[:order] (encoder.py:408) */
            int warm = lpc.order < bs ? lpc.order : bs;
            put_uint(&put, 0, 1);
            put_uint(&put, (uint64_t)(0x20 | (lpc.order - 1)), 6);
            put_uint(&put, 0, 1);
            rc = rice_search(s, s->lpc_res, lpc.res_len, bs, warm, rice_min, rice_max, &porder,
                             &method, err_line);
            if (rc) return rc;
            for (int i = 0; i < warm; i++) put_uint(&put, (uint64_t)s->x[i], bps);
            if (precision - 1 == 15) FAIL(FO_E_ASSERT); /* encoder.py:619 */
            put_uint(&put, (uint64_t)(precision - 1), 4);
            put_uint(&put, (uint64_t)lpc.shift, 5);
            for (int i = 0; i < lpc.ncoef; i++) put_uint(&put, (uint64_t)(int64_t)lpc.q[i], precision);
            put_residual(&put, s, lpc.res_len, bs, warm, porder, method);
        }
        if (inf) {
            inf->type = use_lpc;
            inf->order = use_lpc ? lpc.order : f_order;
            inf->shift = use_lpc ? lpc.shift : 0;
            inf->precision = use_lpc ? precision : 0;
            inf->porder = porder;
            inf->method = method;
            inf->nparams = 1 << porder;
            for (int i = 0; i < (1 << porder) && i < FO_INFO_PARAMS; i++) inf->params[i] = s->params[i];
            if (use_lpc)
                for (int i = 0; i < lpc.ncoef; i++) inf->qcoef[i] = lpc.q[i];
            inf->bits = (int64_t)(put.nbits - bits0);
        }
        if (put.overflow) FAIL(FO_E_NOSPACE);
    }
    /* encoder.py:159-163 */
    put_uint(&put, 0, (int)((8 - (put.nbits & 7)) & 7));
    if (put.overflow || (put.nbits >> 3) + 2 > cap) FAIL(FO_E_NOSPACE);
    uint16_t crc = fo_crc16(out, put.nbits >> 3);
    put_uint(&put, crc, 16);
    if (put.overflow) FAIL(FO_E_NOSPACE);
    *len_out = put.nbits >> 3;
    return FO_OK;
}

static uint64_t frame_bound(int bs, int channels, int bps, int rice_max)
{
    /* Rice bits < N*(bps+3.01) + parameters + warm-up + coefficients */
    uint64_t per = ((uint64_t)bs * (uint64_t)(bps + 4) + 7) / 8 + ((uint64_t)1 << rice_max) + 256;
    return per * (uint64_t)channels + 64;
}

int fo_encode_frames(const int32_t *pcm, uint64_t n_sample_frames, int channels, int bps,
                     int block_size, int max_lpc_order, int qlp_precision, int rice_min,
                     int rice_max, uint64_t first_frame_number, uint32_t flags, uint8_t *out,
                     uint64_t out_cap, uint64_t *out_len, uint64_t *frame_offsets,
                     fo_subframe_info *info, int64_t *err_frame, int *err_line)
{
    int dummy_line;
    if (!err_line) err_line = &dummy_line;
    if (err_frame) *err_frame = -1;
    if (block_size < 1 || channels < 1 || max_lpc_order < 0 || max_lpc_order > FO_MAX_LPC ||
        rice_min < 0 || rice_max > 15 || rice_min > rice_max)
        FAIL(FO_E_ASSERT);
    scratch_t s;
    if (!scratch_init(&s, block_size, 1 << 15)) {
        scratch_free(&s);
        FAIL(FO_E_NOSPACE);
    }
    uint64_t n_frames = (n_sample_frames + (uint64_t)block_size - 1) / (uint64_t)block_size;
    uint64_t pos = 0;
    uint64_t bound = frame_bound(block_size, channels, bps, rice_max);
    uint8_t *tmp = malloc(bound);
    int rc = FO_OK;
    for (uint64_t f = 0; f < n_frames; f++) {
        uint64_t start = f * (uint64_t)block_size;
        int bs = (int)(n_sample_frames - start < (uint64_t)block_size ? n_sample_frames - start
                                                                       : (uint64_t)block_size);
        uint64_t flen = 0;
        rc = encode_frame(&s, pcm + start * (uint64_t)channels, bs, channels, bps, max_lpc_order,
                          qlp_precision, rice_min, rice_max, first_frame_number + f, flags, tmp,
                          bound, &flen, info ? info + f * (uint64_t)channels : NULL, err_line);
        if (rc) {
            if (err_frame) *err_frame = (int64_t)f;
            break;
        }
        if (pos + flen > out_cap) {
            rc = FO_E_NOSPACE;
            *err_line = __LINE__;
            if (err_frame) *err_frame = (int64_t)f;
            break;
        }
        memcpy(out + pos, tmp, flen);
        if (frame_offsets) frame_offsets[f] = pos;
        pos += flen;
    }
    if (rc == FO_OK && frame_offsets) frame_offsets[n_frames] = pos;
    if (out_len) *out_len = pos;
    free(tmp);
    scratch_free(&s);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* multi-threaded fan-out (the reference has none; used for the CPU baseline) */

typedef struct {
    const int32_t *pcm;
    uint64_t n_sample_frames;
    int channels, bps, block_size, L, precision, rice_min, rice_max;
    uint64_t first_number;
    uint32_t flags;
    uint64_t f0, f1; /* frame range */
    uint8_t *buf;
    uint64_t cap, len;
    uint64_t *offs; /* per-frame local offsets, f1 - f0 + 1 */
    int rc, line;
    int64_t err_frame;
} mt_job_t;

static void *mt_worker(void *arg)
{
    mt_job_t *j = arg;
    uint64_t s0 = j->f0 * (uint64_t)j->block_size;
    uint64_t s1 = j->f1 * (uint64_t)j->block_size;
    if (s1 > j->n_sample_frames) s1 = j->n_sample_frames;
    j->rc = fo_encode_frames(j->pcm + s0 * (uint64_t)j->channels, s1 - s0, j->channels, j->bps,
                             j->block_size, j->L, j->precision, j->rice_min, j->rice_max,
                             j->first_number + j->f0, j->flags, j->buf, j->cap, &j->len, j->offs,
                             NULL, &j->err_frame, &j->line);
    return NULL;
}

int fo_encode_frames_mt(const int32_t *pcm, uint64_t n_sample_frames, int channels, int bps,
                        int block_size, int max_lpc_order, int qlp_precision, int rice_min,
                        int rice_max, uint64_t first_frame_number, uint32_t flags, uint8_t *out,
                        uint64_t out_cap, uint64_t *out_len, uint64_t *frame_offsets, int nthreads,
                        int64_t *err_frame, int *err_line)
{
    uint64_t n_frames = (n_sample_frames + (uint64_t)block_size - 1) / (uint64_t)block_size;
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n_frames) nthreads = n_frames ? (int)n_frames : 1;
    mt_job_t *jobs = calloc((size_t)nthreads, sizeof(mt_job_t));
    pthread_t *th = calloc((size_t)nthreads, sizeof(pthread_t));
    uint64_t bound = frame_bound(block_size, channels, bps, rice_max);
    for (int t = 0; t < nthreads; t++) {
        mt_job_t *j = &jobs[t];
        j->pcm = pcm;
        j->n_sample_frames = n_sample_frames;
        j->channels = channels;
        j->bps = bps;
        j->block_size = block_size;
        j->L = max_lpc_order;
        j->precision = qlp_precision;
        j->rice_min = rice_min;
        j->rice_max = rice_max;
        j->first_number = first_frame_number;
        j->flags = flags;
        j->f0 = n_frames * (uint64_t)t / (uint64_t)nthreads;
        j->f1 = n_frames * (uint64_t)(t + 1) / (uint64_t)nthreads;
        j->cap = bound * (j->f1 - j->f0) + 16;
        j->buf = malloc(j->cap);
        j->offs = malloc(sizeof(uint64_t) * (j->f1 - j->f0 + 1));
        pthread_create(&th[t], NULL, mt_worker, j);
    }
    int rc = FO_OK;
    uint64_t pos = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        mt_job_t *j = &jobs[t];
        if (rc == FO_OK && j->rc != FO_OK) {
            rc = j->rc;
            if (err_line) *err_line = j->line;
            if (err_frame) *err_frame = (int64_t)j->f0 + j->err_frame;
        }
        if (rc == FO_OK) {
            if (pos + j->len > out_cap) {
                rc = FO_E_NOSPACE;
                if (err_line) *err_line = __LINE__;
            } else {
                memcpy(out + pos, j->buf, j->len);
                if (frame_offsets)
                    for (uint64_t f = j->f0; f < j->f1; f++) frame_offsets[f] = pos + j->offs[f - j->f0];
                pos += j->len;
            }
        }
        free(j->buf);
        free(j->offs);
    }
    if (rc == FO_OK && frame_offsets) frame_offsets[n_frames] = pos;
    if (out_len) *out_len = pos;
    free(jobs);
    free(th);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* encoder.py:61-83, :170-189                                                */

int fo_stream_prologue(uint32_t sample_rate, int channels, int bps, uint64_t total_samples,
                       int block_size, uint8_t out[42])
{
    put_t put;
    put_init(&put, out, 42);
    put_uint(&put, 0x664C6143, 32); /* "fLaC", common.py:9 */
    put_uint(&put, 1, 1);           /* last */
    put_uint(&put, 0, 7);           /* STREAMINFO */
    put_uint(&put, 34, 24);
    put_uint(&put, (uint64_t)block_size, 16);
    put_uint(&put, (uint64_t)block_size, 16);
    put_uint(&put, 0, 24);
    put_uint(&put, 0, 24);
    put_uint(&put, sample_rate, 20);
    put_uint(&put, (uint64_t)(channels - 1), 3);
    put_uint(&put, (uint64_t)(bps - 1), 5);
    put_uint(&put, total_samples, 36);
    for (int i = 0; i < 16; i++) put_uint(&put, 0, 8); /* md5 = bytes(16) */
    return put.overflow ? FO_E_NOSPACE : FO_OK;
}

/* ------------------------------------------------------------------------- */
/* binary.py:78-141 Get, decoder.py                                          */

typedef struct {
    const uint8_t *buf;
    uint64_t len;  /* bytes */
    uint64_t pos;  /* bits */
    int eof;
} get_t;

static uint64_t get_uint(get_t *g, int n)
{
    uint64_t r = 0;
    for (int i = 0; i < n; i++) {
        uint64_t byte = g->pos >> 3;
        if (byte >= g->len) {
            g->eof = 1; /* binary.py:40 EOFError */
            return 0;
        }
        r = (r << 1) | ((g->buf[byte] >> (7 - (g->pos & 7))) & 1);
        g->pos++;
    }
    return r;
}

/* Get.sint (binary.py:129-131); n == 0 is a negative shift count -> ValueError */
static int64_t get_sint(get_t *g, int n)
{
    uint64_t x = get_uint(g, n);
    return (int64_t)x - (int64_t)((x >> (n - 1)) << n);
}

#define DFAIL(code)                \
    do {                           \
        if (err_line) *err_line = __LINE__; \
        return (code);             \
    } while (0)
#define CHECK_EOF()                \
    do {                           \
        if (g->eof) return FO_E_EOF; \
    } while (0)

static int dec_residual(get_t *g, int bs, int pred_order, int64_t *res, int *err_line)
{
    /* decoder.py:358-421 */
    int method;
    uint64_t m = get_uint(g, 2);
    CHECK_EOF();
    if (m == 0) method = 4;
    else if (m == 1) method = 5;
    else DFAIL(FO_E_VALUE);
    int porder = (int)get_uint(g, 4);
    CHECK_EOF();
    int nparts = 1 << porder;
    if (bs % nparts != 0) DFAIL(FO_E_ASSERT);
    if (!((bs >> porder) > pred_order)) DFAIL(FO_E_ASSERT);
    int pos = 0;
    for (int p = 0; p < nparts; p++) {
        int cnt = p == 0 ? (bs >> porder) - pred_order : (bs >> porder);
        int k = (int)get_uint(g, method);
        CHECK_EOF();
        if (k == (1 << method) - 1) { /* escaped, decoder.py:404-407 */
            int size = (int)get_uint(g, 5);
            CHECK_EOF();
            for (int i = 0; i < cnt; i++) {
                if (size == 0) DFAIL(FO_E_VALUE);
                res[pos++] = get_sint(g, size);
                CHECK_EOF();
            }
        } else {
            for (int i = 0; i < cnt; i++) { /* decoder.py:414-421 */
                uint64_t msb = 0;
                for (;;) {
                    uint64_t b = get_uint(g, 1);
                    CHECK_EOF();
                    if (b) break;
                    msb++;
                }
                uint64_t lsb = get_uint(g, k);
                CHECK_EOF();
                uint64_t u = (msb << k) | lsb;
                res[pos++] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1); /* utils.py:87-88 */
            }
        }
    }
    return FO_OK;
}

/* decoder.py:493-498 */
static void dec_predict(const int32_t *coef, int order, int shift, int64_t *r, int n)
{
    for (int i = order; i < n; i++) {
        int64_t s = 0;
        for (int j = 0; j < order; j++) s += r[i - 1 - j] * (int64_t)coef[j];
        r[i] += s >> shift;
    }
}

static int dec_subframe(get_t *g, int bs, int sample_size, int64_t *out, int *err_line)
{
    /* decoder.py:318-355 */
    if (get_uint(g, 1) != 0) {
        CHECK_EOF();
        DFAIL(FO_E_ASSERT);
    }
    CHECK_EOF();
    int type = (int)get_uint(g, 6);
    CHECK_EOF();
    if (!(type <= 1 || (type >= 8 && type <= 12) || type >= 32)) DFAIL(FO_E_ASSERT);
    int wasted = 0;
    if (get_uint(g, 1)) { /* decoder.py:346-355 (counts zeros; off by one vs the format) */
        CHECK_EOF();
        for (;;) {
            uint64_t b = get_uint(g, 1);
            CHECK_EOF();
            if (b) break;
            wasted++;
        }
    }
    CHECK_EOF();
    int ss = sample_size - wasted;
    if (type == 0) { /* constant */
        if (ss <= 0) DFAIL(FO_E_VALUE);
        int64_t v = get_sint(g, ss);
        CHECK_EOF();
        for (int i = 0; i < bs; i++) out[i] = v;
        return FO_OK;
    }
    if (type == 1) { /* verbatim */
        for (int i = 0; i < bs; i++) {
            if (ss <= 0) DFAIL(FO_E_VALUE);
            out[i] = get_sint(g, ss);
            CHECK_EOF();
        }
        return FO_OK;
    }
    if (type >= 8 && type <= 12) { /* fixed */
        int order = type & 7;
        for (int i = 0; i < order; i++) {
            if (ss <= 0) DFAIL(FO_E_VALUE);
            out[i] = get_sint(g, ss);
            CHECK_EOF();
        }
        int rc = dec_residual(g, bs, order, out + order, err_line);
        if (rc) return rc;
        dec_predict(FIXED_COEF[order], order, 0, out, bs);
        return FO_OK;
    }
    /* lpc */
    int order = (type & 31) + 1;
    for (int i = 0; i < order; i++) {
        if (ss <= 0) DFAIL(FO_E_VALUE);
        out[i] = get_sint(g, ss);
        CHECK_EOF();
    }
    int precision = (int)get_uint(g, 4);
    CHECK_EOF();
    if (precision == 15) DFAIL(FO_E_ASSERT);
    precision += 1;
    int shift = (int)get_sint(g, 5);
    CHECK_EOF();
    int32_t coef[FO_MAX_LPC];
    for (int i = 0; i < order; i++) {
        coef[i] = (int32_t)get_sint(g, precision);
        CHECK_EOF();
    }
    int rc = dec_residual(g, bs, order, out + order, err_line);
    if (rc) return rc;
    if (shift < 0) DFAIL(FO_E_VALUE); /* >> negative */
    dec_predict(coef, order, shift, out, bs);
    return FO_OK;
}

int fo_decode_stream(const uint8_t *stream, uint64_t len, uint32_t *sample_rate, int *bps,
                     int *channels, uint64_t *total_samples, int32_t *pcm_out,
                     uint64_t cap_samples, uint64_t *n_out, int *frame_channels, int *err_line)
{
    get_t gg = {stream, len, 0, 0};
    get_t *g = &gg;
    int dummy;
    if (!err_line) err_line = &dummy;
    if (n_out) *n_out = 0;
    /* decoder.py:34-44 */
    if (len < 4) return FO_E_EOF;
    if (memcmp(stream, "fLaC", 4) != 0) DFAIL(FO_E_ASSERT);
    g->pos = 32;
    int last = (int)get_uint(g, 1);
    int type = (int)get_uint(g, 7);
    (void)get_uint(g, 24);
    CHECK_EOF();
    if (type > 6) DFAIL(FO_E_VALUE); /* MetadataBlockType(x) */
    if (type != 0) DFAIL(FO_E_ASSERT);
    (void)get_uint(g, 16);
    (void)get_uint(g, 16);
    (void)get_uint(g, 24);
    (void)get_uint(g, 24);
    uint32_t rate = (uint32_t)get_uint(g, 20);
    int si_channels = (int)get_uint(g, 3) + 1;
    int si_bps = (int)get_uint(g, 5) + 1;
    uint64_t si_total = get_uint(g, 36);
    for (int i = 0; i < 16; i++) (void)get_uint(g, 8);
    CHECK_EOF();
    while (!last) { /* decoder.py:92-97 skip_metadata */
        last = (int)get_uint(g, 1);
        int t = (int)get_uint(g, 7);
        uint64_t l = get_uint(g, 24);
        CHECK_EOF();
        if (t > 6) DFAIL(FO_E_VALUE);
        if (l > 0) {
            if ((g->pos >> 3) + l > g->len) return FO_E_EOF;
            g->pos += l * 8;
        } else {
            DFAIL(FO_E_VALUE); /* read(buffer, 0): "n must be greater than zero" */
        }
    }
    if (sample_rate) *sample_rate = rate;
    if (bps) *bps = si_bps;
    if (channels) *channels = si_channels;
    if (total_samples) *total_samples = si_total;

    uint64_t produced = 0;
    int fch = 0;
    int64_t *chan[8] = {0};
    int chan_alloc = 0;
    int rc = FO_OK;
    for (;;) { /* decoder.py:100-108 get_frames */
        /* decoder.py:133-185 get_frame_header */
        uint64_t sync = get_uint(g, 15);
        if (g->eof) break;
        if (sync != 0x7FFC) {
            rc = FO_E_ASSERT;
            *err_line = __LINE__;
            break;
        }
        (void)get_uint(g, 1);
        int bsc = (int)get_uint(g, 4);
        if (g->eof) break;
        if (!(bsc > 0 && bsc < 15)) {
            rc = FO_E_ASSERT;
            *err_line = __LINE__;
            break;
        }
        int src = (int)get_uint(g, 4);
        if (g->eof) break;
        if (!(src < 15)) {
            rc = FO_E_ASSERT;
            *err_line = __LINE__;
            break;
        }
        if (src == 11) { /* SAMPLE_RATE_VALUE_DECODING[0b1011] */
            rc = FO_E_KEY;
            *err_line = __LINE__;
            break;
        }
        int chc = (int)get_uint(g, 4);
        if (g->eof) break;
        if (!(chc <= 10)) {
            rc = FO_E_ASSERT;
            *err_line = __LINE__;
            break;
        }
        int ssc = (int)get_uint(g, 3);
        if (g->eof) break;
        if (ssc == 3) {
            rc = FO_E_ASSERT;
            *err_line = __LINE__;
            break;
        }
        if (get_uint(g, 1) != 0 && !g->eof) {
            rc = FO_E_ASSERT;
            *err_line = __LINE__;
            break;
        }
        if (g->eof) break;
        /* coded number: decoder.py:251-262 */
        int b0 = (int)get_uint(g, 8);
        if (g->eof) break;
        int follow = b0 >= 0xFE ? 6 : b0 >= 0xFC ? 5 : b0 >= 0xF8 ? 4 : b0 >= 0xF0 ? 3
                     : b0 >= 0xE0 ? 2 : b0 >= 0xC0 ? 1 : 0;
        for (int i = 0; i < follow; i++) (void)get_uint(g, 8);
        if (g->eof) break;
        int bs;
        if (bsc == 1) bs = 192;
        else if (bsc >= 2 && bsc <= 5) bs = 144 << bsc;
        else if (bsc == 6) bs = (int)get_uint(g, 8) + 1;
        else if (bsc == 7) bs = (int)get_uint(g, 16) + 1;
        else bs = 1 << bsc;
        if (g->eof) break;
        if (src == 12) (void)get_uint(g, 8);
        else if (src == 13 || src == 14) (void)get_uint(g, 16);
        if (g->eof) break;
        (void)get_uint(g, 8); /* crc8, never verified (decoder.py:175) */
        if (g->eof) break;

        static const int ss_tab[8] = {0, 8, 12, 0, 16, 20, 24, 32};
        int sample_size = ssc == 0 ? si_bps : ss_tab[ssc];
        static const int ch_count[11] = {1, 2, 3, 4, 5, 6, 7, 8, 2, 2, 2};
        int nch = ch_count[chc];
        fch = nch;
        if (bs > chan_alloc) {
            for (int c = 0; c < 8; c++) {
                free(chan[c]);
                chan[c] = malloc(sizeof(int64_t) * (size_t)bs);
            }
            chan_alloc = bs;
        }
        int frc = FO_OK;
        for (int c = 0; c < nch; c++) {
            /* Channels.decorrelation_bit, common.py:182-196 */
            int extra = (chc == 8 && c == 1) || (chc == 9 && c == 0) || (chc == 10 && c == 1);
            frc = dec_subframe(g, bs, sample_size + extra, chan[c], err_line);
            if (frc) break;
        }
        if (frc == FO_E_EOF || g->eof) break;
        if (frc) {
            rc = frc;
            break;
        }
        if (g->pos & 7) { /* decoder.py:124-126 */
            uint64_t pad = get_uint(g, (int)(8 - (g->pos & 7)));
            if (g->eof) break;
            if (pad != 0) {
                rc = FO_E_ASSERT;
                *err_line = __LINE__;
                break;
            }
        }
        (void)get_uint(g, 16); /* crc16, never verified (decoder.py:128) */
        if (g->eof) break;

        /* decoder.py:431-451 decode_frame */
        if (chc == 8) { /* L_S */
            for (int i = 0; i < bs; i++) chan[1][i] = chan[0][i] - chan[1][i];
        } else if (chc == 9) { /* S_R */
            for (int i = 0; i < bs; i++) chan[0][i] = chan[0][i] + chan[1][i];
        } else if (chc == 10) { /* M_S, as the reference computes it */
            for (int i = 0; i < bs; i++) {
                int64_t c0 = chan[0][i], c1 = chan[1][i];
                int64_t right = c0 - (c1 >> 1);
                int64_t left = right + c1;
                chan[0][i] = left;
                chan[1][i] = right;
            }
        }
        if (pcm_out) {
            if ((produced + (uint64_t)bs) * (uint64_t)nch > cap_samples) {
                rc = FO_E_NOSPACE;
                *err_line = __LINE__;
                break;
            }
            for (int i = 0; i < bs; i++)
                for (int c = 0; c < nch; c++)
                    pcm_out[(produced + (uint64_t)i) * (uint64_t)nch + (uint64_t)c] = (int32_t)chan[c][i];
        }
        produced += (uint64_t)bs;
    }
    for (int c = 0; c < 8; c++) free(chan[c]);
    if (n_out) *n_out = produced;
    if (frame_channels) *frame_channels = fch;
    return rc;
}
