/*
 * cpu_sim.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal SIMT simulator so that the device code of flac-py_b200/csrc/flacb200_kernels.cuh
 * can be executed and checked against the golden vectors on a machine without a GPU.  One CTA
 * runs at a time; every CUDA thread is a ucontext fiber; __syncthreads and the warp intrinsics
 * are rendezvous points at which a fiber yields until its peers arrive.  CTAs run in launch
 * order, so the look-back always finds a published prefix.  Nothing in the product links this.
 */
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)
#define __restrict__

struct sim_dim3 { unsigned x, y, z; };
struct double2 { double x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }

namespace sim {

struct Fiber {
    ucontext_t ctx;
    char *stack;
    bool done;
};

struct State {
    ucontext_t main_ctx;
    std::vector<Fiber> fibers;
    int cur = 0;
    int nthreads = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    int nbar_count[16] = {0};
    unsigned nbar_gen[16] = {0};
    int wbar_count[64];
    unsigned wbar_gen[64];
    unsigned long long wbuf[64][32];
    unsigned char *dyn_smem = nullptr, *dyn_alloc = nullptr;
    unsigned long long block_serial = 0; /* CTAs run so far */
    std::function<void()> body;
};

inline State &st()
{
    static State s;
    return s;
}

extern sim_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

inline void yield()
{
    State &s = st();
    swapcontext(&s.fibers[(size_t)s.cur].ctx, &s.main_ctx);
}

inline void syncthreads()
{
    State &s = st();
    const unsigned g = s.bar_gen;
    if (++s.bar_count == s.nthreads) {
        s.bar_count = 0;
        s.bar_gen++;
    } else {
        while (s.bar_gen == g) yield();
    }
}

/* bar.sync id, count */
inline void named_barrier(int id, int count)
{
    State &s = st();
    if (id == 0) return syncthreads();
    const unsigned g = s.nbar_gen[id];
    if (++s.nbar_count[id] == count) {
        s.nbar_count[id] = 0;
        s.nbar_gen[id]++;
    } else {
        while (s.nbar_gen[id] == g) yield();
    }
}

inline void syncwarp()
{
    State &s = st();
    const int w = s.cur >> 5;
    const unsigned g = s.wbar_gen[w];
    if (++s.wbar_count[w] == 32) {
        s.wbar_count[w] = 0;
        s.wbar_gen[w]++;
    } else {
        while (s.wbar_gen[w] == g) yield();
    }
}

template <class T>
inline T exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "exchange of <= 8 bytes");
    State &s = st();
    const int w = s.cur >> 5, lane = s.cur & 31;
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    s.wbuf[w][lane] = raw;
    syncwarp();
    T r = v;
    if (src_lane >= 0 && src_lane < 32) memcpy(&r, &s.wbuf[w][src_lane], sizeof(T));
    syncwarp();
    return r;
}

static void fiber_entry()
{
    State &s = st();
    s.body();
    s.fibers[(size_t)s.cur].done = true;
    swapcontext(&s.fibers[(size_t)s.cur].ctx, &s.main_ctx);
}

/* run one CTA of `nthreads` threads */
inline void run_block(unsigned block, unsigned grid, int nthreads, size_t dyn_smem_bytes,
                      const std::function<void()> &body)
{
    State &s = st();
    const size_t stack_bytes = 256 * 1024;
    s.block_serial++;
    s.nthreads = nthreads;
    s.bar_count = 0;
    memset(s.nbar_count, 0, sizeof(s.nbar_count));
    memset(s.wbar_count, 0, sizeof(s.wbar_count));
    s.body = body;
    /* the dynamic shared memory ENDS where its allocation ends (and is 16-byte aligned like on the
     * device), so that an address sanitizer build catches any access behind the launch's window */
    const size_t smem_len = (dyn_smem_bytes + 15) / 16 * 16;
    const size_t alloc_len = (smem_len + 127) / 128 * 128 + 128;
    s.dyn_alloc = (unsigned char *)aligned_alloc(128, alloc_len);
    s.dyn_smem = s.dyn_alloc + (alloc_len - smem_len);
    memset(s.dyn_smem, 0xA5, smem_len); /* shared memory is not zero initialised */
    g_blockIdx = {block, 0, 0};
    g_gridDim = {grid, 1, 1};
    g_blockDim = {(unsigned)nthreads, 1, 1};
    s.fibers.resize((size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        Fiber &f = s.fibers[(size_t)t];
        f.stack = (char *)malloc(stack_bytes);
        f.done = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = stack_bytes;
        f.ctx.uc_link = &s.main_ctx;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    /* FLACB_SIM_ORDER: the order in which the threads of a CTA are resumed between their
     * rendezvous points -- "reverse", or "random[:seed]" (a new permutation every sweep).  A kernel
     * whose result depends on that order has a race between two barriers: the tests run the
     * golden cases under several orders and compare bytes (the pool has no compute-sanitizer). */
    std::vector<int> order((size_t)nthreads);
    for (int t = 0; t < nthreads; t++) order[(size_t)t] = t;
    const char *ord = getenv("FLACB_SIM_ORDER");
    const bool rev = ord && !strncmp(ord, "reverse", 7), rnd = ord && !strncmp(ord, "random", 6);
    unsigned long long rs = 0x9E3779B97F4A7C15ull ^ (unsigned long long)block;
    if (rnd && ord[6] == ':') rs ^= strtoull(ord + 7, nullptr, 0) * 0xD1B54A32D192ED03ull;
    if (rev)
        for (int t = 0; t < nthreads; t++) order[(size_t)t] = nthreads - 1 - t;
    int alive = nthreads;
    while (alive > 0) {
        if (rnd) {
            for (int t = nthreads - 1; t > 0; t--) { /* Fisher-Yates with a xorshift generator */
                rs ^= rs << 13, rs ^= rs >> 7, rs ^= rs << 17;
                const int j = (int)(rs % (unsigned long long)(t + 1));
                const int tmp = order[(size_t)t];
                order[(size_t)t] = order[(size_t)j], order[(size_t)j] = tmp;
            }
        }
        for (int k = 0; k < nthreads; k++) {
            const int t = order[(size_t)k];
            Fiber &f = s.fibers[(size_t)t];
            if (f.done) continue;
            s.cur = t;
            g_threadIdx = {(unsigned)t, 0, 0};
            swapcontext(&s.main_ctx, &f.ctx);
            if (f.done) alive--;
        }
    }
    for (int t = 0; t < nthreads; t++) free(s.fibers[(size_t)t].stack);
    free(s.dyn_alloc);
    s.dyn_smem = nullptr;
}

} /* namespace sim */

#define threadIdx sim::g_threadIdx
#define blockIdx sim::g_blockIdx
#define blockDim sim::g_blockDim
#define gridDim sim::g_gridDim

#define FLACB_DYN_SMEM(name) unsigned char *name = sim::st().dyn_smem
/* statically allocated shared memory of a kernel (a function-local static here) is poisoned by the first
 * thread of every CTA to get there: on the device it holds whatever the CTA before left behind */
#define FLACB_SIM_POISON(obj)                                   \
    do {                                                        \
        static unsigned long long done_ = 0;                    \
        if (done_ != sim::st().block_serial) {                  \
            done_ = sim::st().block_serial;                     \
            memset((void *)&(obj), 0xA5, sizeof(obj));          \
        }                                                       \
    } while (0)

static inline void __syncthreads() { sim::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { sim::syncwarp(); }
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) { sim::yield(); }

template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return sim::exchange(v, (sim::st().cur & 31) ^ m); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    const int lane = sim::st().cur & 31;
    return sim::exchange(v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return sim::exchange(v, src & 31); }
static inline unsigned __ballot_sync(unsigned, int pred);
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    sim::State &s = sim::st();
    const int w = s.cur >> 5, lane = s.cur & 31;
    s.wbuf[w][lane] = pred ? 1ull : 0ull;
    sim::syncwarp();
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if (s.wbuf[w][l]) r |= 1u << l;
    sim::syncwarp();
    return r;
}

template <class T> static inline T __reduce_add_sync(unsigned, T v)
{
    sim::State &s = sim::st();
    const int w = s.cur >> 5, lane = s.cur & 31;
    s.wbuf[w][lane] = (unsigned long long)(long long)v;
    sim::syncwarp();
    T r = 0;
    for (int l = 0; l < 32; l++) r = (T)(r + (T)(long long)s.wbuf[w][l]);
    sim::syncwarp();
    return r;
}
template <class T> static inline T __reduce_min_sync(unsigned, T v)
{
    sim::State &s = sim::st();
    const int w = s.cur >> 5, lane = s.cur & 31;
    s.wbuf[w][lane] = (unsigned long long)v;
    sim::syncwarp();
    T r = v;
    for (int l = 0; l < 32; l++)
        if ((T)s.wbuf[w][l] < r) r = (T)s.wbuf[w][l];
    sim::syncwarp();
    return r;
}
template <class T> static inline T __reduce_max_sync(unsigned, T v)
{
    sim::State &s = sim::st();
    const int w = s.cur >> 5, lane = s.cur & 31;
    s.wbuf[w][lane] = (unsigned long long)v;
    sim::syncwarp();
    T r = v;
    for (int l = 0; l < 32; l++)
        if ((T)s.wbuf[w][l] > r) r = (T)s.wbuf[w][l];
    sim::syncwarp();
    return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31));
}
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __clzll(unsigned long long x) { return x ? __builtin_clzll(x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline unsigned __sad(int a, int b, unsigned c) { const long long d = (long long)a - (long long)b; return c + (unsigned)(d < 0 ? -d : d); }
static inline int __double2hiint(double a) { long long v; memcpy(&v, &a, 8); return (int)(v >> 32); }
static inline double __hiloint2double(int hi, int lo) { long long v = ((long long)hi << 32) | (unsigned)lo; double d; memcpy(&d, &v, 8); return d; }
static inline long long __double2ll_rd(double a) { return (long long)floor(a); }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }

template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicXor(T *p, T v) { T o = *p; *p = o ^ v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
