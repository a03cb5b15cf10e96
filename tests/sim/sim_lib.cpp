/*
 * sim_lib.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Runs the device code of flac-py_b200/csrc/flacb200_kernels.cuh on the CPU SIMT simulator
 * (cpu_sim.h) behind an entry point shaped like flacb_encode_pcm, so that the CPU test suite
 * can check the kernel logic against the golden vectors and the oracle without a GPU.  It is
 * slow (a fiber per CUDA thread) and is never loaded by the product.
 */
#define FLACB_CPU_SIM 1
#include <stdio.h>
#include "cpu_sim.h"

namespace sim {
sim_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
}

#include "../../flac-py_b200/csrc/flacb200_host.h"
#include "../../flac-py_b200/csrc/flacb200_kernels.cuh"
#include "../../flac-py_b200/csrc/flacb200_serialise.cuh"
#include "../../flac-py_b200/csrc/flacb200_choose_mma.cuh"
#include "../../flac-py_b200/csrc/flacb200_decode.cuh"

using namespace flacb;

extern "C" int flacb_sim_encode(const void *pcm, uint64_t n_sample_frames, int channels, int bits,
                                int block_size, int max_lpc_order, int qlp_precision, int rice_min,
                                int rice_max, uint64_t first_frame_number, uint32_t flags,
                                uint8_t *out, uint64_t out_cap, uint64_t *out_len,
                                uint64_t *frame_offsets, flacb_subframe_info *info,
                                flacb_stats *stats, int threads_override, uint32_t plane_mode)
{
    EncodeArgs a = {n_sample_frames, channels, bits, block_size, max_lpc_order, qlp_precision,
                    rice_min, rice_max, first_frame_number, flags};
    int rc = validate_encode_args(a);
    if (out_len) *out_len = 0;
    if (rc) return rc;
    const uint64_t n_frames = (n_sample_frames + (uint64_t)block_size - 1) / (uint64_t)block_size;
    if (stats) memset(stats, 0, sizeof(*stats)), stats->first_bad_frame = -1;
    if (n_frames == 0) {
        if (frame_offsets) frame_offsets[0] = 0;
        return FLACB_OK;
    }

    crc16_table(c_crc16_tab);
    crc16_xpow_tables(c_xpow_lo, c_xpow_hi, XPOW_HI);

    EncParams p;
    memset(&p, 0, sizeof(p));
    p.pcm = (const uint8_t *)pcm;
    p.n_sample_frames = n_sample_frames;
    p.channels = channels;
    p.bps = bits;
    p.cbytes = (flags & FLACB_PCM_INT32) ? 4 : (bits + 7) / 8;
    p.block_size = block_size;
    p.L = max_lpc_order;
    p.P = qlp_precision;
    p.rmin = rice_min;
    p.rmax = rice_max;
    p.first_frame_number = first_frame_number;
    p.flags = flags;
    p.n_frames = (uint32_t)n_frames;
    p.tail_size = (int)(n_sample_frames - (n_frames - 1) * (uint64_t)block_size);
    std::vector<double> wf, wt;
    p.win_full_bad = !tukey_window(block_size, wf, 64);
    p.win_tail_bad = !tukey_window(p.tail_size, wt, 64);
    p.win_full = wf.data();
    p.win_tail = wt.data();
    window_flat_range(wf, block_size, p.flat_full_lo, p.flat_full_hi);
    window_flat_range(wt, p.tail_size, p.flat_tail_lo, p.flat_tail_hi);

    /* plane_mode bits 0-7: 0 = the library's own choice of shape; 1 = one plane, reloaded per channel;
     * 2 = every plane resident.  Bits 9-10: a slower, equally exact flavour of the all-orders filter. */
    auto layout_total = [&](int threads, int all_planes) {
        return (size_t)make_layout(block_size, channels, bits, rice_max, 0, threads, all_planes).total;
    };
    LaunchShape shape = pick_shape(block_size, E, 4096, 227 * 1024, 228 * 1024, layout_total);
    if (shape.threads == 0) return FLACB_E_UNSUPPORTED;
    if ((plane_mode & 0xFFu) == 1) shape.all_planes = 0;
    if ((plane_mode & 0xFFu) == 2) shape.all_planes = 1;
    const int B = threads_override ? threads_override : shape.threads;
    const SmemLayout lay = make_layout(block_size, channels, bits, rice_max, 0, B, shape.all_planes);
    const SmemLayout lay_rec = make_layout(block_size, channels, bits, rice_max, max_lpc_order, B, 0);
    p.all_planes = shape.all_planes;
    p.lay = lay;

    std::vector<uint32_t> frame_bytes(n_frames);
    std::vector<u64> frame_off(n_frames + 1, 0);
    std::vector<int32_t> frame_status(n_frames, 0);
    std::vector<u64> lb(n_frames, 0);
    uint32_t ticket = 0;
    u64 dstats[32] = {0};
    std::vector<uint8_t> slots;
    p.frame_bytes = frame_bytes.data();
    p.frame_off = frame_off.data();
    p.frame_status = frame_status.data();
    p.lb_state = lb.data();
    p.ticket = &ticket;
    p.stats = dstats;
    p.info = info;
    if (flags & FLACB_SLOT_COMPACT) {
        p.slot_stride = (frame_bound(block_size, channels, bits, rice_max) + 15) & ~15ull;
        slots.assign(p.slot_stride * n_frames, 0xEE);
        p.out = slots.data();
        p.out_cap = slots.size();
    } else {
        p.slot_stride = 0;
        p.out = out;
        p.out_cap = out_cap;
    }
    const bool staged = max_lpc_order > 0;
    std::vector<uint8_t> pre;
    std::vector<SubDecision> dec;
    if (staged) {
        /* stages A, B and C1 of the LPC pipeline, same sequence as flacb200.cu */
        p.pre_stride = (lay_rec.lpc_stride + 15u) & ~15u;
        const uint64_t n_sub = n_frames * (uint64_t)channels;
        pre.assign(n_sub * p.pre_stride, 0xA5);
        p.pre = pre.data();
        EncParams pa = p;
        const int front_planes = (plane_mode & 0xFFu) == 1 ? 0 : 1;
        pa.lay = make_layout_front(block_size, channels, bits, rice_max, max_lpc_order, B, front_planes);
        pa.all_planes = front_planes;
        for (uint64_t f = 0; f < n_frames; f++) {
            if (max_lpc_order <= 12)
                sim::run_block((unsigned)f, (unsigned)n_frames, B, pa.lay.total, [&]() { analyse_kernel<12, 576, 1, 0>(pa); });
            else
                sim::run_block((unsigned)f, (unsigned)n_frames, B, pa.lay.total, [&]() { analyse_kernel<32, 576, 1, 0>(pa); });
        }
        const int wpb = 4;
        const unsigned nb = (unsigned)((n_sub + wpb - 1) / wpb);
        EncParams pb = p;
        pb.lay = lay_rec;
        for (unsigned blk = 0; blk < nb; blk++)
            sim::run_block(blk, nb, wpb * 32, (size_t)wpb * p.pre_stride, [&]() { lpc_coefficients_kernel(pb, (uint32_t)n_sub); });
        dec.resize(n_sub);
        memset(dec.data(), 0xA5, n_sub * sizeof(SubDecision));
        p.dec = dec.data();
        EncParams pc = p;
        pc.lay = make_layout_choose(block_size, channels, bits, rice_max, max_lpc_order, B, front_planes);
        pc.all_planes = front_planes;
        int flavour = 2;
        if (bits <= 16 && qlp_precision <= 8) flavour = 0;
        else if (qlp_precision <= 8 && bits <= 24 && max_lpc_order > 16) flavour = 3; /* the library's choice for wide samples and high orders: tensor cores */
        else if (((u64)max_lpc_order << (qlp_precision - 1 + bits - 1)) <= (1ull << 30)) flavour = 1;
        {
            /* plane_mode bits 9-10: another, equally exact flavour (1, 2: scalar; 3: tensor cores, also for 16 bits);
             * bit 15: the scalar flavour the library would pick without the tensor-core path */
            const int v = (int)((plane_mode >> 9) & 3u);
            if (v == 3 && qlp_precision <= 8 && bits <= 24) flavour = 3;
            else if (v == 2 || (v == 1 && ((u64)max_lpc_order << (qlp_precision - 1 + bits - 1)) <= (1ull << 30))) flavour = v;
            if ((plane_mode & 0x8000u) && flavour == 3) flavour = ((u64)max_lpc_order << (qlp_precision - 1 + bits - 1)) <= (1ull << 30) ? 1 : 2;
        }
        const MmaLayout mlay = make_layout_choose_mma(block_size, channels, bits, rice_max, max_lpc_order, 256);
        for (uint64_t f = 0; f < n_frames && flavour == 3; f++) {
            EncParams pm = p;
            pm.lay = lay_rec;
            if (max_lpc_order <= 12)
                sim::run_block((unsigned)f, (unsigned)n_frames, 256, mlay.total, [&]() { choose_mma_kernel<12, 256, 2>(pm, mlay); });
            else
                sim::run_block((unsigned)f, (unsigned)n_frames, 256, mlay.total, [&]() { choose_mma_kernel<32, 256, 2>(pm, mlay); });
        }
        for (uint64_t f = 0; f < n_frames && flavour != 3; f++) {
            const unsigned fu = (unsigned)f, g = (unsigned)n_frames;
            if (max_lpc_order <= 12) {
                if (flavour == 0) sim::run_block(fu, g, B, pc.lay.total, [&]() { choose_kernel<12, 576, 1, 0, 0>(pc); });
                else if (flavour == 1) sim::run_block(fu, g, B, pc.lay.total, [&]() { choose_kernel<12, 576, 1, 1, 0>(pc); });
                else sim::run_block(fu, g, B, pc.lay.total, [&]() { choose_kernel<12, 576, 1, 2, 0>(pc); });
            } else {
                if (flavour == 0) sim::run_block(fu, g, B, pc.lay.total, [&]() { choose_kernel<32, 576, 1, 0, 0>(pc); });
                else if (flavour == 1) sim::run_block(fu, g, B, pc.lay.total, [&]() { choose_kernel<32, 576, 1, 1, 0>(pc); });
                else sim::run_block(fu, g, B, pc.lay.total, [&]() { choose_kernel<32, 576, 1, 2, 0>(pc); });
            }
        }
    }
    /* plane_mode bit 11: the second-generation serialiser (flacb200_serialise.cuh); bits 12-13:
     * samples per thread (0: 16, 1: 8, 2: 32, 3: 12); bit 14: a small staging window (mid-frame flushes) */
    const bool ser2 = (plane_mode & 0x800u) != 0;
    uint16_t crc4[4][256];
    std::vector<uint16_t> crc_w;
    if (ser2) {
        if (rice_max > V2_MAX_RICE || (flags & FLACB_SLOT_COMPACT)) return FLACB_E_UNSUPPORTED;
        const int rsel = (int)((plane_mode >> 12) & 3u);
        const int S = rsel == 1 ? 8 : rsel == 2 ? 32 : rsel == 3 ? 12 : 16;
        int Bt = ((block_size + S - 1) / S + 31) / 32 * 32;
        if (threads_override) Bt = threads_override;
        if (Bt > 1024 || Bt * S < block_size) return FLACB_E_UNSUPPORTED;
        uint32_t budget = 220 * 1024;
        if (plane_mode & 0x4000u) {
            const SerLayout probe = make_ser_layout(block_size, channels, bits, rice_max, Bt, S, 220 * 1024);
            budget = probe.stg + (((subframe_bound_bytes(block_size, bits, rice_max) + 3) / 4 + 32) + 8) * 4 + 64;
        }
        p.slay = make_ser_layout(block_size, channels, bits, rice_max, Bt, S, budget);
        if (p.slay.stg_words == 0) return FLACB_E_UNSUPPORTED;
        if (getenv("FLACB_SIM_TRACE")) fprintf(stderr, "ser2: Bt %d S %d stg_words %u total %u\n", Bt, S, p.slay.stg_words, p.slay.total);
        crc16_slice4(crc4);
        crc_w.resize(frame_bound(block_size, channels, bits, rice_max) + 64);
        crc16_weights(crc_w.data(), crc_w.size());
        p.crc4 = &crc4[0][0];
        p.crc_w = crc_w.data();
        if (getenv("FLACB_SIM_TRACE")) fprintf(stderr, "ser2: tables built\n");
        for (uint64_t f = 0; f < n_frames; f++) {
            const unsigned fu = (unsigned)f, g = (unsigned)n_frames;
            const size_t dyn2 = p.slay.total;
#define SER_RUN(LM, CH)                                                                                         \
    do {                                                                                                        \
        if (S == 8) sim::run_block(fu, g, Bt, dyn2, [&]() { serialise_kernel<LM, CH, 8, 1024, 1>(p); });        \
        else if (S == 16) sim::run_block(fu, g, Bt, dyn2, [&]() { serialise_kernel<LM, CH, 16, 1024, 1>(p); });  \
        else if (S == 12) sim::run_block(fu, g, Bt, dyn2, [&]() { serialise_kernel<LM, CH, 12, 1024, 1>(p); });  \
        else sim::run_block(fu, g, Bt, dyn2, [&]() { serialise_kernel<LM, CH, 32, 1024, 1>(p); });               \
    } while (0)
            if (!staged) SER_RUN(4, false);
            else if (max_lpc_order <= 12) SER_RUN(12, true);
            else SER_RUN(32, true);
#undef SER_RUN
            if (getenv("FLACB_SIM_TRACE")) fprintf(stderr, "ser2: frame %u done\n", fu);
        }
    }
    const size_t dyn = p.lay.total;
    for (uint64_t f = 0; f < n_frames && !ser2; f++) {
        if (!staged)
            sim::run_block((unsigned)f, (unsigned)n_frames, B, dyn, [&]() { encode_frames_kernel<12, 576, 1, false, 0>(p); });
        else if (max_lpc_order <= 12)
            sim::run_block((unsigned)f, (unsigned)n_frames, B, dyn, [&]() { encode_frames_kernel<12, 576, 1, true, 0>(p); });
        else
            sim::run_block((unsigned)f, (unsigned)n_frames, B, dyn, [&]() { encode_frames_kernel<32, 576, 1, true, 0>(p); });
    }
    if (flags & FLACB_SLOT_COMPACT) {
        sim::run_block(0, 1, 256, 0, [&]() { scan_frame_bytes_kernel(frame_bytes.data(), frame_off.data(), (uint32_t)n_frames); });
        for (uint64_t f = 0; f < n_frames; f++)
            sim::run_block((unsigned)f, (unsigned)n_frames, 128, 0, [&]() {
                compact_slots_kernel(slots.data(), p.slot_stride, frame_bytes.data(), frame_off.data(), out, out_cap,
                                     frame_status.data());
            });
    }
    uint64_t valid = frame_off[n_frames];
    for (uint64_t f = 0; f < n_frames; f++) {
        const int s = frame_status[f];
        const bool ood = s >= FLACB_E_ASSERT && s <= FLACB_E_ZERODIV;
        if (s != 0 && !(ood && (flags & FLACB_LENIENT))) {
            rc = s;
            valid = frame_off[f];
            if (stats) stats->first_bad_frame = (int64_t)f;
            break;
        }
    }
    if (out_len) *out_len = valid;
    if (frame_offsets) memcpy(frame_offsets, frame_off.data(), sizeof(u64) * (n_frames + 1));
    if (stats) {
        stats->frames = n_frames;
        stats->subframes = n_frames * (uint64_t)channels;
        for (int o = 0; o < 5; o++) stats->subframes_fixed[o] = dstats[o];
        stats->subframes_lpc = dstats[5];
        stats->out_of_domain = dstats[6];
        stats->bytes_in = n_sample_frames * (uint64_t)channels * (uint64_t)p.cbytes;
        stats->bytes_out = valid;
        stats->threads_per_frame = B;
        stats->smem_bytes = (int32_t)lay.total;
    }
    return rc;
}

/* the chunk plan of the host path (flacb200_host.h), for the host-logic tests */
extern "C" int flacb_sim_chunk_plan(uint64_t n_frames, uint64_t wave, uint64_t frame_pcm, long long want_override,
                                    long long edge_override, long long last_override, uint64_t *out, int cap)
{
    const std::vector<uint64_t> f0 = chunk_plan(n_frames, wave, frame_pcm, want_override, edge_override, last_override);
    for (size_t i = 0; i < f0.size() && (int)i < cap; i++) out[i] = f0[i];
    return (int)f0.size();
}

extern "C" int flacb_sim_prologue(uint32_t sample_rate, int channels, int bits, uint64_t total_samples,
                                  int block_size, uint8_t out[42])
{
    stream_prologue(sample_rate, channels, bits, total_samples, block_size, out);
    return 0;
}

/* flacb_decode_frames on the simulator (same kernel sequence as flacb200.cu) */
static int sim_decode_impl(const uint8_t *frames, uint64_t len, int streaminfo_bits, int32_t *pcm_out,
                           uint64_t cap_samples, uint64_t *n_sample_frames, int *frame_channels,
                           uint64_t *n_candidates, int check_crc8)
{
    if (n_sample_frames) *n_sample_frames = 0;
    if (frame_channels) *frame_channels = 0;
    if (n_candidates) *n_candidates = 0;
    if (len < 2) return FLACB_OK;
    const u64 n_chunks = (len + SYNC_CHUNK - 1) / SYNC_CHUNK;
    std::vector<uint32_t> counts(n_chunks);
    std::vector<u64> offsets(n_chunks + 1);
    for (u64 b = 0; b < n_chunks; b++)
        sim::run_block((unsigned)b, (unsigned)n_chunks, SYNC_THREADS, 0,
                       [&]() { find_sync_kernel(frames, len, counts.data(), offsets.data(), nullptr, 0); });
    sim::run_block(0, 1, 256, 0, [&]() { scan_frame_bytes_kernel(counts.data(), offsets.data(), (uint32_t)n_chunks); });
    const u64 n_cand = offsets[n_chunks];
    if (n_candidates) *n_candidates = n_cand;
    if (n_cand == 0) return FLACB_E_ASSERT;
    std::vector<Candidate> cand(n_cand);
    std::vector<uint32_t> sizes(n_cand), jump(2 * n_cand), reach(n_cand);
    std::vector<u64> tmpoff(n_cand + 1);
    for (u64 b = 0; b < n_chunks; b++)
        sim::run_block((unsigned)b, (unsigned)n_chunks, SYNC_THREADS, 0,
                       [&]() { find_sync_kernel(frames, len, counts.data(), offsets.data(), cand.data(), 1); });
    const unsigned nb = (unsigned)((n_cand + 127) / 128);
    for (unsigned b = 0; b < nb; b++)
        sim::run_block(b, nb, 128, 0, [&]() { parse_headers_kernel(frames, len, streaminfo_bits, cand.data(), (uint32_t)n_cand, sizes.data(), check_crc8); });
    sim::run_block(0, 1, 256, 0, [&]() { scan_frame_bytes_kernel(sizes.data(), tmpoff.data(), (uint32_t)n_cand); });
    std::vector<int32_t> tmp(tmpoff[n_cand] + 16);
    std::vector<SubRec> recs(n_cand * SUBREC_SLOTS);
    memset(recs.data(), 0, recs.size() * sizeof(SubRec));
    const u64 n_warps = (n_cand + DECODE_LANES - 1) / DECODE_LANES;
    const unsigned nbd = (unsigned)((n_warps + DECODE_WARPS - 1) / DECODE_WARPS);
    for (unsigned b = 0; b < nbd; b++)
        sim::run_block(b, nbd, 32 * DECODE_WARPS, 0, [&]() { decode_frames_kernel(frames, len, cand.data(), (uint32_t)n_cand, tmpoff.data(), tmp.data(), recs.data(), check_crc8, DECODE_LANES); });
    const unsigned nbr = (unsigned)((recs.size() + 63) / 64);
    for (unsigned b = 0; b < nbr; b++)
        sim::run_block(b, nbr, 64, 0, [&]() { restore_kernel(recs.data(), (uint32_t)n_cand, tmp.data()); });
    DecodeResult res;
    sim::run_block(0, 1, 1024, 0, [&]() { chain_kernel(cand.data(), (uint32_t)n_cand, len, jump.data(), reach.data(), &res); });
    if (res.status == STATUS_CRC8_MISMATCH || res.status == STATUS_DEFERRED) return STATUS_CRC8_MISMATCH;
    const u64 total = res.n_sample_frames * (u64)(res.frame_channels > 0 ? res.frame_channels : 1);
    if (total > cap_samples) return FLACB_E_NOSPACE;
    if (total)
        for (u64 b = 0; b < n_cand; b++)
            sim::run_block((unsigned)b, (unsigned)n_cand, 256, 0,
                           [&]() { interleave_kernel(cand.data(), tmp.data(), pcm_out, total, res.frame_channels); });
    if (n_sample_frames) *n_sample_frames = res.n_sample_frames;
    if (frame_channels) *frame_channels = res.frame_channels;
    return res.status;
}

extern "C" int flacb_sim_decode(const uint8_t *frames, uint64_t len, int streaminfo_bits, int32_t *pcm_out,
                                uint64_t cap_samples, uint64_t *n_sample_frames, int *frame_channels,
                                uint64_t *n_candidates)
{
    int rc = sim_decode_impl(frames, len, streaminfo_bits, pcm_out, cap_samples, n_sample_frames, frame_channels,
                             n_candidates, 1);
    if (rc == STATUS_CRC8_MISMATCH)
        rc = sim_decode_impl(frames, len, streaminfo_bits, pcm_out, cap_samples, n_sample_frames, frame_channels,
                             n_candidates, 0);
    return rc;
}
