"""The CUDA kernels' logic, run on the CPU SIMT simulator (tests/sim), against the golden streams.

This is what `-m "not gpu"` can say about the device code on a machine without a GPU; the real
parity tests (tests/test_gpu_*.py) run the same comparisons on the B200."""
import numpy as np
import pytest

import oracle as O
import simlib
from helpers import CASES, OOD, case_kwargs, case_pcm, case_stream, ood_pcm

FAST = ["c1_sine", "c2_white", "c2_silence", "c2_clip", "bs1000", "bs255", "bs192_r0", "q8_white", "q15_white24",
        "c3_white", "c4_mixed"]


@pytest.mark.parametrize("name", FAST)
def test_sim_encode_matches_golden(name):
    case = CASES[name]
    x = case_pcm(case)
    rc, data, offs, _, st = simlib.encode_frames(x, case["bits"], **case_kwargs(case))
    assert rc == 0
    assert data == case_stream(case)[42:]
    lens = [4, 4, 34] + [int(b - a) for a, b in zip(offs[:-1], offs[1:])]
    assert lens == case["chunk_lengths"]


@pytest.mark.parametrize("name", ["c2_white", "bs255", "c1_sine"])
def test_sim_slot_mode_and_single_plane_mode(name):
    case = CASES[name]
    x = case_pcm(case)
    ref = case_stream(case)[42:]
    rc, data, _, _, _ = simlib.encode_frames(x, case["bits"], flags=simlib.FLAG_SLOT_COMPACT, **case_kwargs(case))
    assert rc == 0 and data == ref
    # one channel plane in shared memory, reloaded per channel and phase (plane_mode=1), and
    # every plane resident (plane_mode=2)
    for mode in (1, 2):
        rc, data, _, _, st = simlib.encode_frames(x, case["bits"], plane_mode=mode, **case_kwargs(case))
        assert rc == 0 and data == ref, mode


@pytest.mark.parametrize("name", ["c2_white", "c2_clip"])
def test_sim_all_orders_filter_flavours_agree(name):
    """choose_kernel sums |residual| of every LPC order with dp2a on sample pairs by default here
    (16-bit PCM, 5-bit coefficients); plain 32-bit (plane_mode bits 9-10 = 1) and per-order
    32/64-bit multiply-adds (= 2) give the same sums, hence the same bytes"""
    case = CASES[name]
    x = case_pcm(case)
    ref = case_stream(case)[42:]
    for flavour in (1, 2):
        rc, data, _, _, _ = simlib.encode_frames(x, case["bits"], plane_mode=flavour << 9, **case_kwargs(case))
        assert rc == 0 and data == ref, flavour


SHAPES = [(3, 16, 12, 12, 2304, "mixed"), (5, 24, 8, 9, 1024, "white"), (1, 16, 12, 8, 4608, "white"),
          (6, 20, 12, 5, 576, "mixed"), (2, 12, 4, 5, 192, "white"), (4, 24, 32, 14, 4096, "white"),
          (2, 16, 32, 5, 4096, "white"), (2, 16, 32, 8, 1024, "white")]


@pytest.mark.parametrize("ch,bits,L,q,bs,cls", SHAPES)
def test_sim_other_channel_counts_widths_and_precisions_against_the_oracle(ch, bits, L, q, bs, cls):
    """3 / 5 / 6 / 4 channels, 12 / 20-bit samples, precisions 8 (coefficient bytes at full range),
    9, 12, 14 (the 32/64-bit flavour of the all-orders filter), order 32 on 16-bit samples (dp2a bodies
    up to 32 wide), short blocks"""
    import flacb200_synth as S
    x = S.pcm(bs * 3 + bs // 3 + 5, ch, bits, 48000, cls, 99 + ch + bits)
    ref = O.encode_frames(x, bits, bs, L, q, 0, 4)
    rc, data, offs, _, st = simlib.encode_frames(x, bits, bs, L, q, 0, 4)
    assert rc == 0 and data == ref.data and (offs == ref.frame_offsets).all()


def test_sim_other_thread_counts_and_int32_input():
    case = CASES["c2_white"]
    x = case_pcm(case)
    ref = case_stream(case)[42:]
    for threads in (64, 192, 320):  # several tiles per block, partly filled last tile
        rc, data, _, _, _ = simlib.encode_frames(x, 16, threads=threads, **case_kwargs(case))
        assert rc == 0 and data == ref, threads
    rc, data, _, _, _ = simlib.encode_frames(x, 16, int32_input=True, **case_kwargs(case))
    assert rc == 0 and data == ref


def test_sim_long_unary_runs_cross_the_staging_window():
    """spikes in near-silence make Rice codes hundreds of bits long; with a 256-word staging
    window (test build) a tile's codes span many windows: the packer's multi-pass path"""
    rng = np.random.default_rng(3)
    x = rng.integers(-2, 3, size=(4608 * 2, 2)).astype(np.int32)
    x[100::173] = 30000
    x[77::911, 1] = -32768
    ref = O.encode_frames(x, 16, 4608, 12, 5, 0, 0)
    for small in (False, True):
        rc, data, offs, _, _ = simlib.encode_frames(x, 16, 4608, 12, 5, 0, 0, small_window=small)
        assert rc == 0 and data == ref.data, small
    case = CASES["c2_white"]
    rc, data, _, _, _ = simlib.encode_frames(case_pcm(case), 16, small_window=True, **case_kwargs(case))
    assert rc == 0 and data == case_stream(case)[42:]


@pytest.mark.parametrize("name", sorted(OOD))
def test_sim_out_of_domain_status_and_lenient_stream(name):
    case = OOD[name]
    x = ood_pcm(case)
    kw = {k: case[k] for k in ("block_size", "max_lpc_order", "qlp_precision", "rice_min", "rice_max")}
    rc, data, offs, _, st = simlib.encode_frames(x, case["bits"], **kw)
    names = {1: "AssertionError", 2: "ValueError", 3: "ZeroDivisionError"}
    assert names.get(rc) == case["raises"]
    rc, data, offs, _, st = simlib.encode_frames(x, case["bits"], flags=simlib.FLAG_LENIENT, **kw)
    assert rc == 0 and st.out_of_domain >= 1
    stream = O.stream_prologue(case["rate"], 2, case["bits"], x.shape[0], case["block_size"]) + data
    assert (O.decode_stream(stream)[4] == x).all()


def test_sim_first_frame_number():
    case = CASES["bs192_r0"]
    x = case_pcm(case)
    kw = case_kwargs(case)
    ref = O.encode_frames(x, 16, first_frame_number=1000, **kw)
    rc, data, _, _, _ = simlib.encode_frames(x, 16, first_frame_number=1000, **kw)
    assert rc == 0 and data == ref.data


@pytest.mark.parametrize("name", ["c2_white", "many_frames", "c3_white", "bs255", "c1_sine", "c4_white"])
def test_sim_decode_matches_oracle(name):
    case = CASES[name]
    s = case_stream(case)
    rc, pcm, fch, ncand = simlib.decode_frames(s[42:], case["bits"])
    try:
        ref = O.decode_stream(s)[4]
        assert rc == 0 and pcm.shape == ref.shape and (pcm == ref).all()
        assert (pcm == case_pcm(case)).all()
    except O.OracleError as e:
        assert rc == e.code


def test_sim_decode_truncated_and_garbage():
    case = CASES["c2_white"]
    s = case_stream(case)[42:]
    rc, pcm, fch, _ = simlib.decode_frames(s[:-100], 16)
    assert rc == 0 and pcm.shape == (4608 * 3, 2)
    rc, pcm, fch, _ = simlib.decode_frames(b"\x00" * 100 + s, 16)  # no sync at the start
    assert rc == 1 and pcm.shape[0] == 0
    rc, pcm, fch, _ = simlib.decode_frames(s + b"\x12\x34\x56", 16)  # junk behind the last frame
    assert rc == 1 and pcm.shape == (14688, 2)


def test_sim_decode_ignores_crcs_like_the_reference():
    """a frame with a wrong header CRC-8 / frame CRC-16 still decodes (decoder.py:128, :175)"""
    case = CASES["c2_white"]
    stream = bytearray(case_stream(case))
    enc = O.encode_frames(case_pcm(case), 16, **case_kwargs(case))
    f1 = 42 + int(enc.frame_offsets[1])
    assert stream[f1:f1 + 2] == b"\xff\xf8"
    stream[f1 + 5] ^= 0x5A          # CRC-8 of frame 1
    stream[-1] ^= 0xFF              # CRC-16 of the last frame
    ref = O.decode_stream(bytes(stream))[4]
    rc, pcm, fch, _ = simlib.decode_frames(bytes(stream[42:]), 16)
    assert rc == 0 and pcm.shape == ref.shape and (pcm == ref).all()


def test_sim_decode_variable_block_sizes_take_the_unfiltered_path():
    """frames larger than frame 0's are deferred by the fast path; the decode then repeats
    unfiltered and matches the serial decoder"""
    import flacb200_synth as S
    a = S.pcm(1000 * 2, 2, 16, 44100, "white", 21)
    b = S.pcm(4096 + 500, 2, 16, 44100, "sine", 22)
    ea = O.encode_frames(a, 16, 1000, 8, 6, 0, 3)
    eb = O.encode_frames(b, 16, 4096, 8, 6, 0, 3)
    frames = ea.data + eb.data
    stream = O.stream_prologue(44100, 2, 16, a.shape[0] + b.shape[0], 4096) + frames
    ref = O.decode_stream(stream)[4]
    rc, pcm, fch, _ = simlib.decode_frames(frames, 16)
    assert rc == 0 and pcm.shape == ref.shape and (pcm == ref).all()
    assert (pcm == np.concatenate([a, b])).all()


@pytest.mark.parametrize("name", sorted(__import__("helpers").DECODE_CASES))
def test_sim_decode_crafted_streams_like_the_reference(name):
    """constant / verbatim subframes, escaped partitions, 5-bit Rice method, wasted-bits flag,
    L/S, S/R, M/S, 3 channels, reserved type, negative LPC shift: the samples the reference
    decoder yields, and the exception it ends with (fixtures from the real reference)"""
    from helpers import DECODE_CASES, STATUS_NAMES, decode_case
    case = DECODE_CASES[name]
    stream, want = decode_case(case)
    rc, pcm, fch, _ = simlib.decode_frames(stream[42:], 16)
    assert STATUS_NAMES.get(rc, "?") == case["raises"]
    assert pcm.shape[0] == want.shape[0]
    if want.shape[0]:
        assert pcm.shape == want.shape and (pcm == want).all()


# ---- second-generation serialiser (flacb200_serialise.cuh): plane_mode bit 11; bits 12-13 pick the
# samples per thread (16 / 8 / 32 / 12), bit 14 a staging window of one worst-case subframe (mid-frame flushes)
SER2 = 0x800
SER2_MODES = {"s16": SER2, "s8": SER2 | 0x1000, "s32": SER2 | 0x2000, "s12": SER2 | 0x3000, "s16_smallwin": SER2 | 0x4000}


@pytest.mark.parametrize("mode", sorted(SER2_MODES))
@pytest.mark.parametrize("name", FAST + ["many_frames"])
def test_sim_serialise_kernel_matches_golden(name, mode):
    case = CASES[name]
    per_thread = {"s8": 8, "s12": 12}.get(mode, 16)
    if case["block_size"] > per_thread * 1024:
        pytest.skip(f"{per_thread} samples per thread: blocks up to {per_thread * 1024}")
    x = case_pcm(case)
    rc, data, offs, _, st = simlib.encode_frames(x, case["bits"], plane_mode=SER2_MODES[mode], **case_kwargs(case))
    assert rc == 0
    assert data == case_stream(case)[42:]
    lens = [4, 4, 34] + [int(b - a) for a, b in zip(offs[:-1], offs[1:])]
    assert lens == case["chunk_lengths"]


@pytest.mark.parametrize("ch,bits,L,q,bs,cls", SHAPES + [(2, 16, 0, 5, 1000, "mixed"), (1, 24, 0, 5, 333, "white"), (7, 8, 6, 6, 96, "white")])
def test_sim_serialise_kernel_other_shapes_against_the_oracle(ch, bits, L, q, bs, cls):
    """channel counts without a vector load path, odd block sizes (partitions that are not a multiple of
    a thread's span, unaligned frame starts), fixed predictors only, 8-bit samples"""
    import flacb200_synth as S
    x = S.pcm(bs * 3 + bs // 3 + 5, ch, bits, 48000, cls, 99 + ch + bits)
    ref = O.encode_frames(x, bits, bs, L, q, 0, 4)
    for mode in (SER2, SER2 | 0x1000, SER2 | 0x3000, SER2 | 0x4000):
        rc, data, offs, info, st = simlib.encode_frames(x, bits, bs, L, q, 0, 4, plane_mode=mode, want_info=True)
        assert rc == 0 and data == ref.data and (offs == ref.frame_offsets).all(), hex(mode)
        rc1, data1, _, info1, _ = simlib.encode_frames(x, bits, bs, L, q, 0, 4, want_info=True)
        for a, b in zip(info, info1):  # the introspection records of both serialisers agree
            assert bytes(a) == bytes(b)


@pytest.mark.parametrize("name", sorted(OOD))
def test_sim_serialise_kernel_out_of_domain_statuses(name):
    """same status, same first bad frame and, in lenient mode, same bytes as the first-generation kernel"""
    case = OOD[name]
    x = ood_pcm(case)
    kw = {k: case[k] for k in ("block_size", "max_lpc_order", "qlp_precision", "rice_min", "rice_max")}
    for flags in (0, simlib.FLAG_LENIENT):
        a = simlib.encode_frames(x, case["bits"], flags=flags, **kw)
        b = simlib.encode_frames(x, case["bits"], flags=flags, plane_mode=SER2, **kw)
        assert a[0] == b[0] and a[1] == b[1] and (a[2] == b[2]).all()
        assert a[4].first_bad_frame == b[4].first_bad_frame and a[4].out_of_domain == b[4].out_of_domain


def test_sim_serialise_kernel_partial_threads_and_rice_min():
    """more threads than spans (idle threads behind the block), rice_min > 0 with an odd block (no admissible order)"""
    import flacb200_synth as S
    x = S.pcm(2304 * 2 + 77, 2, 16, 44100, "mixed", 5)
    rc, data, offs, _, _ = simlib.encode_frames(x, 16, 2304, 8, 7, 1, 6, plane_mode=SER2, threads=192)
    old = simlib.encode_frames(x, 16, 2304, 8, 7, 1, 6)
    assert rc == old[0] == 1 and data == old[1]  # the 77-sample tail has no admissible order >= 1: AssertionError
    with pytest.raises(O.OracleError):
        O.encode_frames(x, 16, 2304, 8, 7, 1, 6)
    full = x[:2304 * 2]
    assert simlib.encode_frames(full, 16, 2304, 8, 7, 1, 6, plane_mode=SER2, threads=192)[1] == O.encode_frames(full, 16, 2304, 8, 7, 1, 6).data
    a = simlib.encode_frames(x, 16, 2304, 8, 7, 1, 6, flags=simlib.FLAG_LENIENT)
    b = simlib.encode_frames(x, 16, 2304, 8, 7, 1, 6, flags=simlib.FLAG_LENIENT, plane_mode=SER2, threads=192)
    assert a[0] == b[0] and a[1] == b[1]


@pytest.mark.parametrize("order", ["reverse", "random:1", "random:2"])
@pytest.mark.parametrize("name", ["c2_white", "c1_sine", "bs255", "c4_mixed"])
def test_sim_thread_order_does_not_change_the_stream(name, order, monkeypatch):
    """Race check on the simulator (compute-sanitizer is closed on the GPU pool): the threads of a CTA are
    resumed in reverse / in a new random order after every rendezvous.  A read that races with a write
    between two barriers -- in-place residuals, atomic-or packing, scan scratch reuse -- would change bytes.
    Both serialising kernels, the analysis / choice stages and the decoder."""
    monkeypatch.setenv("FLACB_SIM_ORDER", order)
    case = CASES[name]
    x = case_pcm(case)
    ref = case_stream(case)[42:]
    for mode in (0, SER2, SER2 | 0x3000, SER2 | 0x4000):
        rc, data, offs, _, _ = simlib.encode_frames(x, case["bits"], plane_mode=mode, **case_kwargs(case))
        assert rc == 0 and data == ref, hex(mode)
    rc, pcm, fch, _ = simlib.decode_frames(ref, case["bits"])
    if case["channels"] == 2:
        assert rc == 0 and (pcm == x).all()


@pytest.mark.parametrize("name", ["c3_white", "c4_mixed", "c2_white", "q8_white", "bs255", "bs192_r0"])
def test_sim_tensor_core_filter_matches_golden_and_the_scalar_flavours(name):
    """choose_mma_kernel (plane_mode bits 9-10 = 3): sample byte planes x coefficient bytes through an emulated
    mma.m16n8k32 -- same sums of |residual| for every order as the scalar flavours (bit 15), same bytes as the goldens;
    three planes for 24-bit samples, two for 16-bit, tails shorter than a tile, orders 12 and 32"""
    case = CASES[name]
    x = case_pcm(case)
    a = simlib.encode_frames(x, case["bits"], plane_mode=3 << 9, want_info=True, **case_kwargs(case))
    b = simlib.encode_frames(x, case["bits"], plane_mode=0x8000 | (2 << 9), want_info=True, **case_kwargs(case))
    assert a[0] == 0 and a[1] == case_stream(case)[42:] == b[1]
    L = case["max_lpc_order"]
    for ia, ib in zip(a[3], b[3]):
        assert [ia.lpc_err[m] for m in range(L)] == [ib.lpc_err[m] for m in range(L)]


@pytest.mark.parametrize("name", ["c3_white", "bs255"])
def test_sim_tensor_core_filter_64_bit_path(name, monkeypatch):
    """the epilogue that keeps 64-bit residuals (prediction sums that do not fit 32 bits once shifted: no filter a
    sane signal produces) is forced on: same bytes"""
    case = CASES[name]
    x = case_pcm(case)
    monkeypatch.setenv("FLACB_SIM_MMA_WIDE", "1")
    a = simlib.encode_frames(x, case["bits"], plane_mode=3 << 9, **case_kwargs(case))
    assert a[0] == 0 and a[1] == case_stream(case)[42:]


@pytest.mark.parametrize("bits,block,L,channels,frames", [(16, 4608, 12, 2, 12), (24, 4096, 32, 2, 4)])
def test_sim_autocorrelation_bits_match_the_oracle_over_many_frames(bits, block, L, channels, frames):
    """every lag of every subframe of a run of frames of the bench's mixed signal, bit pattern against bit
    pattern (the reference's compensated sum() is the correctly rounded exact sum; encoder.py:443-450).  Several
    CTAs in a row with the simulator's poisoned static shared memory: state a frame must not inherit from the
    frame before (a stale per-frame maximum once cost 1 lag in 200 its last bit on the GPU only)."""
    import flacb200_synth as S
    x = S.pcm(frames * block, channels, bits, 44100, "mixed", 1234, start=0)
    ref = O.encode_frames(x, bits, block, L, 5, 0, 5, want_info=True)
    rc, data, _, info, _ = simlib.encode_frames(x, bits, block, L, 5, 0, 5, want_info=True)
    assert rc == 0 and data == ref.data
    bad = 0
    for a, r in zip(info, ref.info):
        for lag in range(L + 1):
            bad += np.float64(a.autoc[lag]).tobytes() != np.float64(r.autoc[lag]).tobytes()
    assert bad == 0
