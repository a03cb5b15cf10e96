"""GPU parity tests proper: the CUDA encode path, called through the package and the C ABI,
against the golden streams minted from the reference and against the CPU oracle."""
import numpy as np
import pytest

import flacb200_synth as S
import oracle as O
from helpers import CASES, OOD, case_kwargs, case_pcm, case_stream, hex_f64, ood_pcm

pytestmark = pytest.mark.gpu


def _params(case):
    from flac.encoder import EncoderParameters
    return EncoderParameters(block_size=case["block_size"],
                             rice_partition_order=range(case["rice_min"], case["rice_max"] + 1),
                             lpc_order=range(case["max_lpc_order"] + 1),
                             qlp_precision=case["qlp_precision"])


def _encode(x, bits, params, **kw):
    from flac import _native as N
    from flac.encoder import encode_frames
    sh = encode_frames(bits, x.shape[1], S.pack_le(x, bits), params, **kw)[0]
    if sh.error:
        raise sh.error
    return sh, N


@pytest.mark.parametrize("mode", ["lookback", "slots"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_stream_bytes(name, mode):
    case = CASES[name]
    x = case_pcm(case)
    from flac import _native as N
    sh, _ = _encode(x, case["bits"], _params(case), flags=N.SLOT_COMPACT if mode == "slots" else 0)
    assert sh.rc == 0
    ref = case_stream(case)
    assert N.stream_prologue(case["rate"], case["channels"], case["bits"], case["n"], case["block_size"]) == ref[:42]
    assert sh.out.tobytes() == ref[42:]
    lens = [4, 4, 34] + [int(b - a) for a, b in zip(sh.offs[:-1], sh.offs[1:])]
    assert lens == case["chunk_lengths"]


@pytest.mark.parametrize("ser", ["1", "2,0", "2,1", "2,2", "2,3", "2,4", "2,0,1"])
@pytest.mark.parametrize("name", ["c1_sine", "c2_white", "c2_clip", "c3_white", "c4_mixed", "bs255", "bs1000", "q15_white24", "many_frames"])
def test_golden_stream_bytes_every_serialiser(name, ser, monkeypatch):
    """FLACB200_SERIALISE=1: the first-generation serialising kernel (encode_frames_kernel); 2,<v>[,<frames per
    SM>]: serialise_kernel with 16 / 16 / 8 / 32 / 12 samples per thread (the default picks the variant with the most
    resident threads; a variant that cannot hold the block falls back to the first-generation kernel)"""
    monkeypatch.setenv("FLACB200_SERIALISE", ser)
    case = CASES[name]
    x = case_pcm(case)
    sh, _ = _encode(x, case["bits"], _params(case))
    assert sh.rc == 0
    assert sh.out.tobytes() == case_stream(case)[42:]
    lens = [4, 4, 34] + [int(b - a) for a, b in zip(sh.offs[:-1], sh.offs[1:])]
    assert lens == case["chunk_lengths"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_stage_values(name):
    """fixed errors, autocorrelation bits, quantised coefficients, LPC errors, Rice decisions"""
    case = CASES[name]
    x = case_pcm(case)
    sh, _ = _encode(x, case["bits"], _params(case), want_info=True)
    L = case["max_lpc_order"]
    lags = bad_lags = 0
    for info, ref in zip(sh.info, case["subframes"]):
        assert ("lpc" if info.type else "fixed") == ref["type"]
        assert info.order == ref["order"]
        assert info.fixed_order == ref["fixed_order"]
        if ref["fixed_err"] is not None:
            assert list(info.fixed_err) == ref["fixed_err"]
        if L > 0:
            want = [hex_f64(h) for h in ref["autoc"]]
            got = [info.autoc[i] for i in range(L + 1)]
            lags += L + 1
            bad_lags += sum(1 for a, b in zip(got, want) if a != b)
            if got == want:
                for m, (qc, shf) in enumerate(ref["quant"]):
                    assert list(info.lpc_q[m][:info.lpc_ncoef[m]]) == qc
                    assert info.lpc_shift[m] == shf
                assert info.lpc_order == ref["lpc_order"]
                assert info.lpc_err[info.lpc_order - 1] == ref["lpc_total"]
        assert info.porder == ref["porder"]
        assert info.method == ref["method"]
        assert list(info.params[:min(info.nparams, 64)]) == ref["params"][:64]
    # the double-double accumulation reproduces Python's compensated sum() on every lag seen
    assert bad_lags == 0, f"{bad_lags}/{lags} autocorrelation lags differ in the last bit"


@pytest.mark.parametrize("name", sorted(OOD))
def test_out_of_domain_raises_like_the_reference(name):
    case = OOD[name]
    x = ood_pcm(case)
    from flac.encoder import encode_pcm
    with pytest.raises(Exception) as ei:
        b"".join(encode_pcm(case["rate"], case["bits"], x.shape[1], S.pack_le(x, case["bits"]), _params(case)))
    assert type(ei.value).__name__ == case["raises"]


@pytest.mark.parametrize("name", sorted(OOD))
def test_out_of_domain_lenient_round_trips(name):
    case = OOD[name]
    x = ood_pcm(case)
    from flac.encoder import encode_pcm
    stream = b"".join(encode_pcm(case["rate"], case["bits"], x.shape[1], S.pack_le(x, case["bits"]),
                                 _params(case), strict=False))
    rate, bits, ch, total, pcm = O.decode_stream(stream)
    assert (pcm == x).all()


def test_module_api_iterator_matches_golden():
    from flac.encoder import encode
    case = CASES["c2_white"]
    x = case_pcm(case)
    chunks = list(encode(case["rate"], case["bits"], case["channels"], case["n"], iter(x.tolist()), _params(case)))
    assert [len(c) for c in chunks] == case["chunk_lengths"]
    assert b"".join(chunks) == case_stream(case)


def test_first_frame_number_and_sharding_are_exact():
    """frames [a, b) encoded separately with their coded numbers concatenate to the whole"""
    case = CASES["many_frames"]
    x = case_pcm(case)
    p = _params(case)
    whole, _ = _encode(x, case["bits"], p)
    bs = case["block_size"]
    cut = 57
    a, _ = _encode(x[:cut * bs], case["bits"], p)
    b, _ = _encode(x[cut * bs:], case["bits"], p, first_frame_number=cut)
    assert a.out.tobytes() + b.out.tobytes() == whole.out.tobytes()


@pytest.mark.parametrize("bits,cls,L,bs", [(16, "mixed", 12, 4608), (24, "mixed", 12, 4608), (16, "white", 8, 4096),
                                           (24, "white", 32, 16384), (16, "mixed", 0, 4608), (8, "white", 6, 1152)])
def test_against_oracle_on_fresh_inputs(bits, cls, L, bs):
    x = S.pcm(bs * 9 + bs // 3 + 17, 2, bits, 96000, cls, 4242 + bits + L)
    if bits == 8:
        x = (x // 1).astype(np.int32)
    from flac.encoder import EncoderParameters
    p = EncoderParameters(block_size=bs, rice_partition_order=range(0, 7), lpc_order=range(L + 1), qlp_precision=7)
    try:
        ref = O.encode_frames(x, bits, bs, L, 7, 0, 6)
    except O.OracleError:
        pytest.skip("input outside the reference's domain")
    sh, _ = _encode(x, bits, p)
    assert sh.rc == 0
    assert sh.out.tobytes() == ref.data
    assert (sh.offs == ref.frame_offsets).all()


@pytest.mark.parametrize("ch,bits,L,q,bs,cls", [(3, 16, 12, 12, 2304, "mixed"), (5, 24, 8, 9, 1024, "white"),
                                                (1, 16, 12, 8, 4608, "white"), (6, 20, 12, 5, 576, "mixed"),
                                                (2, 12, 4, 5, 192, "white"), (4, 24, 32, 14, 4096, "white"),
          (2, 16, 32, 5, 4096, "white"), (2, 16, 32, 8, 1024, "white")])
def test_other_channel_counts_widths_and_precisions_against_the_oracle(ch, bits, L, q, bs, cls):
    """3 / 5 / 6 / 4 channels, 12 / 20-bit samples, precisions 8 (coefficient bytes at full range),
    9, 12, 14 (the 32/64-bit flavour of the all-orders filter), order 32 on 16-bit samples (dp2a bodies
    up to 32 wide), short blocks; many frames"""
    x = S.pcm(bs * 40 + bs // 3 + 5, ch, bits, 48000, cls, 99 + ch + bits)
    from flac.encoder import EncoderParameters
    p = EncoderParameters(block_size=bs, rice_partition_order=range(0, 5), lpc_order=range(L + 1), qlp_precision=q)
    try:
        ref = O.encode_frames(x, bits, bs, L, q, 0, 4, threads=8)
    except O.OracleError:
        pytest.skip("input outside the reference's domain")
    sh, _ = _encode(x, bits, p)
    assert sh.rc == 0
    assert sh.out.tobytes() == ref.data
    assert (sh.offs == ref.frame_offsets).all()


def test_large_config2_slice_round_trip_and_oracle_spot_check():
    """60 s of config-2 audio: oracle bytes on a slice, oracle decode of the whole"""
    n = 44100 * 60
    x = S.pcm(n, 2, 16, 44100, "mixed", 0xF1AC0002)
    from flac.encoder import EncoderParameters, encode_pcm
    p = EncoderParameters(block_size=4608, rice_partition_order=range(0, 6), lpc_order=range(13), qlp_precision=5)
    stream = b"".join(encode_pcm(44100, 16, 2, S.pack_le(x, 16), p))
    enc = O.encode_frames(x, 16, 4608, 12, 5, 0, 5, threads=16)
    assert stream[42:] == enc.data
    rate, bits, ch, total, pcm = O.decode_stream(stream)
    assert (rate, bits, ch, total) == (44100, 16, 2, n)
    assert (pcm == x).all()


def test_chunked_host_pipeline_matches_single_launch(monkeypatch):
    """host buffers: the library cuts the frame range into chunks with several in flight (H2D,
    kernels, D2H overlapped, one stream per copy direction); bytes, offsets and error position
    must not depend on the cut (FLACB200_CHUNK_FRAMES is taken as it is: 71 frames in chunks
    of 5 / 16 / 33 / one chunk)"""
    from flac import _native as N
    x = S.pcm(4608 * 70 + 999, 2, 16, 44100, "mixed", 0xF1AC0002)
    p = _params(CASES["c2_white"])
    ref = O.encode_frames(x, 16, 4608, 12, 5, 0, 5, threads=8)
    for chunk in ("5", "16", "33", "100000"):
        monkeypatch.setenv("FLACB200_CHUNK_FRAMES", chunk)
        sh, _ = _encode(x, 16, p)
        assert sh.rc == 0 and sh.out.tobytes() == ref.data and (sh.offs == ref.frame_offsets).all(), chunk
    # an out-of-domain frame (digital silence) in the third chunk: everything in front of it is returned
    y = x.copy()
    y[4608 * 50:4608 * 51] = 0
    monkeypatch.setenv("FLACB200_CHUNK_FRAMES", "16")
    sh, _ = _encode(y, 16, p)
    assert sh.rc == N.E_ZERODIV and sh.stats.first_bad_frame == 50
    assert sh.out.tobytes() == O.encode_frames(y[:4608 * 50], 16, 4608, 12, 5, 0, 5, threads=8).data


def test_sharding_over_two_gpus_in_one_process():
    """encode_pcm(devices=[0, 1]): frame ranges on two GPUs from two host threads, byte ranges
    concatenated on the host"""
    from flac import _native as N
    if N.lib().flacb_device_count() < 2:
        pytest.skip("needs two GPUs")
    from flac.encoder import encode_pcm
    case = CASES["many_frames"]
    x = case_pcm(case)
    stream = b"".join(encode_pcm(case["rate"], case["bits"], 2, S.pack_le(x, 16), _params(case), devices=[0, 1]))
    assert stream == case_stream(case)


def test_subframe_records_use_the_reference_data_model():
    from flac.common import RiceCodingMethod, SubframeFixed, SubframeLPC, SubframeTypeFixed, SubframeTypeLPC
    from flac.encoder import subframe_records
    case = CASES["c2_white"]
    sh, _ = _encode(case_pcm(case), case["bits"], _params(case), want_info=True)
    recs = subframe_records(sh, case["qlp_precision"])
    assert len(recs) == len(case["subframes"])
    for (header, body, residual), ref in zip(recs, case["subframes"]):
        if ref["type"] == "lpc":
            assert isinstance(header.type_, SubframeTypeLPC) and isinstance(body, SubframeLPC)
            assert body.coefficients == ref["quant"][ref["order"] - 1][0] and body.shift == ref["quant"][ref["order"] - 1][1]
        else:
            assert isinstance(header.type_, SubframeTypeFixed) and isinstance(body, SubframeFixed)
        assert header.type_.order == body.order == ref["order"]
        assert residual.partition_order == ref["porder"] and residual.coding_method is RiceCodingMethod(ref["method"])
        assert [p.parameter for p in residual.partitions][:64] == ref["params"][:64]


@pytest.mark.parametrize("flavour", ["3", "2", "1"])
@pytest.mark.parametrize("name", ["c3_white", "c3_clip", "c4_mixed", "c4_white", "c2_white", "q8_white", "bs255", "many_frames"])
def test_golden_stream_bytes_every_filter_flavour(name, flavour, monkeypatch):
    """stage C1, the all-orders LPC filter: FLACB200_FLAVOUR=3 byte planes on the tensor cores (mma.sync s8; the
    default for samples wider than 16 bits, forced here for 16-bit samples too), 2 / 1 the scalar flavours"""
    monkeypatch.setenv("FLACB200_FLAVOUR", flavour)
    case = CASES[name]
    sh, _ = _encode(case_pcm(case), case["bits"], _params(case), want_info=True)
    assert sh.rc == 0
    assert sh.out.tobytes() == case_stream(case)[42:]
    for info, ref in zip(sh.info, case["subframes"]):
        if case["max_lpc_order"] > 0 and ref.get("lpc_total") is not None:
            assert info.lpc_order == ref["lpc_order"] and info.lpc_err[info.lpc_order - 1] == ref["lpc_total"]
