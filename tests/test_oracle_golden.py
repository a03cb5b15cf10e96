"""Pin the CPU oracle (oracle/flac_oracle.c) against the reference's own outputs
(tests/golden/, minted by tests/golden/make_golden.py from turlando/flac-py)."""
import numpy as np
import pytest

import oracle as O
from helpers import CASES, GOLD, OOD, case_kwargs, case_pcm, case_stream, hex_f64, ood_pcm, sha


@pytest.mark.parametrize("name", sorted(CASES))
def test_encode_bytes_and_decisions(name):
    case = CASES[name]
    x = case_pcm(case)
    enc = O.encode_frames(x, case["bits"], want_info=True, **case_kwargs(case))
    stream = O.stream_prologue(case["rate"], case["channels"], case["bits"], case["n"],
                               case["block_size"]) + enc.data
    assert sha(stream) == case["stream_sha256"]
    assert stream == case_stream(case)
    # generator chunking of the reference: 4, 4, 34, then one chunk per frame
    lens = [4, 4, 34] + [int(b - a) for a, b in zip(enc.frame_offsets[:-1], enc.frame_offsets[1:])]
    assert lens == case["chunk_lengths"]
    L = case["max_lpc_order"]
    for info, ref in zip(enc.info, case["subframes"]):
        assert ("lpc" if info.type else "fixed") == ref["type"]
        assert info.order == ref["order"]
        assert info.fixed_order == ref["fixed_order"]
        if ref["fixed_err"] is not None:
            assert list(info.fixed_err) == ref["fixed_err"]
        if L > 0:
            assert [info.autoc[i] for i in range(L + 1)] == [hex_f64(h) for h in ref["autoc"]]
            for m, (qc, sh) in enumerate(ref["quant"]):
                assert list(info.lpc_q[m][:info.lpc_ncoef[m]]) == qc
                assert info.lpc_shift[m] == sh
            assert info.lpc_order == ref["lpc_order"]
            assert info.lpc_err[info.lpc_order - 1] == ref["lpc_total"]
        assert info.porder == ref["porder"]
        assert info.method == ref["method"]
        assert list(info.params[:info.nparams]) == ref["params"][:O.INFO_PARAMS]


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if "decode" in c))
def test_decode_matches_reference_decoder(name):
    case = CASES[name]
    rate, bits, ch, total, pcm = O.decode_stream(case_stream(case))
    d = case["decode"]
    assert (rate, bits, ch, total) == (d["rate"], d["bits"], d["channels"], d["total"])
    assert sha(np.ascontiguousarray(pcm, dtype=np.int32).tobytes()) == d["pcm_sha256"]
    assert (pcm == case_pcm(case)).all()


@pytest.mark.parametrize("name", sorted(OOD))
def test_out_of_domain_raises_like_the_reference(name):
    case = OOD[name]
    x = ood_pcm(case)
    kw = {k: case[k] for k in ("block_size", "max_lpc_order", "qlp_precision", "rice_min", "rice_max")}
    with pytest.raises(O.OracleError) as ei:
        O.encode_frames(x, case["bits"], **kw)
    assert ei.value.exc_class.__name__ == case["raises"]


def test_threads_give_identical_bytes():
    case = CASES["many_frames"]
    x = case_pcm(case)
    a = O.encode_frames(x, case["bits"], **case_kwargs(case))
    b = O.encode_frames(x, case["bits"], threads=5, **case_kwargs(case))
    assert a.data == b.data and (a.frame_offsets == b.frame_offsets).all()


def test_primitive_known_answers():
    kat = GOLD["kat"]
    assert O.crc8(b"123456789") == kat["crc8_123456789"] == 0xF4
    assert O.crc16(b"123456789") == kat["crc16_123456789"] == 0xFEE8
    for v, h in kat["coded_number"].items():
        assert O.coded_number(int(v)).hex() == h
    with pytest.raises(ValueError):
        O.coded_number(1 << 31)
    for n, hs in kat["tukey"].items():
        assert O.tukey(int(n)).tolist() == [hex_f64(h) for h in hs]
    for item in kat["float_sum"]:
        x = np.array([hex_f64(h) for h in item["x"]])
        assert O.neumaier_sum(x) == hex_f64(item["sum"])


def test_frame_headers_known_answers():
    # SURVEY.md 9.8 worked bytes + headers minted from encoder.py:194 put_frame_header
    for key, h in GOLD["kat"]["frame_header"].items():
        bs, num = (int(t) for t in key.split(":"))
        x = np.ones((bs, 2), dtype=np.int32)
        x[::2] = -3
        x[1::3] = 2
        enc = O.encode_frames(x, 16, bs, max_lpc_order=0, rice_min=0, rice_max=0,
                              first_frame_number=num)
        assert enc.data[:len(h) // 2].hex() == h
    assert GOLD["kat"]["frame_header"]["4608:0"] == "fff8501000f9"
    assert GOLD["kat"]["frame_header"]["16384:70000"] == "fff8e010f09185b030"
    assert GOLD["kat"]["frame_header"]["864:200"] == "fff87010c388035f02"


@pytest.mark.reference
def test_live_reference_spot_check():
    """When the real reference is mounted, re-derive one case end to end."""
    import refharness as R
    import flacb200_synth as S
    x = S.pcm(2000 + 300, 2, 16, 44100, "white", 99)
    chunks = R.encode_chunks(x.tolist(), x.shape[0], 2, 44100, 16, 1000, 10, 6, 1, 3)
    ours = O.encode_stream(x, 44100, 16, 1000, max_lpc_order=10, qlp_precision=6, rice_min=1,
                           rice_max=3)
    assert b"".join(chunks) == ours


@pytest.mark.parametrize("name", sorted(__import__("helpers").DECODE_CASES))
def test_oracle_decodes_crafted_streams_like_the_reference(name):
    """stream features the reference's encoder never produces (constant / verbatim subframes,
    escaped partitions, 5-bit Rice method, wasted-bits flag, L/S, S/R, M/S, 3 channels, reserved
    type, negative LPC shift): fixtures minted from the real reference decoder by
    tests/golden/make_decode_golden.py"""
    from helpers import DECODE_CASES, decode_case
    case = DECODE_CASES[name]
    stream, want = decode_case(case)
    try:
        pcm = O.decode_stream(stream)[4]
        assert case["raises"] is None
        assert pcm.shape == want.shape and (pcm == want).all()
    except O.OracleError as e:
        assert e.exc_class.__name__ == case["raises"]
