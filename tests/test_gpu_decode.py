"""GPU decode (BASELINE.json config 5) against the reference decoder's results (golden) and the
CPU oracle, through the package and the C ABI."""
import io
import os
import subprocess
import sys
import wave

import numpy as np
import pytest

import flacb200_synth as S
import oracle as O
from helpers import CASES, case_pcm, case_stream, sha

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if "decode" in c))
def test_golden_streams_decode_like_the_reference(name):
    from flac.decoder import decode
    case = CASES[name]
    rate, bits, ch, total, samples = decode(io.BytesIO(case_stream(case)))
    d = case["decode"]
    assert (rate, bits, ch, total) == (d["rate"], d["bits"], d["channels"], d["total"])
    pcm = np.array(list(samples), dtype=np.int32)
    assert sha(np.ascontiguousarray(pcm).tobytes()) == d["pcm_sha256"]
    assert (pcm == case_pcm(case)).all()


@pytest.mark.parametrize("name", sorted(n for n, c in CASES.items() if "decode" not in c))
def test_streams_the_reference_cannot_decode_fail_the_same_way(name):
    """mono / 8-channel streams carry the L_R nibble (encoder.py:95): the reference decoder
    asserts on them, and so does the oracle"""
    from flac.decoder import decode
    case = CASES[name]
    try:
        O.decode_stream(case_stream(case))
        pytest.skip("the oracle decodes this one")
    except O.OracleError as e:
        want = e.exc_class
    with pytest.raises(want):
        rate, bits, ch, total, samples = decode(io.BytesIO(case_stream(case)))
        list(samples)


def test_truncated_stream_drops_the_last_frame():
    from flac.decoder import decode_frames
    case = CASES["c2_white"]
    s = case_stream(case)
    cut = s[:len(s) - 100]
    ref = O.decode_stream(cut)[4]
    pcm, fch = decode_frames(cut[42:], 16)
    assert fch == 2 and pcm.shape == ref.shape and (pcm == ref).all()
    assert pcm.shape[0] == 4608 * 3


def test_spec_channels_makes_mono_and_multichannel_decodable():
    from flac.encoder import EncoderParameters, encode_pcm
    from flac.decoder import decode
    for ch, bits in ((1, 16), (8, 24), (3, 16)):
        x = S.pcm(4608 * 2 + 777, ch, bits, 48000, "mixed", 77 + ch)
        p = EncoderParameters(block_size=4608, rice_partition_order=range(0, 6), lpc_order=range(13), qlp_precision=5)
        stream = b"".join(encode_pcm(48000, bits, ch, S.pack_le(x, bits), p, spec_channels=True))
        ref = O.decode_stream(stream)[4]
        rate, sbits, sch, total, samples = decode(io.BytesIO(stream))
        pcm = np.array(list(samples), dtype=np.int32)
        assert (sch, total) == (ch, x.shape[0])
        assert (pcm == x).all() and (ref == x).all()


def test_config5_round_trip_of_a_config2_stream():
    """60 s of the config-2 stream: GPU encode, GPU decode, bit-exact PCM; the oracle decoder
    (pinned to the reference decoder by the golden fixtures) agrees"""
    from flac.encoder import EncoderParameters, encode_pcm
    from flac.decoder import decode_pcm
    n = 44100 * 60
    x = S.pcm(n, 2, 16, 44100, "mixed", 0xF1AC0002)
    p = EncoderParameters(block_size=4608, rice_partition_order=range(0, 6), lpc_order=range(13), qlp_precision=5)
    stream = b"".join(encode_pcm(44100, 16, 2, S.pack_le(x, 16), p))
    info, pcm = decode_pcm(io.BytesIO(stream))
    assert (info.sample_rate, info.sample_size, info.channels, info.samples) == (44100, 16, 2, n)
    assert pcm.shape == x.shape and (pcm == x).all()
    ref = O.decode_stream(stream)[4]
    assert (ref == pcm).all()


def test_cli_encode_decode_round_trip(tmp_path):
    x = S.pcm(44100 * 3 + 123, 2, 16, 44100, "mixed", 5)
    wav_in, flac_out, wav_out = tmp_path / "in.wav", tmp_path / "out.flac", tmp_path / "out.wav"
    with wave.open(str(wav_in), "wb") as w:
        w.setnchannels(2), w.setsampwidth(2), w.setframerate(44100)
        w.writeframes(S.pack_le(x, 16))
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = os.path.join(root, "flac-py_b200")
    r = subprocess.run([sys.executable, "-m", "flac", "encode", str(wav_in), str(flac_out)], env=env,
                       capture_output=True, text=True, check=True)
    assert r.stdout.startswith("Encoding completed in ")
    assert flac_out.read_bytes() == O.encode_stream(x, 44100, 16, 4608, max_lpc_order=12, qlp_precision=5,
                                                    rice_min=0, rice_max=5)
    r = subprocess.run([sys.executable, "-m", "flac", "decode", str(flac_out), str(wav_out)], env=env,
                       capture_output=True, text=True, check=True)
    assert r.stdout.startswith("Decoding completed in ")
    with wave.open(str(wav_out), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 44100, x.shape[0])
        assert w.readframes(w.getnframes()) == S.pack_le(x, 16)


def test_decode_ignores_crcs_like_the_reference():
    """a wrong header CRC-8 / frame CRC-16 does not stop the reference (decoder.py:128, :175):
    the CRC-8 filter of the fast path must fall back to the unfiltered scan"""
    from flac.decoder import decode_frames
    case = CASES["c2_white"]
    stream = bytearray(case_stream(case))
    from helpers import case_kwargs
    enc = O.encode_frames(case_pcm(case), 16, **case_kwargs(case))
    f1 = 42 + int(enc.frame_offsets[1])
    stream[f1 + 5] ^= 0x5A
    stream[-1] ^= 0xFF
    ref = O.decode_stream(bytes(stream))[4]
    pcm, fch = decode_frames(bytes(stream[42:]), 16)
    assert pcm.shape == ref.shape and (pcm == ref).all()


def test_decode_variable_block_sizes():
    """frames larger than frame 0's: the fast path defers them, the unfiltered repeat decodes them"""
    from flac.decoder import decode_frames
    a = S.pcm(1000 * 2, 2, 16, 44100, "white", 21)
    b = S.pcm(4096 + 500, 2, 16, 44100, "sine", 22)
    frames = O.encode_frames(a, 16, 1000, 8, 6, 0, 3).data + O.encode_frames(b, 16, 4096, 8, 6, 0, 3).data
    pcm, fch = decode_frames(frames, 16)
    assert fch == 2 and (pcm == np.concatenate([a, b])).all()


@pytest.mark.parametrize("name", sorted(__import__("helpers").DECODE_CASES))
def test_crafted_streams_decode_like_the_reference(name):
    """constant / verbatim subframes, escaped partitions, 5-bit Rice method, wasted-bits flag,
    L/S, S/R, M/S, 3 channels, reserved type, negative LPC shift: through flac.decoder.decode, the
    samples the reference decoder yields and the exception it ends with (fixtures from the real
    reference, tests/golden/make_decode_golden.py)"""
    from flac.decoder import decode
    from helpers import DECODE_CASES, decode_case
    case = DECODE_CASES[name]
    stream, want = decode_case(case)
    rate, bits, ch, total, samples = decode(io.BytesIO(stream))
    got, raised = [], None
    try:
        for row in samples:
            got.append(row)
    except Exception as e:  # noqa: BLE001
        raised = type(e).__name__
    assert raised == case["raises"]
    assert got == case["samples"]


def test_scan_frames_gives_the_frame_index_and_headers():
    """flacb_scan_frames / flac.decoder.scan_frames: the offsets of the frames of a golden stream are the
    chunk boundaries the reference's encoder yielded; frame_headers() parses them into flac.common.FrameHeader"""
    from flac.common import BlockingStrategy, Channels
    from flac.decoder import frame_headers, scan_frames
    for name in ("many_frames", "c2_white", "c3_white", "bs255"):
        case = CASES[name]
        s = case_stream(case)[42:]
        offs, sizes, ns, rc = scan_frames(s, case["bits"])
        assert rc == 0 and ns == case["n"]
        want = np.cumsum([0] + case["chunk_lengths"][3:])
        assert (offs == want).all()
        assert sizes.sum() == case["n"] and (sizes[:-1] == case["block_size"]).all()
        hs = frame_headers(s, case["bits"])
        assert [h.coded_number for h in hs] == list(range(len(sizes)))
        assert all(h.blocking_strategy is BlockingStrategy.Fixed and h.channels is Channels.L_R and h.sample_rate is None
                   and h.sample_size is None for h in hs)
        assert [h.block_size for h in hs] == [int(v) for v in sizes]


def test_decode_sizes_its_output_from_the_stream_not_from_streaminfo():
    """a STREAMINFO sample count of 2^36 - 1 (or 1) must not drive the allocation: the library says
    how much the stream decodes to and hands the decoded frames over without decoding twice"""
    import io
    from flac.decoder import decode_frames, decode_pcm
    case = CASES["c2_white"]
    s = bytearray(case_stream(case))
    x = case_pcm(case)
    for total in ((1 << 36) - 1, 1):
        v = int.from_bytes(s[18:26], "big")
        v = (v & ~((1 << 36) - 1)) | total
        s[18:26] = v.to_bytes(8, "big")
        info, pcm = decode_pcm(io.BytesIO(bytes(s)))
        assert info.samples == total and (pcm == x).all()
    pcm, fch = decode_frames(bytes(s[42:]), 16, cap_sample_frames=3)  # far too small a hint
    assert fch == 2 and (pcm == x).all()
