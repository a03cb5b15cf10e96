"""Host side of the drop-in package: bit reader / writer, coded numbers, CRCs, helpers, the
parameter object and the CLI parser.  The known answers are the ones the reference's own tests
pin (test/test_binary.py, test/test_coded_number.py) plus the ones SURVEY.md section 4 lists."""
import doctest
import io

import pytest

from flac import binary, coded_number, common, crc, utils
from flac.binary import Get, Put
from flac.encoder import EncoderParameters


def test_get_widths_1_to_9_then_unaligned():
    g = Get(io.BytesIO((0b_11010010_00100001_00000100_00001000_00001000_00000111).to_bytes(6, "big")))
    assert [g.uint(n) for n in range(1, 10)] == [1, 2, 4, 8, 16, 32, 64, 128, 256]
    assert g.uint(3) == 0b111


def test_get_wide_reads():
    assert Get(io.BytesIO((0b10000000_00000001).to_bytes(2, "big"))).uint(16) == 0b10000000_00000001
    g = Get(io.BytesIO(bytes([0xFF, 0x00, 0xAA, 0x55, 0x80])))
    assert g.uint(4) == 0xF and g.uint(25) == 0b1111000000001010101001010 and g.uint(3) == 0b101
    assert g.uint(8) == 0x80
    with pytest.raises(EOFError):
        g.uint(1)
    g = Get(io.BytesIO(bytes([0b11100000])))
    assert g.sint(3) == -1 and g.bool() is False


def test_put_aligned_unaligned_multibyte():
    p = Put()
    p.uint(0b1, 1), p.uint(0b10, 2), p.uint(0b100, 3), p.uint(0b10, 2)
    assert p.buffer == bytes([0b11010010])
    p = Put()
    p.uint(0xABCD, 16), p.uint(0x5, 4), p.uint(0xFFF, 12)
    assert p.buffer == bytes([0xAB, 0xCD, 0x5F, 0xFF])
    p = Put()
    p.uint(-1, 5), p.uint(-4, 3)  # two's complement by masking (binary.py:172)
    assert p.buffer == bytes([0b11111100])
    p = Put()
    p.uint(1, 3)
    with pytest.raises(AssertionError):
        p.buffer


def test_put_streaminfo_field_widths():
    """the field widths of STREAMINFO (a regression the reference pins in test_put_5)"""
    p = Put()
    for v, n in ((4608, 16), (4608, 16), (0, 24), (0, 24), (44100, 20), (1, 3), (15, 5), (22050, 36)):
        p.uint(v, n)
    p.bytes(bytes(16))
    assert p.buffer.hex() == "120012000000000000000ac442f000005622" + "00" * 16


@pytest.mark.parametrize("decoded,encoded,size", [
    (0b00000000, 0b00000000, 1), (0b00011111, 0b00011111, 1), (0b01111111, 0b01111111, 1),
    (0b10000000, 0b11000010_10000000, 2), (0b10101010_1, 0b11000101_10010101, 2),
    (0b10101010_101, 0b11010101_10010101, 2)])
def test_coded_number_known_answers(decoded, encoded, size):
    assert int.from_bytes(coded_number.encode(decoded), "big") == encoded
    assert coded_number.decode(encoded.to_bytes(size, "big")) == decoded


def test_coded_number_sizes_and_round_trip():
    for x, size in ((5742, 3), (21093, 3), (374999, 4), ((1 << 26) - 1, 5), ((1 << 31) - 1, 6)):
        e = coded_number.encode(x)
        assert len(e) == size == coded_number.required_bytes(x) == coded_number.following_bytes(e[0]) + 1
        assert coded_number.decode(e) == x
    assert coded_number.encode(70000).hex() == "f09185b0" and coded_number.encode(200).hex() == "c388"
    with pytest.raises(ValueError):
        coded_number.encode(1 << 31)


def test_crc_known_answers():
    assert crc.crc8(b"123456789", common.CRC8_POLYNOMIAL) == 0xF4
    assert crc.crc16(b"123456789", common.CRC16_POLYNOMIAL) == 0xFEE8
    assert crc.crc_table(8, 0x107)[:4] == [0, 7, 0xE, 9]
    assert crc.crc_table(16, 0x18005)[:4] == [0, 0x8005, 0x800F, 0xA]
    assert crc.crc8(bytes.fromhex("fff8501000"), 0x107) == 0xF9  # frame 0 header of SURVEY 9.8


def test_doctests():
    for mod in (binary, utils):
        assert doctest.testmod(mod).failed == 0


def test_utils():
    assert utils.clamp(5, 0, 3) == 3 and utils.clamp(-5, 0, 3) == 0 and utils.clamp(2, 0, 3) == 2
    assert [utils.zigzag_encode(v) for v in (0, -1, 1, -2, 2)] == [0, 1, 2, 3, 4]
    assert all(utils.zigzag_decode(utils.zigzag_encode(v)) == v for v in range(-300, 300))
    assert utils.log2i(1) == 0 and utils.log2i(4608) == 12
    assert utils.argparse_range("8") == range(0, 9) and utils.argparse_range("3,4") == range(3, 5)
    with pytest.raises(AssertionError):
        utils.argparse_range("5,2")
    with pytest.raises(ValueError):
        list(utils.batch(iter([1]), 0))


def test_common_tables():
    assert common.BLOCK_SIZE_ENCODING[4608] == 0b0101 and common.BLOCK_SIZE_ENCODING[16384] == 0b1110
    assert common.BLOCK_SIZE_ENCODING[192] == 1 and common.BLOCK_SIZE_ENCODING[256] == 8
    assert common.CHANNELS_ENCODING[common.Channels.L_R] == 1 and common.Channels.M_S.count == 2
    assert common.Channels.FL_FR_FC_LFE_BL_BR_SL_SR.count == 8
    assert common.FIXED_PREDICTOR_COEFFICIENTS[4] == (4, -6, 4, -1)


def test_encoder_parameters_asserts():
    EncoderParameters(block_size=4608, rice_partition_order=range(6), lpc_order=range(13), qlp_precision=5)
    with pytest.raises(AssertionError):
        EncoderParameters(4608, range(6), range(1, 13), 5)
    with pytest.raises(AssertionError):
        EncoderParameters(4608, range(6), range(34), 5)
    with pytest.raises(AssertionError):
        EncoderParameters(4608, range(6), range(13), 4)


def test_encode_asserts_lpc_order_for_low_rates_before_touching_the_gpu():
    from flac.encoder import encode
    p = EncoderParameters(4608, range(6), range(33), 5)
    with pytest.raises(AssertionError):
        next(encode(44100, 16, 2, 0, iter([]), p))


def test_cli_parser_defaults_match_the_reference():
    from flac.__main__ import make_argument_parser
    a = make_argument_parser().parse_args(["encode", "in.wav", "out.flac"])
    assert (a.block_size, a.max_lpc_order, a.qlp_coeff_precision, a.rice_partition_order) == (4608, 12, 5, range(0, 6))
    a = make_argument_parser().parse_args(["encode", "in.wav", "out.flac", "-b", "16384", "-l", "32", "-q", "7", "-r", "2,8"])
    assert (a.block_size, a.max_lpc_order, a.qlp_coeff_precision, a.rice_partition_order) == (16384, 32, 7, range(2, 9))
    a = make_argument_parser().parse_args(["decode", "in.flac", "out.wav"])
    assert a.action == "decode"


def test_complete_streaminfo_fills_frame_sizes_and_md5():
    """extension: the fields the reference leaves at zero (encoder.py:75-81)"""
    import hashlib
    from flac.encoder import complete_streaminfo
    import oracle as O
    block = O.stream_prologue(44100, 2, 16, 1000, 4608)[8:]
    assert len(block) == 34 and block[4:10] == bytes(6) and block[18:] == bytes(16)
    pcm = bytes(range(256)) * 4
    out = complete_streaminfo(block, [14, 9000, 77], pcm)
    assert out[:4] == block[:4] and out[10:18] == block[10:18]          # block sizes, rate/channels/bits/total untouched
    assert int.from_bytes(out[4:7], "big") == 14 and int.from_bytes(out[7:10], "big") == 9000
    assert out[18:] == hashlib.md5(pcm).digest()
    assert complete_streaminfo(block, []) == block


def test_seektable_block_is_skipped_by_the_decoder_and_points_at_frames():
    """extension: a SEEKTABLE block built from the frame offsets; the
    reference's decoder (here: its restatement) skips it and decodes the same PCM, and every seek
    point lands on a frame sync code"""
    from flac.encoder import seektable_block
    import flacb200_synth as S
    import oracle as O
    x = S.pcm(1000 * 7 + 333, 2, 16, 44100, "white", 5)
    enc = O.encode_frames(x, 16, 1000, 8, 6, 0, 3)
    pro = bytearray(O.stream_prologue(44100, 2, 16, x.shape[0], 1000))
    assert pro[4] == 0x80
    pro[4] = 0x00                                             # STREAMINFO is no longer the last block
    block = seektable_block(enc.frame_offsets, 1000, x.shape[0], every=3)
    assert block[0] == 0x83 and int.from_bytes(block[1:4], "big") == 18 * 3
    stream = bytes(pro) + block + enc.data
    assert (O.decode_stream(stream)[4] == x).all()
    for k in range(3):
        p = block[4 + 18 * k:4 + 18 * (k + 1)]
        sample, off, count = int.from_bytes(p[:8], "big"), int.from_bytes(p[8:16], "big"), int.from_bytes(p[16:], "big")
        assert sample == 3 * k * 1000 and enc.data[off:off + 2] == b"\xff\xf8"
        assert count == (1000 if 3 * k < 7 else 333)


def test_chunk_plan_of_the_host_path():
    """whole waves, about 16 MB, short first and last chunk; every frame in exactly one chunk"""
    import simlib
    # config 2: 18 KB frames, waves of 444 -> 888-frame chunks between one-wave edges
    p = simlib.chunk_plan(5743, 444, 4608 * 2 * 2)
    sizes = [b - a for a, b in zip(p[:-1], p[1:])]
    assert p[0] == 0 and p[-1] == 5743 and sizes[0] == 444 and sizes[-1] == 444 and set(sizes[1:-2]) == {888}
    # config 4: 110 KB frames: one wave is already 49 MB, chunks are not doubled
    sizes = [b - a for a, b in zip(*(lambda q: (q[:-1], q[1:]))(simlib.chunk_plan(3125, 444, 4608 * 8 * 3)))]
    assert set(sizes[:-1]) == {444} and sizes[-1] < 2 * 444
    # config 3: waves of 148 big frames: two waves per chunk
    sizes = [b - a for a, b in zip(*(lambda q: (q[:-1], q[1:]))(simlib.chunk_plan(3516, 148, 16384 * 2 * 3)))]
    assert sizes[0] == 148 and sizes[-1] == 148 and 296 in sizes
    # LPC configurations: a third of a wave first, two thirds last (what flacb_encode_pcm asks for)
    p = simlib.chunk_plan(5743, 444, 4608 * 2 * 2, 0, 148, 296)
    sizes = [b - a for a, b in zip(p[:-1], p[1:])]
    assert p[0] == 0 and p[-1] == 5743 and sizes[0] == 148 and sizes[-1] == 296 and set(sizes[1:-2]) == {888}
    for n in (1, 2, 71, 443, 444, 887, 888, 889, 1333, 5743, 100001):
        for wave in (1, 148, 444):
            for pcm in (64, 18432, 110592):
                for want, edge, last in ((0, 0, 0), (5, 0, 0), (16, 0, 0), (100000, 0, 0), (888, 7, 0), (0, 49, 98), (0, 148, 296)):
                    q = simlib.chunk_plan(n, wave, pcm, want, edge, last)
                    assert q[0] == 0 and q[-1] == n and all(b > a for a, b in zip(q[:-1], q[1:])), (n, wave, pcm, want, edge, last, q[:6])


def test_data_model_records_match_the_reference_names_and_tables():
    """flac.common carries every record, enum and code table of the reference's data model (common.py:26-420)"""
    import flac.common as c
    for name in ("FrameHeader", "SubframeHeader", "SubframeTypeConstant", "SubframeTypeVerbatim", "SubframeTypeFixed",
                 "SubframeTypeLPC", "SubframeConstant", "SubframeVerbatim", "SubframeFixed", "SubframeLPC", "Frame",
                 "RicePartition", "EscapedPartition", "Residual", "RiceCodingMethod", "BlockSizeValue", "BlockSizeUncommon8",
                 "BlockSizeUncommon16", "SampleRateValue", "SampleRateFromStreaminfo", "SampleSizeValue",
                 "SAMPLE_RATE_VALUE_ENCODING", "SAMPLE_RATE_VALUE_DECODING", "SAMPLE_SIZE_ENCODING", "SAMPLE_SIZE_DECODING"):
        assert hasattr(c, name), name
    assert c.RiceCodingMethod.Rice4Bit.value == 4 and c.RiceCodingMethod.Rice5Bit.value == 5
    assert c.SAMPLE_RATE_VALUE_ENCODING[c.SampleRateValue.V_44_1_kHz] == 0b1001 and c.SampleRateValue.V_22_05_kHz.value == 22050
    assert c.SAMPLE_RATE_VALUE_DECODING[0b1100] is c.SampleRateValue.V_96_kHz
    assert c.SAMPLE_SIZE_ENCODING[c.SampleSizeValue.V_24] == 0b110 and 0b011 not in c.SAMPLE_SIZE_DECODING
    assert c.Channels.S_R.decorrelation_bit == [1, 0] and c.Channels.M_S.decorrelation_bit == [0, 1]
    assert c.Channels.FL_FR_FC_LFE_BL_BR.decorrelation_bit == [0] * 6
    lpc = c.SubframeLPC([1, 2, 3], 5, 7, [1, -2, 3], [0] * 10)
    assert lpc.order == 3 and repr(lpc) == "SubframeLPC(order=3, precision=5, shift=7, coefficients=[1, -2, 3])"
    assert repr(c.SubframeFixed([1], [2, 3])) == "SubframeFixed(order=1)"
    r = c.Residual(c.RiceCodingMethod.Rice4Bit, [c.RicePartition(3, [1, 2])] * 8)
    assert r.partition_order == 3 and repr(r.partitions[0]) == "RicePartition(parameter=3, samples_count=2)"
    h = c.FrameHeader(c.BlockingStrategy.Fixed, 4608, None, c.Channels.L_R, None, 7)
    assert h.crc is None and c.Frame(h, [], 0).header is h
