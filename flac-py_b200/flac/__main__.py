"""`python -m flac {encode,decode}`: the reference's command line (flac/__main__.py:114-212) --
same sub-commands, options, defaults and timing messages -- in front of the B200 path.

WAV I/O is done in bulk (the reference reads and writes one sample frame at a time, and its
sample parser is only correct when channels == bytes per sample, flac/__main__.py:84-92; here
every width is parsed as little-endian PCM of that width)."""
from __future__ import annotations

from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser
from pathlib import Path
from timeit import default_timer as timer
from wave import open as wave_open

import numpy as np

from flac.decoder import decode_pcm
from flac.encoder import EncoderParameters, encode_pcm
from flac.utils import argparse_range

ACTION_ENCODE = "encode"
ACTION_DECODE = "decode"

DEFAULT_BLOCK_SIZE = 4608
DEFAULT_MAX_LPC_ORDER = 12
DEFAULT_QLP_COEFF_PRECISION = 5
DEFAULT_RICE_PARTITION_ORDER = "5"


def _pcm_bytes(pcm: np.ndarray, width: int) -> bytes:
    """int32 (n, channels) -> interleaved little-endian PCM of `width` bytes per sample"""
    if width == 2:
        return pcm.astype("<i2").tobytes()
    if width == 4:
        return pcm.astype("<i4").tobytes()
    if width == 1:
        return pcm.astype("i1").tobytes()
    return np.ascontiguousarray(pcm.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :width]).tobytes()


def cmd_decode(path_in: Path, path_out: Path):
    with path_in.open("rb") as f, wave_open(str(path_out), mode="wb") as w:
        time_start = timer()
        streaminfo, pcm = decode_pcm(f)
        # Python's wave module only handles whole-byte sample sizes (flac/__main__.py:34-36)
        assert streaminfo.sample_size % 8 == 0
        w.setframerate(streaminfo.sample_rate)
        w.setsampwidth(streaminfo.sample_size // 8)
        w.setnchannels(pcm.shape[1] if pcm.size else streaminfo.channels)
        w.setnframes(streaminfo.samples)
        w.writeframes(_pcm_bytes(pcm, streaminfo.sample_size // 8))
        time_end = timer()
        delta = "{0:.6g}".format(time_end - time_start)
        print(f"Decoding completed in {delta} seconds")


def cmd_encode(path_in: Path, path_out: Path, block_size: int, max_lpc_order: int,
               qlp_coeffs_precision: int, rice_partition_order: range, complete_streaminfo: bool = False):
    with wave_open(str(path_in), mode="rb") as f_in, path_out.open("wb") as f_out:
        parameters = EncoderParameters(
            block_size=block_size,
            lpc_order=range(max_lpc_order + 1),
            qlp_precision=qlp_coeffs_precision,
            rice_partition_order=rice_partition_order,
        )
        sample_rate = f_in.getframerate()
        sample_size_bytes = f_in.getsampwidth()
        channels = f_in.getnchannels()
        frames = f_in.getnframes()
        time_start = timer()
        payload = f_in.readframes(frames)
        lengths = []
        for bs in encode_pcm(sample_rate, sample_size_bytes * 8, channels, payload, parameters, frames=frames):
            f_out.write(bs)
            lengths.append(len(bs))
        if complete_streaminfo:
            # not in the reference: frame sizes and the MD5 of the PCM, which the reference leaves at
            # zero (encoder.py:75-81), patched into the STREAMINFO block at byte 8 of the stream
            from .encoder import complete_streaminfo as _complete, stream_prologue
            block = stream_prologue(sample_rate, sample_size_bytes * 8, channels, frames, parameters)[2]
            f_out.seek(8)
            f_out.write(_complete(block, lengths[3:], payload))
        time_end = timer()
        delta = "{0:.6g}".format(time_end - time_start)
        print(f"Encoding completed in {delta} seconds")


def make_argument_parser():
    parser = ArgumentParser(prog="flac", formatter_class=ArgumentDefaultsHelpFormatter)
    action = parser.add_subparsers(title="action", dest="action", required=True)

    decode = action.add_parser(ACTION_DECODE, formatter_class=ArgumentDefaultsHelpFormatter)
    decode.add_argument("infile", type=Path, metavar="infile.flac")
    decode.add_argument("outfile", type=Path, metavar="outfile.wav")

    encode = action.add_parser(ACTION_ENCODE, formatter_class=ArgumentDefaultsHelpFormatter)
    encode.add_argument("infile", type=Path, metavar="infile.wav")
    encode.add_argument("outfile", type=Path, metavar="outfile.flac")
    encode.add_argument(
        "-b", "--block-size", type=int, default=DEFAULT_BLOCK_SIZE, metavar="N",
        help="Blocksize in samples. For subset streams this must be <= 4608 if the samplerate <= 48kHz. "
             "For  subset streams with higher samplerates it must be <= 16384.")
    encode.add_argument(
        "-l", "--max-lpc-order", type=int, default=DEFAULT_MAX_LPC_ORDER, metavar="N",
        help="Specifies  the  maximum LPC order. This number must be <= 32. For subset streams, it must be "
             "<= 12 if the sample rate is <= 48kHz. (0 = fixed predictors only, an extension.)")
    encode.add_argument(
        "-q", "--qlp-coeff-precision", type=int, default=DEFAULT_QLP_COEFF_PRECISION, metavar="N",
        help="Precision of the quantized linear-predictor coefficients. (min is 5)")
    encode.add_argument(
        "-r", "--rice-partition-order", type=argparse_range, default=DEFAULT_RICE_PARTITION_ORDER,
        metavar="[M,]N",
        help="[min,]max residual partition order (0..15). min defaults to 0 if unspecified.")
    encode.add_argument(
        "--complete-streaminfo", action="store_true",
        help="Extension (not in the reference): write the minimum / maximum frame size and the MD5 of the "
             "PCM into STREAMINFO instead of the reference's zeros.")
    return parser


def main():
    args = make_argument_parser().parse_args()
    if args.action == ACTION_DECODE:
        cmd_decode(args.infile, args.outfile)
    if args.action == ACTION_ENCODE:
        rpo = args.rice_partition_order
        if isinstance(rpo, str):  # argparse does not run `type` on a string default
            rpo = argparse_range(rpo)
        cmd_encode(args.infile, args.outfile, args.block_size, args.max_lpc_order,
                   args.qlp_coeff_precision, rpo, args.complete_streaminfo)


if __name__ == "__main__":
    main()
