"""`flac.decoder` of the drop-in: same entry point as the reference (flac/decoder.py:31 decode).
The stream prologue (magic, STREAMINFO, other metadata blocks: decoder.py:34-44, :66-97) is
parsed on the host; every audio frame -- header, subframes, Rice residuals, prediction restore,
inter-channel decorrelation (decoder.py:100-498, binary.py:78-141) -- is decoded by
libflacb200.so on a B200, all frames in parallel."""
from __future__ import annotations

from io import BufferedIOBase
from typing import Iterator, Optional, Sequence

import numpy as np

from flac import _native as N
from flac.binary import Get
from flac.common import MAGIC, MetadataBlockHeader, MetadataBlockType, Streaminfo


def get_metadata_block_header(get: Get) -> MetadataBlockHeader:
    return MetadataBlockHeader(last=get.bool(), type=MetadataBlockType(get.uint(7)), length=get.uint(24))


def get_metadata_block_streaminfo(get: Get) -> Streaminfo:
    return Streaminfo(
        min_block_size=get.uint(16), max_block_size=get.uint(16),
        min_frame_size=get.uint(24), max_frame_size=get.uint(24),
        sample_rate=get.uint(20), channels=get.uint(3) + 1, sample_size=get.uint(5) + 1,
        samples=get.uint(36), md5=get.bytes(16))


def skip_metadata(get: Get):
    while True:
        header = get_metadata_block_header(get)
        get.bytes(header.length)
        if header.last is True:
            return


def decode_frames_status(frames, sample_size: int, *, devices: Optional[Sequence[int]] = None,
                         cap_sample_frames: Optional[int] = None) -> tuple[np.ndarray, int, int]:
    """Decode back-to-back audio frames (bytes-like) on the GPU.
    -> (int32 array (n, frame_channels), frame_channels, status): the samples of the frames in
    front of the first one the reference decoder fails in, and that failure's status."""
    raw = np.frombuffer(frames, dtype=np.uint8) if not isinstance(frames, np.ndarray) else frames
    ctx = N.default_context(devices[0] if devices else None)
    # Room for the decoded samples.  STREAMINFO's 36-bit sample count is only a hint (the reference never
    # allocates from it, and a wrong or hostile value must not turn into a terabyte of np.empty): it is
    # clamped to what a stream of this many bytes plausibly holds; when the guess is too small the library
    # says how much the stream decodes to and hands the frames it has already decoded over
    # (flacb_decode_output), so nothing is decoded twice and nothing is sized from untrusted data.
    guess = max(raw.size * 16, 1 << 16)
    cap = min((cap_sample_frames + 65536) * 8, guess) if cap_sample_frames else guess
    out = np.empty(cap, dtype=np.int32)
    rc, n, fch, _ = ctx.decode_raw(raw.ctypes.data if raw.size else 0, raw.size, sample_size, 0, out.ctypes.data, cap)
    if rc == N.E_NOSPACE:
        cap = n * max(fch, 1)
        out = np.empty(cap, dtype=np.int32)
        rc, n, fch, _ = ctx.decode_output(0, out.ctypes.data, cap)
    if rc in (N.E_CUDA, N.E_UNSUPPORTED, N.E_INTERNAL, N.E_NOSPACE):
        N.raise_for_status(rc, "decode")
    fch = max(fch, 1)
    return out[:n * fch].reshape(-1, fch), fch, rc


def scan_frames(frames, sample_size: int, *, devices: Optional[Sequence[int]] = None) -> tuple[np.ndarray, np.ndarray, int, int]:
    """The frame index of back-to-back audio frames (bytes-like): byte offsets of the frames the reference's
    get_frames loop (decoder.py:100-108) visits plus the end of the last one, their block sizes, the number of
    sample frames and the status the decode would end with.  Found in parallel on the GPU (every sync candidate
    is parsed, the real frames are the chain from offset 0); the reference finds frame k by parsing 0 .. k-1."""
    raw = np.frombuffer(frames, dtype=np.uint8) if not isinstance(frames, np.ndarray) else frames
    ctx = N.default_context(devices[0] if devices else None)
    cap = max(raw.size // 64, 64)
    while True:
        rc, offs, sizes, n, ns, _ = ctx.scan_frames(raw.ctypes.data if raw.size else 0, raw.size, sample_size, 0, cap)
        if rc != N.E_NOSPACE:
            break
        cap = n
    if rc in (N.E_CUDA, N.E_UNSUPPORTED, N.E_INTERNAL):
        N.raise_for_status(rc, "scan")
    return offs, sizes, ns, rc


def frame_headers(frames, sample_size: int, **kw) -> list:
    """`FrameHeader` records (flac.common, reference common.py:263-271) of the frames scan_frames() finds,
    parsed on the host from the few header bytes of each (decoder.py:133-185, :188-262)."""
    from flac import coded_number
    from flac.common import BlockingStrategy, CHANNELS_DECODING, FrameHeader, SAMPLE_RATE_VALUE_DECODING, SAMPLE_SIZE_DECODING
    offs, sizes, _, _ = scan_frames(frames, sample_size, **kw)
    raw = bytes(frames) if not isinstance(frames, (bytes, bytearray)) else frames
    out = []
    for k in range(len(offs) - 1):
        h = raw[int(offs[k]):int(offs[k]) + 16]
        rate_code, ss_code = h[2] & 15, (h[3] >> 1) & 7
        nbytes = coded_number.following_bytes(h[4]) + 1
        number = coded_number.decode(h[4:4 + nbytes])
        p = 4 + nbytes + {6: 1, 7: 2}.get(h[2] >> 4, 0)
        rate = None
        if rate_code in SAMPLE_RATE_VALUE_DECODING:
            rate = SAMPLE_RATE_VALUE_DECODING[rate_code].value
        elif rate_code in (12, 13, 14):
            w = 1 if rate_code == 12 else 2
            rate = int.from_bytes(h[p:p + w], "big") * (1000 if rate_code == 12 else 10 if rate_code == 14 else 1)
            p += w
        out.append(FrameHeader(BlockingStrategy(h[1] & 1), int(sizes[k]), rate, CHANNELS_DECODING[h[3] >> 4],
                               SAMPLE_SIZE_DECODING[ss_code].value if ss_code in SAMPLE_SIZE_DECODING else None, number, h[p]))
    return out


def decode_frames(frames, sample_size: int, *, devices: Optional[Sequence[int]] = None,
                  cap_sample_frames: Optional[int] = None) -> tuple[np.ndarray, int]:
    """-> (int32 array (n, frame_channels), frame_channels); raises what the reference raises."""
    pcm, fch, rc = decode_frames_status(frames, sample_size, devices=devices, cap_sample_frames=cap_sample_frames)
    N.raise_for_status(rc, "decode")
    return pcm, fch


def decode(buffer: BufferedIOBase) -> tuple[int, int, int, int, Iterator[list[int]]]:
    """Reference signature (flac/decoder.py:31-63)."""
    get = Get(buffer)
    assert get.bytes(4) == MAGIC
    header = get_metadata_block_header(get)
    assert header.type == MetadataBlockType.Streaminfo
    streaminfo = get_metadata_block_streaminfo(get)
    if header.last is False:
        skip_metadata(get)

    def samples() -> Iterator[list[int]]:
        # like the reference's generator: the sample frames of the good frames first, then the
        # exception of the frame that fails
        pcm, _, rc = decode_frames_status(buffer.read(), streaminfo.sample_size,
                                          cap_sample_frames=streaminfo.samples or None)
        for row in pcm.tolist():
            yield row
        N.raise_for_status(rc, "decode")

    return (streaminfo.sample_rate, streaminfo.sample_size, streaminfo.channels, streaminfo.samples, samples())


def decode_pcm(buffer: BufferedIOBase) -> tuple[Streaminfo, np.ndarray]:
    """Bulk entry used by the CLI: STREAMINFO and the decoded samples as one int32 array."""
    get = Get(buffer)
    assert get.bytes(4) == MAGIC
    header = get_metadata_block_header(get)
    assert header.type == MetadataBlockType.Streaminfo
    streaminfo = get_metadata_block_streaminfo(get)
    if header.last is False:
        skip_metadata(get)
    pcm, _ = decode_frames(buffer.read(), streaminfo.sample_size, cap_sample_frames=streaminfo.samples or None)
    return streaminfo, pcm
