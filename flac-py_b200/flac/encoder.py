"""`flac.encoder` of the drop-in: same entry points as the reference (flac/encoder.py:33
EncoderParameters, flac/encoder.py:48 encode), with the whole per-frame pipeline -- frame
header, fixed / LPC subframe analysis, Rice partition search, bit packing, CRC-8 / CRC-16
(flac/encoder.py:87-806, flac/binary.py:144-216, flac/crc.py, flac/coded_number.py) -- executed
by libflacb200.so on a B200.  Only the 42-byte stream prologue is produced on the host.

`encode()` keeps the reference's signature (an iterator of per-sample-frame lists in, an iterator
of byte chunks out: magic, metadata header, STREAMINFO, then one chunk per frame).  `encode_pcm()`
is the bulk entry the CLI uses: interleaved little-endian PCM bytes (a WAV payload) or an integer
array in, the same chunks out, optionally sharded by frame range over several GPUs.

Differences from the reference, all opt-in or unavoidable:
  * `lpc_order=range(1)` (max LPC order 0) means "fixed predictors only"; the reference crashes
    on it (encoder.py:404) and has no such switch.
  * `strict=False` encodes inputs on which the reference raises (digital silence, DC, ...).
  * `spec_channels=True` writes the real channel assignment instead of the reference's constant
    L_R nibble (encoder.py:95), which its own decoder needs for anything but stereo.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from itertools import islice
from typing import Iterator, Optional, Sequence

import numpy as np

from flac import _native as N
from flac.common import MAGIC


@dataclass
class EncoderParameters:
    block_size: int
    rice_partition_order: range
    lpc_order: range
    qlp_precision: int

    def __post_init__(self):
        assert self.lpc_order.start == 0
        assert self.lpc_order.stop <= 33  # Max LPC order is 32
        assert self.qlp_precision >= 5

    @property
    def max_lpc_order(self) -> int:
        return max(self.lpc_order.stop - 1, 0)


# frames handed to the GPU per call by the iterator entry point / bulk entry point
_ITER_CHUNK_BYTES = 64 << 20
_BULK_CHUNK_BYTES = 1 << 30


def _check(sample_rate: int, parameters: EncoderParameters):
    if sample_rate <= 48_000:
        assert parameters.lpc_order.stop <= 13
    r = parameters.rice_partition_order
    if len(r) == 0 or r.step != 1:
        raise NotImplementedError("rice_partition_order must be a non-empty range with step 1")


def stream_prologue(sample_rate: int, sample_size: int, channels: int, frames: int,
                    parameters: EncoderParameters) -> list[bytes]:
    """The three chunks the reference yields before the first frame (encoder.py:61-83)."""
    p = N.stream_prologue(sample_rate, channels, sample_size, frames, parameters.block_size)
    assert p[:4] == MAGIC
    return [p[:4], p[4:8], p[8:]]


def complete_streaminfo(streaminfo: bytes, frame_lengths, pcm_le: Optional[bytes] = None) -> bytes:
    """STREAMINFO with the fields the reference leaves at zero (encoder.py:75-81): minimum / maximum
    frame size in bytes (24 bits each, from the frame lengths the encode returns anyway) and, if
    the interleaved little-endian PCM is given, its MD5 (16 bytes, as the FLAC format defines
    it).  `streaminfo` is the 34-byte block (the third chunk of `stream_prologue`).  An
    extension: the reference never writes these, so a stream patched with this differs from the
    reference's in exactly these 22 bytes."""
    import hashlib
    assert len(streaminfo) == 34
    lens = [int(v) for v in frame_lengths]
    out = bytearray(streaminfo)
    if lens:
        lo, hi = min(lens), max(lens)
        assert hi < (1 << 24)
        out[4:7] = lo.to_bytes(3, "big")
        out[7:10] = hi.to_bytes(3, "big")
    if pcm_le is not None:
        out[18:34] = hashlib.md5(pcm_le).digest()
    return bytes(out)


def seektable_block(frame_offsets, block_size: int, total_samples: int, every: int = 10,
                    last: bool = True) -> bytes:
    """A SEEKTABLE metadata block (type 3, common.py:30) from the frame byte offsets the encode
    returns anyway: one seek point per `every` frames (first sample of the frame, byte offset from
    the first frame header, samples in the frame).  An extension:
    the reference never writes one, its decoder skips it (decoder.py:43-44, :92-97), and so does
    this package's.  To use it, clear the "last" bit of the STREAMINFO header (first byte of the
    second chunk of `stream_prologue`: 0x80 -> 0x00) and put this block in front of the frames."""
    offs = [int(v) for v in frame_offsets]
    n_frames = len(offs) - 1
    points = bytearray()
    for f in range(0, n_frames, max(1, int(every))):
        first = f * block_size
        count = min(block_size, total_samples - first)
        points += first.to_bytes(8, "big") + offs[f].to_bytes(8, "big") + count.to_bytes(2, "big")
    assert len(points) < (1 << 24)
    header = bytes([(0x80 if last else 0x00) | 3]) + len(points).to_bytes(3, "big")
    return header + bytes(points)


class _Shard:
    """Frames [f0, f1) of one call on one device."""

    def __init__(self, ctx: N.Context, f0: int, f1: int):
        self.ctx, self.f0, self.f1 = ctx, f0, f1
        self.rc = N.OK
        self.out: Optional[np.ndarray] = None
        self.offs: Optional[np.ndarray] = None
        self.stats: Optional[N.Stats] = None
        self.info = None
        self.error: Optional[BaseException] = None


def _run_shard(sh: _Shard, raw: np.ndarray, n_sample_frames: int, channels: int, bits: int,
               p: EncoderParameters, first_frame_number: int, flags: int, want_info: bool, copy_out: bool = True):
    try:
        bs = p.block_size
        cbytes = 4 if flags & N.PCM_INT32 else (bits + 7) // 8
        s0 = sh.f0 * bs
        s1 = min(sh.f1 * bs, n_sample_frames)
        n = s1 - s0
        nfr = sh.f1 - sh.f0
        rmin, rmax = p.rice_partition_order.start, p.rice_partition_order.stop - 1
        cap = N.encode_bound(n, channels, bits, bs, rmax)
        offs = np.zeros(nfr + 1, dtype=np.uint64)
        info = (N.SubframeInfo * max(nfr * channels, 1))() if want_info else None
        part = raw[s0 * channels * cbytes:s1 * channels * cbytes]
        # through the context's page-locked staging memory: one host copy in, and the stream is read
        # out of page-locked memory (the library overlaps the copies of its chunks with its kernels)
        pin_in = sh.ctx.staging("in", part.size)
        pin_out = sh.ctx.staging("out", cap)
        np.copyto(pin_in.array[:part.size], part)
        rc, out_len, st = sh.ctx.encode_raw(pin_in.ptr, n, channels, bits, bs, p.max_lpc_order,
                                            p.qlp_precision, rmin, rmax, first_frame_number + sh.f0,
                                            flags, pin_out.ptr, cap, offs, info)
        out = pin_out.array[:out_len]
        sh.rc, sh.out, sh.offs, sh.stats, sh.info = rc, (out.copy() if copy_out else out), offs, st, info
    except BaseException as e:  # noqa: BLE001
        sh.error = e


def _as_raw(pcm, channels: int, sample_size: int) -> tuple[np.ndarray, int, int]:
    """-> (uint8 view, n_sample_frames, extra flags)"""
    if isinstance(pcm, np.ndarray) and pcm.dtype != np.uint8:
        a = np.ascontiguousarray(pcm, dtype=np.int32).reshape(-1, channels)
        return a.view(np.uint8).reshape(-1), a.shape[0], N.PCM_INT32
    raw = np.frombuffer(pcm, dtype=np.uint8) if not isinstance(pcm, np.ndarray) else pcm.reshape(-1)
    width = (sample_size + 7) // 8
    assert raw.size % (width * channels) == 0, "PCM length is not a whole number of sample frames"
    return raw, raw.size // (width * channels), 0


def encode_frames(sample_size: int, channels: int, pcm, parameters: EncoderParameters, *,
                  first_frame_number: int = 0, strict: bool = True, spec_channels: bool = False,
                  devices: Optional[Sequence[int]] = None, want_info: bool = False,
                  flags: int = 0, copy_out: bool = True):
    """Encode the frames of `pcm` (no prologue).  Returns (data: np.uint8 array per shard list,
    frame chunks iterator helper data).  Internal building block of encode_pcm / encode."""
    raw, n, f_in = _as_raw(pcm, channels, sample_size)
    flags |= f_in
    if not strict:
        flags |= N.LENIENT
    if spec_channels:
        flags |= N.SPEC_CHANNELS
    bs = parameters.block_size
    assert bs >= 1
    nfr = (n + bs - 1) // bs
    if devices is None:
        ctxs = [N.default_context()]
    else:
        ctxs = [N.default_context(d) for d in devices]
    g = max(1, min(len(ctxs), nfr))
    shards = [_Shard(ctxs[k], nfr * k // g, nfr * (k + 1) // g) for k in range(g)]
    # copy_out=False: a shard's `out` is a view of its context's staging memory, valid until the next call
    args = (raw, n, channels, sample_size, parameters, first_frame_number, flags, want_info, copy_out)
    if g == 1:
        _run_shard(shards[0], *args)
    else:
        ths = [threading.Thread(target=_run_shard, args=(sh,) + args) for sh in shards]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    return shards


def subframe_records(shard: _Shard, precision: int) -> list:
    """What the kernels decided, as the reference's records (flac.common, reference common.py:306-420):
    one (SubframeHeader, Residual-shaped summary) pair per subframe of a shard encoded with want_info=True.
    The records carry the decisions (type, order, shift, coefficients, coding method, Rice parameters), not
    the samples: warm-up and residual lists hold `None` placeholders of the right length."""
    from flac.common import (Residual, RiceCodingMethod, RicePartition, SubframeFixed, SubframeHeader, SubframeLPC,
                             SubframeTypeFixed, SubframeTypeLPC)
    out = []
    for inf in shard.info or []:
        lpc = bool(inf.type)
        header = SubframeHeader(SubframeTypeLPC(inf.order) if lpc else SubframeTypeFixed(inf.order), 0)
        warm = [None] * inf.order
        if lpc:
            m = inf.order - 1
            body = SubframeLPC(warm, precision, inf.shift, list(inf.lpc_q[m][:inf.lpc_ncoef[m]]), [])
        else:
            body = SubframeFixed(warm, [])
        parts = [RicePartition(int(k), []) for k in inf.params[:min(inf.nparams, len(inf.params))]]
        parts += [RicePartition(-1, [])] * (inf.nparams - len(parts))  # beyond what the record holds
        out.append((header, body, Residual(RiceCodingMethod(inf.method), parts)))
    return out


def _yield_shards(shards: list[_Shard]) -> Iterator[bytes]:
    """One chunk per frame in stream order; raises where the reference raises, after the frames
    before the failing one have been yielded (the reference's generator behaves the same)."""
    for sh in shards:
        if sh.error is not None:
            raise sh.error
        data, offs = sh.out, sh.offs
        nfr = sh.f1 - sh.f0
        good = nfr if sh.rc == N.OK else max(int(sh.stats.first_bad_frame), 0)
        mv = memoryview(data)
        for f in range(good):
            yield bytes(mv[int(offs[f]):int(offs[f + 1])])
        if sh.rc != N.OK:
            N.raise_for_status(sh.rc, f"frame {sh.f0 + good}")


def encode_pcm(sample_rate: int, sample_size: int, channels: int, pcm, parameters: EncoderParameters,
               *, frames: Optional[int] = None, strict: bool = True, spec_channels: bool = False,
               devices: Optional[Sequence[int]] = None) -> Iterator[bytes]:
    """Bulk entry: `pcm` is the interleaved little-endian WAV payload (bytes-like, width
    ceil(sample_size / 8)) or an integer array of shape (n, channels).  Yields exactly the chunks
    `encode()` yields."""
    _check(sample_rate, parameters)
    raw, n, f_in = _as_raw(pcm, channels, sample_size)
    yield from stream_prologue(sample_rate, sample_size, channels, n if frames is None else frames, parameters)
    bs = parameters.block_size
    cbytes = 4 if f_in else (sample_size + 7) // 8
    per_call = max(1, _BULK_CHUNK_BYTES * max(1, len(devices or [0])) // (bs * channels * cbytes)) * bs
    done = 0
    while done < n:
        m = min(per_call, n - done)
        part = raw[done * channels * cbytes:(done + m) * channels * cbytes]
        shards = encode_frames(sample_size, channels, part, parameters, first_frame_number=done // bs,
                               strict=strict, spec_channels=spec_channels, devices=devices, flags=f_in, copy_out=False)
        yield from _yield_shards(shards)  # every frame is copied out of the staging memory here, before the next call
        done += m


def encode(
        sample_rate: int,
        sample_size: int,
        channels: int,
        frames: int,
        samples: Iterator[list[int]],
        parameters: EncoderParameters,
        *,
        strict: bool = True,
        spec_channels: bool = False,
) -> Iterator[bytes]:
    """Reference signature (flac/encoder.py:48-55).  The sample iterator is consumed a few
    thousand blocks at a time; each batch is one GPU call."""
    _check(sample_rate, parameters)
    yield from stream_prologue(sample_rate, sample_size, channels, frames, parameters)
    bs = parameters.block_size
    blocks = max(1, _ITER_CHUNK_BYTES // (bs * channels * 4))
    number = 0
    it = iter(samples)
    while True:
        rows = list(islice(it, blocks * bs))
        if not rows:
            return
        a = np.asarray(rows, dtype=np.int64)
        if a.ndim != 2 or a.shape[1] < channels:
            raise IndexError("list index out of range")  # x[c] in encoder.py:102
        a = np.ascontiguousarray(a[:, :channels]).astype(np.int32)
        shards = encode_frames(sample_size, channels, a, parameters, first_frame_number=number,
                               strict=strict, spec_channels=spec_channels)
        yield from _yield_shards(shards)
        number += (len(rows) + bs - 1) // bs
        if len(rows) < blocks * bs:
            return
