"""MSB-first bit reader / writer with the reference's interface (flac/binary.py): mask (:6),
extract (:19), Get (:78), Put (:144).  Used by the host side of the drop-in (stream prologue,
STREAMINFO parsing) and by the reference's own unit tests, which run unchanged against this
module.  The frames themselves are packed on the GPU (csrc/flacb200_kernels.cuh)."""
from __future__ import annotations

from io import BufferedIOBase


def mask(n: int) -> int:
    """n low bits set.

    >>> bin(mask(0))
    '0b0'
    >>> bin(mask(1))
    '0b1'
    >>> bin(mask(2))
    '0b11'
    >>> bin(mask(3))
    '0b111'
    """
    return ~(-1 << n)


def extract(x: int, size: int, start: int, stop: int) -> int:
    """Bits [start, stop) of the size-bit word x, counted from the most significant bit.

    >>> bin(extract(0b1, 1, 0, 1))
    '0b1'
    >>> bin(extract(0b10101010, 8, 0, 8))
    '0b10101010'
    >>> bin(extract(0b10101010, 8, 2, 5))
    '0b101'
    """
    return (x >> (size - stop)) & mask(stop - start)


def read(buffer: BufferedIOBase, n: int) -> bytes:
    data = buffer.read(n)
    if n < 1:
        raise ValueError("n must be greater than zero.")
    if len(data) != n:
        raise EOFError()
    return data


class _Binary:
    """Bit position inside the current byte, shared by Get and Put."""

    def __init__(self):
        self._bit_offset = 0

    @property
    def is_aligned(self) -> bool:
        return self._bit_offset == 0

    @property
    def bit_offset(self) -> int:
        return self._bit_offset

    @property
    def bits_until_alignment(self) -> int:
        return -self._bit_offset % 8


class Get(_Binary):
    """Reads big-endian bit fields from a binary stream, one byte at a time."""

    def __init__(self, buffer: BufferedIOBase):
        super().__init__()
        self._buffer = buffer
        self._held = 0  # the byte the bit offset points into

    def uint(self, n: int) -> int:
        value = 0
        todo = n
        while todo > 0:
            if self._bit_offset == 0:
                self._held = read(self._buffer, 1)[0]
            avail = 8 - self._bit_offset
            take = todo if todo < avail else avail
            value = (value << take) | extract(self._held, 8, self._bit_offset, self._bit_offset + take)
            self._bit_offset = (self._bit_offset + take) & 7
            todo -= take
        return value

    def sint(self, n: int) -> int:
        x = self.uint(n)
        return x - ((x >> (n - 1)) << n)

    def bool(self) -> bool:
        return self.uint(1) == 1

    def bytes(self, n: int) -> bytes:
        assert self._bit_offset == 0
        return read(self._buffer, n)


class Put(_Binary):
    """Appends big-endian bit fields; `buffer` returns the bytes once byte aligned."""

    def __init__(self):
        super().__init__()
        self._done = bytearray()
        self._acc = 0  # bits of the unfinished byte, left aligned count in _bit_offset

    @property
    def buffer(self) -> bytes:
        assert self._bit_offset == 0
        return bytes(self._done)

    def uint(self, x: int, n: int):
        if n == 0:
            return
        acc = (self._acc << n) | (x & mask(n))  # masking = two's complement for negatives
        nbits = self._bit_offset + n
        whole, rest = divmod(nbits, 8)
        if whole:
            self._done += (acc >> rest).to_bytes(whole, "big")
        self._acc = acc & mask(rest)
        self._bit_offset = rest

    def bool(self, x: bool):
        self.uint(1 if x is True else 0, 1)

    def bytes(self, bs: bytes):
        assert self._bit_offset == 0
        self._done += bs
        return len(bs)
