"""UTF-8 style variable-length numbers of FLAC frame headers, with the reference's interface
(flac/coded_number.py): encode (:7), required_bytes (:24), decode (:44), following_bytes (:58).
The GPU writes the coded frame number itself (build_frame_header in csrc/flacb200_kernels.cuh);
this module serves the host side and the reference's unit tests."""
from __future__ import annotations

# (largest bit length, total bytes)
_SIZES = ((7, 1), (11, 2), (16, 3), (21, 4), (26, 5), (31, 6))


def required_bytes(x: int) -> int:
    n = x.bit_length()
    for limit, size in _SIZES:
        if n <= limit:
            return size
    raise ValueError(f"Cannot encode coded number: {x}")


def encode(x: int) -> bytes:
    assert 0 <= x.bit_length() <= 36
    size = required_bytes(x)
    if size == 1:
        return bytes([x])
    out = bytearray(size)
    for k in range(size - 1, 0, -1):  # continuation bytes, least significant group last
        out[k] = 0x80 | (x & 0x3F)
        x >>= 6
    out[0] = ((0xFF << (8 - size)) & 0xFF) | x
    return bytes(out)


def following_bytes(x: int) -> int:
    """Number of continuation bytes announced by the first byte."""
    ones = 0
    while ones < 7 and x & (0x80 >> ones):
        ones += 1
    return ones - 1 if ones >= 2 else 0


def decode(bs: bytes) -> int:
    size = following_bytes(bs[0]) + 1
    assert size == len(bs)
    if size == 1:
        return bs[0]
    value = bs[0] & (0xFF >> (size + 1))
    for b in bs[1:]:
        value = (value << 6) | (b & 0x3F)
    return value
