"""CRC-8 / CRC-16 with the reference's interface (flac/crc.py): crc_table (:4), crc8 (:17),
crc16 (:25) -- MSB first, no reflection, no final xor.  Host-side helper (tests, tools); frame
CRCs are computed on the GPU (build_frame_header / flush_window in csrc/flacb200_kernels.cuh)."""
from __future__ import annotations

from functools import lru_cache


@lru_cache(maxsize=None)
def crc_table(n: int, generator: int) -> list[int]:
    poly = (1 << n) | generator
    top = 1 << n
    table = []
    for byte in range(256):
        reg = byte << (n - 8)
        for _ in range(8):
            reg <<= 1
            if reg & top:
                reg ^= poly
        table.append(reg)
    return table


def _crc(bs: bytes, n: int, generator: int, initial_value: int) -> int:
    table = crc_table(n, generator)
    keep = (1 << n) - 1
    reg = initial_value
    for b in bs:
        reg = ((reg << 8) & keep) ^ table[((reg >> (n - 8)) ^ b) & 0xFF]
    return reg


def crc8(bs: bytes, generator: int, initial_value: int = 0) -> int:
    return _crc(bs, 8, generator, initial_value)


def crc16(bs: bytes, generator: int, initial_value: int = 0) -> int:
    return _crc(bs, 16, generator, initial_value)
