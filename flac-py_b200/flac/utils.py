"""Small helpers with the reference's names (flac/utils.py): argparse_range (:12), batch (:31),
clamp (:43), group (:61), invert_dict (:69), log2i (:73), strictly_increasing (:83),
zigzag_decode (:87), zigzag_encode (:91), Enum (:97).  Host-side only; the GPU path does not
use them, the CLI and the tests do."""
from __future__ import annotations

from enum import Enum as _Enum
from typing import Iterator, TypeVar

T = TypeVar("T")


def strictly_increasing(xs: list[int]) -> bool:
    return all(b > a for a, b in zip(xs, xs[1:]))


def argparse_range(s: str) -> range:
    """'[min,]max' -> range(min, max + 1)

    >>> argparse_range('5')
    range(0, 6)
    >>> argparse_range('2,5')
    range(2, 6)
    """
    parts = [int(tok) for tok in s.split(",")]
    assert len(parts) in (1, 2)
    assert strictly_increasing(parts)
    lo = parts[0] if len(parts) == 2 else 0
    return range(lo, parts[-1] + 1)


def batch(it: Iterator[T], n: int) -> Iterator[list[T]]:
    """Lists of n consecutive items; the last one may be shorter.

    >>> [x for x in batch(iter('ABCDEFG'), 3)]
    [['A', 'B', 'C'], ['D', 'E', 'F'], ['G']]
    """
    if n < 1:
        raise ValueError("n must be greater than zero")
    chunk: list[T] = []
    for item in it:
        chunk.append(item)
        if len(chunk) == n:
            yield chunk
            chunk = []
    if chunk:
        yield chunk


def clamp(x: int, a: int, b: int) -> int:
    return a if x < a else b if x > b else x


def group(xs, n: int):
    """
    >>> group([1, 2, 3, 4, 5, 6], 2)
    [[1, 2], [3, 4], [5, 6]]
    """
    return [xs[k:k + n] for k in range(0, len(xs), n)]


def invert_dict(d: dict) -> dict:
    return dict((v, k) for k, v in d.items())


def log2i(x: int) -> int:
    "Base 2 integer logarithm"
    assert x > 0
    return x.bit_length() - 1


def zigzag_decode(x: int) -> int:
    return -((x + 1) >> 1) if x & 1 else x >> 1


def zigzag_encode(x: int) -> int:
    assert (-1 << 63) < x < (1 << 64) - 1
    return 2 * x if x >= 0 else -2 * x - 1


class Enum(_Enum):
    @classmethod
    def values(cls):
        return {m.value for m in cls}
