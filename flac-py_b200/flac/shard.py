"""Frame-range sharding of one stream over several GPUs / ranks.

Frames share no state except their index (reference flac/encoder.py:87, :97: the coded frame
number), so rank r of R encodes frames [F*r/R, F*(r+1)/R) with `first_frame_number` set to the
start of its range, and the stream is the 42-byte prologue followed by the ranks' byte ranges
in order.  No data-path collective: the only thing ranks exchange is one integer each (their
byte count), from which every rank knows where its bytes belong in the output file."""
from __future__ import annotations

from typing import Sequence


def frame_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced, order-preserving split."""
    assert 0 <= rank < world
    return n_frames * rank // world, n_frames * (rank + 1) // world


def sample_range(n_sample_frames: int, block_size: int, rank: int, world: int) -> tuple[int, int]:
    n_frames = (n_sample_frames + block_size - 1) // block_size
    f0, f1 = frame_range(n_frames, rank, world)
    return f0 * block_size, min(f1 * block_size, n_sample_frames)


def byte_offsets(sizes: Sequence[int], prologue: int = 42) -> list[int]:
    """File offset of every rank's byte range (plus the total as the last entry)."""
    out = [prologue]
    for s in sizes:
        out.append(out[-1] + int(s))
    return out
