"""Stream constants and the few records the host side of the drop-in uses, named as in the
reference's data model (flac/common.py): MAGIC (:9), FRAME_SYNC_CODE (:10), CRC polynomials
(:12-13), FIXED_PREDICTOR_COEFFICIENTS (:15-21), MetadataBlockType (:26), MetadataBlockHeader
(:37), Streaminfo (:46), BlockingStrategy (:60), BLOCK_SIZE_ENCODING (:85), Channels (:168) and
its code / count tables (:199-230), RiceCodingMethod (:401).

The reference materialises every frame as a tree of Python dataclasses (FrameHeader, Subframe*,
Residual, RicePartition ...) before writing it; here frames only ever exist as bytes produced on
the GPU, so those per-frame records have no counterpart.  What the kernels decided per subframe
can be inspected through `flac.encoder.encode_pcm(..., want_info=True)`."""
from __future__ import annotations

from dataclasses import dataclass
from enum import auto

from flac.utils import Enum, invert_dict

MAGIC = b"fLaC"
FRAME_SYNC_CODE = 0b111111111111100
CRC8_POLYNOMIAL = 0b100000111
CRC16_POLYNOMIAL = 0b11000000000000101

# row o: x[i] is predicted as sum(c[j] * x[i-1-j])
FIXED_PREDICTOR_COEFFICIENTS = ((), (1,), (2, -1), (3, -3, 1), (4, -6, 4, -1))


class MetadataBlockType(Enum):
    Streaminfo = 0
    Padding = 1
    Application = 2
    Seektable = 3
    VorbisComment = 4
    Cuesheet = 5
    Picture = 6


@dataclass(frozen=True)
class MetadataBlockHeader:
    last: bool
    type: MetadataBlockType
    length: int


@dataclass(frozen=True)
class Streaminfo:
    min_block_size: int
    max_block_size: int
    min_frame_size: int
    max_frame_size: int
    sample_rate: int
    channels: int
    sample_size: int
    samples: int
    md5: bytes


class BlockingStrategy(Enum):
    Fixed = 0
    Variable = 1


class RiceCodingMethod(Enum):
    Rice4Bit = 0
    Rice5Bit = 1


# block size -> 4-bit frame header code; anything else is written as 0110 + 8 bits or
# 0111 + 16 bits of (size - 1)
BLOCK_SIZE_ENCODING = {192: 0b0001}
BLOCK_SIZE_ENCODING.update({576 << k: 0b0010 + k for k in range(4)})
BLOCK_SIZE_ENCODING.update({256 << k: 0b1000 + k for k in range(8)})


class Channels(Enum):
    M = auto()
    L_R = auto()
    L_R_C = auto()
    FL_FR_BL_BR = auto()
    FL_FR_FC_BL_BR = auto()
    FL_FR_FC_LFE_BL_BR = auto()
    FL_FR_FC_LFE_BC_SL_SR = auto()
    FL_FR_FC_LFE_BL_BR_SL_SR = auto()
    L_S = auto()
    S_R = auto()
    M_S = auto()

    @property
    def count(self) -> int:
        return CHANNELS_COUNT[self]


CHANNELS_ENCODING = {c: code for code, c in enumerate(Channels)}
CHANNELS_DECODING = invert_dict(CHANNELS_ENCODING)
CHANNELS_COUNT = {c: (code + 1 if code < 8 else 2) for c, code in CHANNELS_ENCODING.items()}
