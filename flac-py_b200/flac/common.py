"""Stream constants and records of the drop-in, named as in the reference's data model
(flac/common.py): MAGIC (:9), FRAME_SYNC_CODE (:10), CRC polynomials (:12-13),
FIXED_PREDICTOR_COEFFICIENTS (:15-21), MetadataBlockType (:26), MetadataBlockHeader (:37),
Streaminfo (:46), BlockingStrategy (:60), block size records and code table (:67-105), sample rate
records and tables (:110-163), Channels and its tables (:168-230), sample size records and tables
(:235-258), FrameHeader (:263), SubframeType* (:276-299), SubframeHeader (:306), Subframe* (:314-361),
Frame (:367), RicePartition / EscapedPartition (:378-396), RiceCodingMethod (:401), Residual (:406).

The reference materialises every frame as a tree of these records before it writes it; here the
frames are produced as bytes on the GPU and the records are never on the hot path.  They exist so
that code written against `flac.common` keeps importing and constructing them, and they are what
`flac.decoder.frame_headers()` (the parallel sync scan, flacb_scan_frames) and
`flac.encoder.subframe_records()` (the kernels' per-subframe decisions) hand back."""
from __future__ import annotations

from dataclasses import dataclass
from enum import auto
from typing import Optional, Sequence, Union

from flac.utils import Enum, invert_dict, log2i

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
    """width of the Rice parameter field (the value is the width, as in the reference)"""
    Rice4Bit = 4
    Rice5Bit = 5


# block size -> 4-bit frame header code; anything else is written as 0110 + 8 bits or
# 0111 + 16 bits of (size - 1)
BLOCK_SIZE_ENCODING = {192: 0b0001}
BLOCK_SIZE_ENCODING.update({576 << k: 0b0010 + k for k in range(4)})
BLOCK_SIZE_ENCODING.update({256 << k: 0b1000 + k for k in range(8)})


@dataclass(frozen=True)
class BlockSizeValue:
    size: int


@dataclass(frozen=True)
class BlockSizeUncommon8:
    """0110: the size minus one follows the coded number in 8 bits"""


@dataclass(frozen=True)
class BlockSizeUncommon16:
    """0111: the size minus one follows the coded number in 16 bits"""


BlockSize = Union[BlockSizeValue, BlockSizeUncommon8, BlockSizeUncommon16]


class SampleRateFromStreaminfo:
    pass


# the rates the 4-bit header field names, in code order 0001 .. 1010, then 96 kHz at 1100
_RATE_CODES = ((88_200, 0b0001), (176_400, 0b0010), (192_000, 0b0011), (8_000, 0b0100), (16_000, 0b0101), (22_050, 0b0110),
               (24_000, 0b0111), (32_000, 0b1000), (44_100, 0b1001), (48_000, 0b1010), (96_000, 0b1100))


def _rate_name(hz: int) -> str:
    khz = hz / 1000
    return "V_" + (f"{khz:g}".replace(".", "_")) + "_kHz"


SampleRateValue = Enum("SampleRateValue", {_rate_name(hz): hz for hz, _ in _RATE_CODES})


class SampleRateUncommon8:
    pass


class SampleRateUncommon16:
    pass


class SampleRateUncommon16_10:
    pass


SampleRate = Union[SampleRateFromStreaminfo, SampleRateValue, SampleRateUncommon8, SampleRateUncommon16, SampleRateUncommon16_10]
SAMPLE_RATE_VALUE_ENCODING = {SampleRateValue(hz): code for hz, code in _RATE_CODES}
SAMPLE_RATE_VALUE_DECODING = invert_dict(SAMPLE_RATE_VALUE_ENCODING)


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

    @property
    def decorrelation_bit(self) -> list[int]:
        """extra sample-size bit per subframe: the side channel of L/S, S/R and M/S carries one"""
        if self is Channels.S_R:
            return [1, 0]
        if self in (Channels.L_S, Channels.M_S):
            return [0, 1]
        return [0] * self.count


CHANNELS_ENCODING = {c: code for code, c in enumerate(Channels)}
CHANNELS_DECODING = invert_dict(CHANNELS_ENCODING)
CHANNELS_COUNT = {c: (code + 1 if code < 8 else 2) for c, code in CHANNELS_ENCODING.items()}


@dataclass(frozen=True)
class SampleSizeFromStreaminfo:
    pass


SampleSizeValue = Enum("SampleSizeValue", {f"V_{b}": b for b in (8, 12, 16, 20, 24, 32)})
SampleSize = Union[SampleSizeFromStreaminfo, SampleSizeValue]
# 3-bit header field: 001 8, 010 12, (011 reserved), 100 16, 101 20, 110 24, 111 32
SAMPLE_SIZE_ENCODING = {SampleSizeValue(b): code for b, code in ((8, 1), (12, 2), (16, 4), (20, 5), (24, 6), (32, 7))}
SAMPLE_SIZE_DECODING = invert_dict(SAMPLE_SIZE_ENCODING)


@dataclass(frozen=True)
class FrameHeader:
    blocking_strategy: BlockingStrategy
    block_size: int
    sample_rate: Optional[int]  # None: as in STREAMINFO
    channels: Channels
    sample_size: Optional[int]  # None: as in STREAMINFO
    coded_number: int
    crc: Optional[int] = None


@dataclass(frozen=True)
class SubframeTypeConstant:
    pass


@dataclass(frozen=True)
class SubframeTypeVerbatim:
    pass


@dataclass(frozen=True)
class SubframeTypeFixed:
    order: int


@dataclass(frozen=True)
class SubframeTypeLPC:
    order: int


SubframeType = Union[SubframeTypeConstant, SubframeTypeVerbatim, SubframeTypeFixed, SubframeTypeLPC]


@dataclass(frozen=True)
class SubframeHeader:
    type_: SubframeType
    wasted_bits: int


@dataclass(frozen=True, repr=False)
class SubframeConstant:
    sample: int
    block_size: int  # not in the stream: how many samples the constant stands for

    def __repr__(self):
        return "SubframeConstant()"


@dataclass(frozen=True, repr=False)
class SubframeVerbatim:
    samples: list[int]

    def __repr__(self):
        return "SubframeVerbatim()"


@dataclass(frozen=True, repr=False)
class SubframeFixed:
    warmup: list[int]
    residual: list[int]

    @property
    def order(self) -> int:
        return len(self.warmup)

    def __repr__(self):
        return f"SubframeFixed(order={self.order})"


@dataclass(frozen=True, repr=False)
class SubframeLPC:
    warmup: list[int]
    precision: int
    shift: int
    coefficients: list[int]
    residual: list[int]

    @property
    def order(self) -> int:
        return len(self.warmup)

    def __repr__(self):
        return f"SubframeLPC(order={self.order}, precision={self.precision}, shift={self.shift}, coefficients={self.coefficients})"


Subframe = Union[SubframeConstant, SubframeVerbatim, SubframeFixed, SubframeLPC]


@dataclass(frozen=True)
class Frame:
    header: FrameHeader
    subframes: list[Subframe]
    crc: int


@dataclass(frozen=True, repr=False)
class RicePartition:
    parameter: int
    residual: list[int]  # zig-zag encoded

    def __repr__(self):
        return f"RicePartition(parameter={self.parameter}, samples_count={len(self.residual)})"


@dataclass(frozen=True, repr=False)
class EscapedPartition:
    residual: list[int]

    def __repr__(self):
        return f"EscapedPartition(samples_count={len(self.residual)})"


ResidualPartition = Union[RicePartition, EscapedPartition]


@dataclass(frozen=True, repr=False)
class Residual:
    coding_method: RiceCodingMethod
    partitions: Sequence[ResidualPartition]

    @property
    def partition_order(self) -> int:
        return log2i(len(self.partitions))

    def __repr__(self):
        return f"Residual(coding_method={self.coding_method}, partition_order={self.partition_order}, partitions={self.partitions})"
