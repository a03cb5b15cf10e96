"""ctypes wrapper around the CPU oracle (oracle/flac_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product (flac-py_b200/)
never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libflac_oracle.so")

FO_OK, FO_E_ASSERT, FO_E_VALUE, FO_E_ZERODIV, FO_E_EOF, FO_E_NOSPACE, FO_E_OVERFLOW, FO_E_KEY, \
    FO_E_NOTIMPL = range(9)
FLAG_SPEC_CHANNELS = 1
MAX_LPC = 32
INFO_PARAMS = 512

_EXC = {
    FO_E_ASSERT: AssertionError, FO_E_VALUE: ValueError, FO_E_ZERODIV: ZeroDivisionError,
    FO_E_EOF: EOFError, FO_E_NOSPACE: MemoryError, FO_E_OVERFLOW: OverflowError,
    FO_E_KEY: KeyError, FO_E_NOTIMPL: NotImplementedError,
}


class SubframeInfo(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("order", C.c_int32), ("shift", C.c_int32), ("precision", C.c_int32),
        ("porder", C.c_int32), ("method", C.c_int32), ("fixed_order", C.c_int32),
        ("lpc_order", C.c_int32),
        ("fixed_err", C.c_int64 * 5),
        ("lpc_err", C.c_int64 * MAX_LPC),
        ("lpc_shift", C.c_int32 * MAX_LPC),
        ("lpc_ncoef", C.c_int32 * MAX_LPC),
        ("lpc_q", (C.c_int32 * MAX_LPC) * MAX_LPC),
        ("autoc", C.c_double * (MAX_LPC + 1)),
        ("qcoef", C.c_int32 * MAX_LPC),
        ("bits", C.c_int64),
        ("nparams", C.c_int32),
        ("params", C.c_uint8 * INFO_PARAMS),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flac_oracle.c")
    hdr = os.path.join(_HERE, "flac_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or (os.path.exists(src) and os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src),
                                                                          os.path.getmtime(hdr))))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libflac_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, u64p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
        L.fo_encode_frames.restype = C.c_int
        L.fo_encode_frames.argtypes = [
            i32p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_uint64, C.c_uint32, u8p, C.c_uint64, u64p, u64p, C.POINTER(SubframeInfo),
            C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.fo_encode_frames_mt.restype = C.c_int
        L.fo_encode_frames_mt.argtypes = [
            i32p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_uint64, C.c_uint32, u8p, C.c_uint64, u64p, u64p, C.c_int,
            C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.fo_stream_prologue.restype = C.c_int
        L.fo_stream_prologue.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_int, u8p]
        L.fo_decode_stream.restype = C.c_int
        L.fo_decode_stream.argtypes = [
            u8p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_int), u64p,
            i32p, C.c_uint64, u64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fo_crc8.restype = C.c_uint8
        L.fo_crc8.argtypes = [u8p, C.c_uint64]
        L.fo_crc16.restype = C.c_uint16
        L.fo_crc16.argtypes = [u8p, C.c_uint64]
        L.fo_coded_number.restype = C.c_int
        L.fo_coded_number.argtypes = [C.c_uint64, u8p]
        L.fo_neumaier_sum.restype = C.c_double
        L.fo_neumaier_sum.argtypes = [C.POINTER(C.c_double), C.c_uint64]
        L.fo_tukey.restype = None
        L.fo_tukey.argtypes = [C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int, frame: int, line: int):
        self.code, self.frame, self.line = code, frame, line
        self.exc_class = _EXC.get(code, RuntimeError)
        super().__init__(f"oracle status {code} ({self.exc_class.__name__}) at frame {frame}, "
                         f"flac_oracle.c:{line}")


@dataclass
class Encoded:
    data: bytes               # frames only (no prologue)
    frame_offsets: np.ndarray  # n_frames + 1
    info: list | None          # SubframeInfo per (frame, channel), flattened


def frame_bound(block_size: int, channels: int, bps: int, rice_max: int) -> int:
    per = (block_size * (bps + 4) + 7) // 8 + (1 << rice_max) + 256
    return per * channels + 64


def encode_frames(pcm: np.ndarray, bps: int, block_size: int, max_lpc_order: int = 12,
                  qlp_precision: int = 5, rice_min: int = 0, rice_max: int = 5,
                  first_frame_number: int = 0, flags: int = 0, want_info: bool = False,
                  threads: int = 1) -> Encoded:
    """pcm: int32 (n, channels).  Raises OracleError where the reference raises."""
    a = np.ascontiguousarray(pcm, dtype=np.int32)
    n, ch = a.shape
    nfr = (n + block_size - 1) // block_size
    cap = frame_bound(block_size, ch, bps, rice_max) * max(nfr, 1)
    out = np.empty(cap, dtype=np.uint8)
    offs = np.zeros(nfr + 1, dtype=np.uint64)
    out_len = C.c_uint64(0)
    ef, el = C.c_int64(-1), C.c_int(0)
    L = lib()
    u8p, u64p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    if threads > 1 and not want_info:
        rc = L.fo_encode_frames_mt(a.ctypes.data_as(i32p), n, ch, bps, block_size, max_lpc_order,
                                   qlp_precision, rice_min, rice_max, first_frame_number, flags,
                                   out.ctypes.data_as(u8p), cap, C.byref(out_len),
                                   offs.ctypes.data_as(u64p), threads, C.byref(ef), C.byref(el))
        info = None
    else:
        info = (SubframeInfo * max(nfr * ch, 1))() if want_info else None
        rc = L.fo_encode_frames(a.ctypes.data_as(i32p), n, ch, bps, block_size, max_lpc_order,
                                qlp_precision, rice_min, rice_max, first_frame_number, flags,
                                out.ctypes.data_as(u8p), cap, C.byref(out_len),
                                offs.ctypes.data_as(u64p), info, C.byref(ef), C.byref(el))
    if rc != FO_OK:
        raise OracleError(rc, ef.value, el.value)
    return Encoded(out[:out_len.value].tobytes(), offs, list(info)[:nfr * ch] if info else None)


def stream_prologue(sample_rate: int, channels: int, bps: int, total_samples: int,
                    block_size: int) -> bytes:
    buf = (C.c_uint8 * 42)()
    rc = lib().fo_stream_prologue(sample_rate, channels, bps, total_samples, block_size, buf)
    assert rc == FO_OK
    return bytes(buf)


def encode_stream(pcm: np.ndarray, sample_rate: int, bps: int, block_size: int, **kw) -> bytes:
    enc = encode_frames(pcm, bps, block_size, **kw)
    return stream_prologue(sample_rate, pcm.shape[1], bps, pcm.shape[0], block_size) + enc.data


def decode_stream(stream: bytes, cap_samples: int | None = None):
    """-> (sample_rate, bps, channels, total_samples, pcm int32 (n, frame_channels))."""
    L = lib()
    buf = np.frombuffer(stream, dtype=np.uint8)
    u8p, u64p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    if cap_samples is None:
        cap_samples = max(len(stream) * 16, 1 << 16)
    while True:
        out = np.empty(cap_samples, dtype=np.int32)
        rate, bps, ch, fch, line = C.c_uint32(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        total, n_out = C.c_uint64(), C.c_uint64()
        rc = L.fo_decode_stream(buf.ctypes.data_as(u8p), len(stream), C.byref(rate), C.byref(bps),
                                C.byref(ch), C.byref(total), out.ctypes.data_as(i32p), cap_samples,
                                C.byref(n_out), C.byref(fch), C.byref(line))
        if rc == FO_E_NOSPACE:
            cap_samples *= 4
            continue
        if rc != FO_OK:
            raise OracleError(rc, -1, line.value)
        fc = max(fch.value, 1)
        return (rate.value, bps.value, ch.value, total.value,
                out[:n_out.value * fc].reshape(-1, fc).copy())


def crc8(b: bytes) -> int:
    a = (C.c_uint8 * len(b)).from_buffer_copy(b) if b else (C.c_uint8 * 1)()
    return lib().fo_crc8(a, len(b))


def crc16(b: bytes) -> int:
    a = (C.c_uint8 * len(b)).from_buffer_copy(b) if b else (C.c_uint8 * 1)()
    return lib().fo_crc16(a, len(b))


def coded_number(x: int) -> bytes:
    buf = (C.c_uint8 * 7)()
    n = lib().fo_coded_number(x, buf)
    if n == -1:
        raise AssertionError
    if n < 0:
        raise ValueError(f"Cannot encode coded number: {x}")
    return bytes(buf[:n])


def neumaier_sum(x: np.ndarray) -> float:
    a = np.ascontiguousarray(x, dtype=np.float64)
    return lib().fo_neumaier_sum(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)


def tukey(n: int) -> np.ndarray:
    w = np.empty(n, dtype=np.float64)
    lib().fo_tukey(n, w.ctypes.data_as(C.POINTER(C.c_double)))
    return w
