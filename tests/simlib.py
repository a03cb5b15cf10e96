"""ctypes wrapper around the CPU SIMT simulator build of the CUDA kernels (tests/sim/).

TEST INFRASTRUCTURE ONLY: lets `-m "not gpu"` tests run the device code of
flac-py_b200/csrc/flacb200_kernels.cuh on the CPU and compare it with the golden vectors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SIM = os.path.join(_HERE, "sim")
_LIB = os.path.join(_SIM, "libflacb_sim.so")
_ROOT = os.path.dirname(_HERE)

MAX_LPC = 32
INFO_PARAMS = 64

FLAG_SPEC_CHANNELS = 0x01
FLAG_LENIENT = 0x08
FLAG_SLOT_COMPACT = 0x10
FLAG_PCM_INT32 = 0x20


class Stats(C.Structure):
    _fields_ = [
        ("frames", C.c_uint64), ("subframes", C.c_uint64), ("subframes_fixed", C.c_uint64 * 5),
        ("subframes_lpc", C.c_uint64), ("out_of_domain", C.c_uint64), ("bytes_in", C.c_uint64),
        ("bytes_out", C.c_uint64), ("first_bad_frame", C.c_int64), ("kernel_ms", C.c_float),
        ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("threads_per_frame", C.c_int32),
        ("smem_bytes", C.c_int32), ("stage_ms", C.c_float * 4),
    ]


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
        ("bits", C.c_int64),
        ("nparams", C.c_int32),
        ("params", C.c_uint8 * INFO_PARAMS),
    ]


def build(force: bool = False, small_window: bool = False) -> str:
    """small_window: a variant with a 256-word staging window, so that tests reach the packer's
    multi-pass path (with the product's 4096 words a tile can never outgrow the window)."""
    srcs = [os.path.join(_SIM, "sim_lib.cpp"), os.path.join(_SIM, "cpu_sim.h"),
            os.path.join(_ROOT, "flac-py_b200", "csrc", "flacb200_kernels.cuh"),
            os.path.join(_ROOT, "flac-py_b200", "csrc", "flacb200_host.h"),
            os.path.join(_ROOT, "include", "flacb200.h")]
    srcs.append(os.path.join(_ROOT, "flac-py_b200", "csrc", "flacb200_decode.cuh"))
    target = _LIB.replace(".so", "_smallwin.so") if small_window else _LIB
    stale = not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(s) for s in srcs)
    if force or stale:
        extra = ["-DFLACB_STG_WORDS=256"] if small_window else []
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", _SIM,
                        "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", *extra,
                        "-o", target, srcs[0]], check=True)
    return target


_lib = None
_libs = {}


def lib(small_window: bool = False):
    global _lib
    if small_window:
        if "small" not in _libs:
            _libs["small"] = _declare(C.CDLL(build(small_window=True)))
        return _libs["small"]
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.flacb_sim_encode.restype = C.c_int
        L.flacb_sim_encode.argtypes = [
            C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p,
            C.POINTER(SubframeInfo), C.POINTER(Stats), C.c_int, C.c_uint32]
        _lib = L
    return _lib


def _declare(L):
    L.flacb_sim_encode.restype = C.c_int
    L.flacb_sim_encode.argtypes = [
        C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
        C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p,
        C.POINTER(SubframeInfo), C.POINTER(Stats), C.c_int, C.c_uint32]
    return L


def pack_pcm(pcm: np.ndarray, bits: int) -> np.ndarray:
    a = np.ascontiguousarray(pcm, dtype=np.int32)
    nb = (bits + 7) // 8
    if nb == 2:
        return a.astype("<i2").view(np.uint8).reshape(-1)
    if nb == 3:
        return np.ascontiguousarray(a.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1)
    if nb == 1:
        return a.astype("i1").view(np.uint8).reshape(-1)
    return a.astype("<i4").view(np.uint8).reshape(-1)


def encode_frames(pcm: np.ndarray, bits: int, block_size: int, max_lpc_order: int = 12,
                  qlp_precision: int = 5, rice_min: int = 0, rice_max: int = 5,
                  first_frame_number: int = 0, flags: int = 0, want_info: bool = False,
                  threads: int = 0, plane_mode: int = 0, int32_input: bool = False,
                  small_window: bool = False):
    """-> (status, data, frame_offsets, info list | None, Stats)"""
    n, ch = pcm.shape
    if int32_input:
        raw = np.ascontiguousarray(pcm, dtype="<i4").view(np.uint8).reshape(-1)
        flags |= FLAG_PCM_INT32
    else:
        raw = pack_pcm(pcm, bits)
    nfr = (n + block_size - 1) // block_size
    per = (block_size * (bits + 4) + 7) // 8 + (1 << rice_max) + 256
    cap = (per * ch + 64) * max(nfr, 1)
    out = np.zeros(cap, dtype=np.uint8)
    offs = np.zeros(nfr + 1, dtype=np.uint64)
    out_len = C.c_uint64(0)
    info = (SubframeInfo * max(nfr * ch, 1))() if want_info else None
    st = Stats()
    rc = lib(small_window).flacb_sim_encode(raw.ctypes.data, n, ch, bits, block_size, max_lpc_order,
                                qlp_precision, rice_min, rice_max, first_frame_number, flags,
                                out.ctypes.data, cap, C.byref(out_len), offs.ctypes.data, info,
                                C.byref(st), threads, plane_mode)
    return rc, out[:out_len.value].tobytes(), offs, (list(info)[:nfr * ch] if info else None), st


def decode_frames(frames: bytes, streaminfo_bits: int, cap_samples: int | None = None):
    """-> (status, pcm int32 (n, frame_channels), frame_channels, n_candidates)"""
    L = lib()
    L.flacb_sim_decode.restype = C.c_int
    L.flacb_sim_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64,
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    raw = np.frombuffer(frames, dtype=np.uint8)
    cap = cap_samples or max(len(frames) * 16, 1 << 16)
    out = np.zeros(cap, dtype=np.int32)
    n, fch, nc = C.c_uint64(0), C.c_int(0), C.c_uint64(0)
    rc = L.flacb_sim_decode(raw.ctypes.data if raw.size else None, raw.size, streaminfo_bits,
                            out.ctypes.data, cap, C.byref(n), C.byref(fch), C.byref(nc))
    f = max(fch.value, 1)
    return rc, out[:n.value * f].reshape(-1, f).copy(), fch.value, nc.value


def chunk_plan(n_frames: int, wave: int, frame_pcm: int, want: int = 0, edge: int = 0, last: int = 0) -> list[int]:
    """first frame of every chunk of the host path, then n_frames (flacb200_host.h chunk_plan)"""
    L = lib()
    L.flacb_sim_chunk_plan.restype = C.c_int
    L.flacb_sim_chunk_plan.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_longlong, C.c_longlong, C.c_longlong,
                                       C.POINTER(C.c_uint64), C.c_int]
    cap = int(n_frames) + 2
    buf = (C.c_uint64 * cap)()
    n = L.flacb_sim_chunk_plan(n_frames, wave, frame_pcm, want, edge, last, buf, cap)
    assert n <= cap
    return [int(buf[i]) for i in range(n)]
