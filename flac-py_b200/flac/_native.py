"""ctypes binding of libflacb200.so (include/flacb200.h) -- the only way the package computes.

The library is built in-tree by `__graft_entry__.build()` (or `make -C flac-py_b200/csrc`) and
loaded from flac-py_b200/csrc/.  If it is missing, or no CUDA device is usable, every encode /
decode call raises: there is deliberately no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLACB200_LIB", os.path.join(os.path.dirname(_HERE), "csrc", "libflacb200.so"))

OK, E_ASSERT, E_VALUE, E_ZERODIV, E_EOF, E_NOSPACE, E_OVERFLOW, E_KEY, E_NOTIMPL, E_CUDA, \
    E_UNSUPPORTED, E_RANGE, E_INTERNAL = range(13)

SPEC_CHANNELS = 0x01
INPUT_ON_DEVICE = 0x02
OUTPUT_ON_DEVICE = 0x04
LENIENT = 0x08
SLOT_COMPACT = 0x10
PCM_INT32 = 0x20
STAGE_TIMES = 0x40  # time every stage on its own (stats.stage_ms) instead of overlapping them

MAX_LPC = 32
INFO_PARAMS = 64


class NativeError(RuntimeError):
    """The CUDA library is missing or a CUDA call failed."""


# status -> the exception class the reference raises in the same situation
_EXC = {
    E_ASSERT: AssertionError, E_VALUE: ValueError, E_ZERODIV: ZeroDivisionError, E_EOF: EOFError,
    E_NOSPACE: MemoryError, E_OVERFLOW: OverflowError, E_KEY: KeyError,
    E_NOTIMPL: NotImplementedError, E_CUDA: NativeError, E_UNSUPPORTED: NotImplementedError,
    E_RANGE: OverflowError, E_INTERNAL: NativeError,
}


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


EXPORTS = [
    "flacb_version", "flacb_device_count", "flacb_strerror", "flacb_last_cuda_error",
    "flacb_create", "flacb_destroy", "flacb_host_alloc", "flacb_host_free", "flacb_device_alloc",
    "flacb_device_free", "flacb_memcpy_h2d", "flacb_memcpy_d2h", "flacb_encode_bound",
    "flacb_encode_pcm", "flacb_stream_prologue", "flacb_decode_frames", "flacb_decode_output", "flacb_scan_frames", "flacb_timer_begin",
    "flacb_timer_end", "flacb_sync", "flacb_kernel_launches",
]

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C flac-py_b200/csrc` (nvcc, sm_100a).  There is no CPU "
                "fallback.")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        L.flacb_version.restype = i32
        L.flacb_device_count.restype = i32
        L.flacb_strerror.restype = C.c_char_p
        L.flacb_strerror.argtypes = [i32]
        L.flacb_last_cuda_error.restype = C.c_char_p
        L.flacb_create.restype = i32
        L.flacb_create.argtypes = [i32, C.POINTER(vp)]
        L.flacb_destroy.restype = None
        L.flacb_destroy.argtypes = [vp]
        L.flacb_host_alloc.restype = i32
        L.flacb_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.flacb_host_free.restype = None
        L.flacb_host_free.argtypes = [vp]
        L.flacb_device_alloc.restype = i32
        L.flacb_device_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
        L.flacb_device_free.restype = None
        L.flacb_device_free.argtypes = [vp, vp]
        L.flacb_memcpy_h2d.restype = i32
        L.flacb_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
        L.flacb_memcpy_d2h.restype = i32
        L.flacb_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
        L.flacb_encode_bound.restype = u64
        L.flacb_encode_bound.argtypes = [u64, i32, i32, i32, i32]
        L.flacb_encode_pcm.restype = i32
        L.flacb_encode_pcm.argtypes = [vp, vp, u64, i32, i32, i32, i32, i32, i32, i32, u64, u32,
                                       vp, u64, C.POINTER(u64), vp, vp, C.POINTER(Stats)]
        L.flacb_stream_prologue.restype = i32
        L.flacb_stream_prologue.argtypes = [u32, i32, i32, u64, i32, vp]
        L.flacb_decode_frames.restype = i32
        L.flacb_decode_frames.argtypes = [vp, vp, u64, i32, u32, vp, u64, C.POINTER(u64),
                                          C.POINTER(i32), C.POINTER(Stats)]
        L.flacb_timer_begin.restype = i32
        L.flacb_scan_frames.restype = i32
        L.flacb_scan_frames.argtypes = [vp, vp, u64, i32, u32, vp, vp, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]
        L.flacb_decode_output.restype = i32
        L.flacb_decode_output.argtypes = [vp, u32, vp, u64, C.POINTER(u64), C.POINTER(i32), C.POINTER(Stats)]
        L.flacb_timer_begin.argtypes = [vp]
        L.flacb_timer_end.restype = i32
        L.flacb_timer_end.argtypes = [vp, C.POINTER(C.c_float)]
        L.flacb_sync.restype = i32
        L.flacb_sync.argtypes = [vp]
        L.flacb_kernel_launches.restype = u64
        _lib = L
        return L


def raise_for_status(rc: int, what: str = ""):
    if rc == OK:
        return
    L = lib()
    msg = L.flacb_strerror(rc).decode()
    if rc == E_CUDA:
        msg += ": " + L.flacb_last_cuda_error().decode()
    if what:
        msg = f"{what}: {msg}"
    raise _EXC.get(rc, RuntimeError)(msg)


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self._p = C.c_void_p()
        raise_for_status(lib().flacb_host_alloc(C.byref(self._p), max(nbytes, 1)), "host_alloc")
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(self._p.value))[:nbytes]

    @property
    def ptr(self) -> int:
        return self._p.value

    def free(self):
        if self._p:
            self.array = None
            lib().flacb_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001
            pass


class Context:
    """One encoder / decoder context on one GPU (not thread-safe; one per thread)."""

    def __init__(self, device: int | None = None):
        if device is None:
            device = int(os.environ.get("FLACB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        L = lib()
        if L.flacb_device_count() < 1:
            raise NativeError("no CUDA device visible: the flac package has no CPU fallback")
        self._h = C.c_void_p()
        raise_for_status(L.flacb_create(device, C.byref(self._h)), f"flacb_create({device})")
        self.device = device

    def close(self):
        for b in getattr(self, "_staging", {}).values():
            b.free()
        self._staging = {}
        if self._h:
            lib().flacb_destroy(self._h)
            self._h = C.c_void_p()

    def staging(self, which: str, nbytes: int) -> "PinnedBuffer":
        """Page-locked staging memory of the context ("in" / "out"), grown on demand and reused between
        calls: Python objects (bytes, numpy arrays) live in pageable memory, from which the copy engines
        run at a fraction of their rate and never at the same time as the kernels."""
        st = self.__dict__.setdefault("_staging", {})
        b = st.get(which)
        if b is None or b.nbytes < nbytes:
            if b is not None:
                b.free()
            b = st[which] = PinnedBuffer(nbytes + nbytes // 4 + 4096)
        return b

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- raw device memory (benchmarks keep PCM resident) ---------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        raise_for_status(lib().flacb_device_alloc(self._h, C.byref(p), nbytes), "device_alloc")
        return p.value

    def device_free(self, ptr: int):
        lib().flacb_device_free(self._h, C.c_void_p(ptr))

    def h2d(self, dst: int, src: int, nbytes: int):
        raise_for_status(lib().flacb_memcpy_h2d(self._h, C.c_void_p(dst), C.c_void_p(src), nbytes), "h2d")

    def d2h(self, dst: int, src: int, nbytes: int):
        raise_for_status(lib().flacb_memcpy_d2h(self._h, C.c_void_p(dst), C.c_void_p(src), nbytes), "d2h")

    def timer_begin(self):
        raise_for_status(lib().flacb_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        raise_for_status(lib().flacb_timer_end(self._h, C.byref(ms)))
        return ms.value

    def sync(self):
        raise_for_status(lib().flacb_sync(self._h))

    # -- the hot path -----------------------------------------------------------------------
    def encode_raw(self, pcm_ptr: int, n_sample_frames: int, channels: int, bits: int,
                   block_size: int, max_lpc_order: int, qlp_precision: int, rice_min: int,
                   rice_max: int, first_frame_number: int, flags: int, out_ptr: int, out_cap: int,
                   frame_offsets: np.ndarray | None = None, info=None):
        """Thin call-through; returns (status, out_len, Stats)."""
        st = Stats()
        out_len = C.c_uint64(0)
        rc = lib().flacb_encode_pcm(
            self._h, C.c_void_p(pcm_ptr), n_sample_frames, channels, bits, block_size,
            max_lpc_order, qlp_precision, rice_min, rice_max, first_frame_number, flags,
            C.c_void_p(out_ptr), out_cap, C.byref(out_len),
            frame_offsets.ctypes.data_as(C.c_void_p) if frame_offsets is not None else None,
            C.cast(info, C.c_void_p) if info is not None else None, C.byref(st))
        return rc, out_len.value, st

    def decode_raw(self, frames_ptr: int, nbytes: int, streaminfo_bits: int, flags: int,
                   out_ptr: int, cap_samples: int):
        st = Stats()
        n = C.c_uint64(0)
        fch = C.c_int(0)
        rc = lib().flacb_decode_frames(self._h, C.c_void_p(frames_ptr), nbytes, streaminfo_bits,
                                       flags, C.c_void_p(out_ptr), cap_samples, C.byref(n),
                                       C.byref(fch), C.byref(st))
        return rc, n.value, fch.value, st

    def scan_frames(self, frames_ptr: int, nbytes: int, streaminfo_bits: int, flags: int, cap_frames: int):
        """-> (status, offsets uint64 [n + 1], block_sizes uint32 [n], n_sample_frames, frame_channels)"""
        offs = np.zeros(cap_frames + 1, dtype=np.uint64)
        sizes = np.zeros(max(cap_frames, 1), dtype=np.uint32)
        n, ns, fch = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        rc = lib().flacb_scan_frames(self._h, C.c_void_p(frames_ptr), nbytes, streaminfo_bits, flags, offs.ctypes.data,
                                     sizes.ctypes.data, cap_frames, C.byref(n), C.byref(ns), C.byref(fch))
        k = min(n.value, cap_frames)
        return rc, offs[:k + 1], sizes[:k], n.value, ns.value, fch.value

    def decode_output(self, flags: int, out_ptr: int, cap_samples: int):
        """After decode_raw returned E_NOSPACE: write the frames it decoded (they are still in the
        context) to a buffer of the size it reported."""
        st = Stats()
        n = C.c_uint64(0)
        fch = C.c_int(0)
        rc = lib().flacb_decode_output(self._h, flags, C.c_void_p(out_ptr), cap_samples, C.byref(n), C.byref(fch), C.byref(st))
        return rc, n.value, fch.value, st


_default: dict[int, Context] = {}


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = int(os.environ.get("FLACB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _default.get(device)
    if ctx is None:
        ctx = _default[device] = Context(device)
    return ctx


def encode_bound(n_sample_frames: int, channels: int, bits: int, block_size: int, rice_max: int) -> int:
    return lib().flacb_encode_bound(n_sample_frames, channels, bits, block_size, rice_max)


def stream_prologue(sample_rate: int, channels: int, bits: int, total_samples: int, block_size: int) -> bytes:
    buf = (C.c_uint8 * 42)()
    raise_for_status(lib().flacb_stream_prologue(sample_rate, channels, bits, total_samples, block_size, buf))
    return bytes(buf)
