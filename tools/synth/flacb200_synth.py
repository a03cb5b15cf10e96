"""Deterministic, integer-only synthetic PCM for parity tests and benchmarks.

Every sample is a pure function of (seed, class, sample index, channel), so any
slice of a stream can be generated independently and is identical on every host
(numpy uint64/int64 wrap-around arithmetic only; the one float step is the
sine look-up table, which is built once with a fixed recipe and rounded to
integers before use).

Signal classes follow SURVEY.md section 8(d) / BASELINE.md section 3 and are
kept inside the reference encoder's domain: no digital
silence, no DC, every Rice partition has mean zig-zag >= 1.

    sine      3-5 partials per channel, about -6 dBFS in total, +-3 LSB dither
    lowpass   white noise through a short integer FIR (rho = +0.9 decay)
    highpass  same with rho = -0.9
    white     near-white gaussian-ish noise, sigma about full-scale / 8
              (the only class on which the reference picks SubframeLPC)
    silence   dither only, uniform in [-3, 3] at the 16-bit level
    clip      sine at +6 dB over full scale, hard clipped, then re-dithered

`mixed` cycles through the six classes in segments of SEGMENT sample frames.
"""
from __future__ import annotations

import numpy as np

CLASSES = ("sine", "lowpass", "highpass", "white", "silence", "clip")
SEGMENT = 1 << 16  # sample frames per class segment in "mixed" streams

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_LUT_BITS = 12
_LUT = None


def _sine_lut() -> np.ndarray:
    """Q30 sine table, 2**_LUT_BITS entries (one full period)."""
    global _LUT
    if _LUT is None:
        n = 1 << _LUT_BITS
        k = np.arange(n, dtype=np.float64)
        _LUT = np.round(np.sin(2.0 * np.pi * (k + 0.5) / n) * (1 << 30)).astype(np.int64)
    return _LUT


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def _rand_u64(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    base = np.uint64((seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        return _mix64(idx.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + base)


def _uniform(seed: int, stream: int, idx: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Uniform integers in [lo, hi] (inclusive)."""
    span = hi - lo + 1
    r = _rand_u64(seed, stream, idx) >> np.uint64(33)  # 31 bits
    return (r.astype(np.int64) * span >> 31) + lo


def _gauss(seed: int, stream: int, idx: np.ndarray, sigma: int) -> np.ndarray:
    """Sum of four uniform 16-bit values, centred, scaled to about `sigma`."""
    r = _rand_u64(seed, stream, idx)
    s = np.zeros(idx.shape, dtype=np.int64)
    for k in range(4):
        s += ((r >> np.uint64(16 * k)) & np.uint64(0xFFFF)).astype(np.int64)
    s -= 2 * 65535
    # std of the sum is 65536 / sqrt(3) = 37837; scale by sigma / 37837
    return (s * sigma) // 37837


def _sine(idx: np.ndarray, step: int, phase0: int, amp: int) -> np.ndarray:
    lut = _sine_lut()
    with np.errstate(over="ignore"):
        ph = (idx.astype(np.uint64) * np.uint64(step) + np.uint64(phase0)) & np.uint64(0xFFFFFFFF)
    k = (ph >> np.uint64(32 - _LUT_BITS)).astype(np.int64)
    return (lut[k] * amp) >> 30


def _class_samples(cls: str, idx: np.ndarray, ch: int, bits: int, rate: int, seed: int) -> np.ndarray:
    fs = (1 << (bits - 1)) - 1
    st = 16 * ch  # stream id base: decorrelates channels
    if cls == "sine":
        nparts = 3 + (seed + ch) % 3
        out = np.zeros(idx.shape, dtype=np.int64)
        for p in range(nparts):
            f = 110 * (p + 1) * (ch + 2) + 37 * p + (seed % 13)
            step = (f << 32) // rate
            out += _sine(idx, step, (p * 0x1234567 + ch * 0x89ABCDE) & 0xFFFFFFFF, fs // (2 * nparts))
        return out + _uniform(seed, st + 1, idx, -3, 3)
    if cls in ("lowpass", "highpass"):
        sigma = 1000 if bits <= 16 else (1 << 18)
        rho_q15 = 29491 if cls == "lowpass" else -29491  # +-0.9 in Q15
        out = np.zeros(idx.shape, dtype=np.int64)
        c = 1 << 15
        # 16-tap FIR of an AR(1) impulse response; sliceable (needs idx - k >= 0)
        for k in range(16):
            e = _gauss(seed, st + 2, np.maximum(idx - k, 0), sigma)
            e = np.where(idx - k >= 0, e, 0)
            out += (e * c) >> 15
            c = (c * rho_q15) >> 15
        # scale back towards sigma (gain of the filter is about 2.3)
        return (out * 7) >> 4
    if cls == "white":
        return _gauss(seed, st + 3, idx, fs // 8)
    if cls == "silence":
        # Dither only, +-3 LSB at the 16-bit level.  For wider samples it is two-tap shaped dither at that
        # level: on WHITE low-level noise the reference's best LPC order and its best fixed predictor give
        # sums of |residual| that differ by a short random walk, and they tie -- the reference's
        # `assert False`, encoder.py:157 -- about once in 3 000 subframes (measured: 151 frames of the
        # 375 000 of config 4, 3 of config 3's 21 094).  With a little correlation the (sign-inverted) LPC
        # predictor loses by a margin and the long 24-bit corpora stay inside the reference's domain.
        if bits <= 16:
            return _uniform(seed, st + 4, idx, -3, 3)
        d = 3 << (bits - 16)
        prev = np.where(idx >= 1, _uniform(seed, st + 4, np.maximum(idx - 1, 0), -d, d), 0)
        return _uniform(seed, st + 4, idx, -d, d) + prev
    if cls == "clip":
        f = 1000 + 100 * ch if bits <= 16 else 1500 + 100 * ch
        step = (f << 32) // rate
        x = _sine(idx, step, ch * 0x3243F6A, 2 * fs) + _gauss(seed, st + 5, idx, 3)
        x = np.clip(x, -fs - 1, fs)
        d = 3 if bits <= 16 else 12
        x = x + _uniform(seed, st + 6, idx, -d, d)
        return x
    raise ValueError(f"unknown signal class {cls!r}")


def pcm(n: int, channels: int, bits: int = 16, rate: int = 44100, cls: str = "mixed",
        seed: int = 0xF1AC0000, start: int = 0) -> np.ndarray:
    """Return int32 array of shape (n, channels): sample frames [start, start+n)."""
    idx = np.arange(start, start + n, dtype=np.int64)
    fs = (1 << (bits - 1)) - 1
    out = np.empty((n, channels), dtype=np.int32)
    for c in range(channels):
        if cls == "mixed":
            x = np.zeros(n, dtype=np.int64)
            seg = idx // SEGMENT
            for k, name in enumerate(CLASSES):
                m = (seg % len(CLASSES)) == k
                if m.any():
                    x[m] = _class_samples(name, idx[m], c, bits, rate, seed)
        else:
            x = _class_samples(cls, idx, c, bits, rate, seed)
        out[:, c] = np.clip(x, -fs - 1, fs).astype(np.int32)
    return out


def pack_le(samples: np.ndarray, bits: int) -> bytes:
    """Interleaved int32 (n, channels) -> little-endian packed PCM bytes (WAV payload): containers
    of ceil(bits / 8) bytes, sign extended (12-bit samples travel in 2 bytes, 20-bit in 3)."""
    a = np.ascontiguousarray(samples, dtype=np.int32)
    if bits not in (8, 16, 24, 32) and 4 <= bits <= 32:
        return pack_le(a, 8 * ((bits + 7) // 8))
    if bits == 16:
        return a.astype("<i2").tobytes()
    if bits == 24:
        b = a.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        return np.ascontiguousarray(b).tobytes()
    if bits == 32:
        return a.astype("<i4").tobytes()
    if bits == 8:  # WAV 8-bit is unsigned; we keep signed two's complement here
        return a.astype("i1").tobytes()
    raise ValueError(f"unsupported sample size {bits}")


def unpack_le(data: bytes, channels: int, bits: int) -> np.ndarray:
    """Inverse of pack_le -> int32 (n, channels)."""
    if bits == 16:
        a = np.frombuffer(data, dtype="<i2").astype(np.int32)
    elif bits == 24:
        b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        a = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        a = (a ^ 0x800000) - 0x800000
    elif bits == 32:
        a = np.frombuffer(data, dtype="<i4").astype(np.int32)
    elif bits == 8:
        a = np.frombuffer(data, dtype="i1").astype(np.int32)
    else:
        raise ValueError(f"unsupported sample size {bits}")
    return a.reshape(-1, channels)


# ---- the same generator on the device (tools/synth/flacb200_synth.cu): benchmark corpora that no host
# could hold or produce in time are written straight into HBM, bit for bit what pcm() returns ----
_dev_lib = None
_dev_lut_set = set()


def device_lib():
    import ctypes as C
    import os
    import subprocess
    global _dev_lib
    if _dev_lib is None:
        here = os.path.dirname(os.path.abspath(__file__))
        so = os.path.join(here, "libflacb200_synth.so")
        src = os.path.join(here, "flacb200_synth.cu")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", here, "-B", "libflacb200_synth.so"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.flacb_synth_set_lut.restype = C.c_int
        L.flacb_synth_set_lut.argtypes = [C.c_int, C.c_void_p]
        L.flacb_synth_pcm.restype = C.c_int
        L.flacb_synth_pcm.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]
        _dev_lib = L
    return _dev_lib


def pcm_device(device: int, dst_ptr: int, n: int, channels: int, bits: int = 16, rate: int = 44100, cls: str = "mixed",
               seed: int = 0xF1AC0000, start: int = 0) -> int:
    """Write pack_le(pcm(n, channels, bits, rate, cls, seed, start), bits) to device memory at dst_ptr
    (a pointer on CUDA device `device`); returns the number of bytes written."""
    L = device_lib()
    if device not in _dev_lut_set:
        lut = np.ascontiguousarray(_sine_lut(), dtype=np.int64)
        rc = L.flacb_synth_set_lut(device, lut.ctypes.data)
        if rc:
            raise RuntimeError(f"flacb_synth_set_lut failed ({rc})")
        _dev_lut_set.add(device)
    k = 6 if cls == "mixed" else CLASSES.index(cls)
    rc = L.flacb_synth_pcm(device, dst_ptr, start, n, channels, bits, rate, k, seed & 0xFFFFFFFFFFFFFFFF)
    if rc:
        raise RuntimeError(f"flacb_synth_pcm failed ({rc})")
    return n * channels * ((bits + 7) // 8)
