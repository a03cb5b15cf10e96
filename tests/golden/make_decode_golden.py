#!/usr/bin/env python3
"""Mint decoder fixtures from the REAL reference decoder (turlando/flac-py, flac/decoder.py) for
stream features its own encoder never produces: constant and verbatim subframes, escaped Rice
partitions, the wasted-bits flag, left/side, side/right and mid/side channel assignments, the
5-bit Rice method, a reserved subframe type.

Run in the build container, where the reference is mounted at /root/reference:

    python tests/golden/make_decode_golden.py

The streams are crafted bit by bit here (with the reference's own bit writer and CRC helpers);
every one is decoded by the unmodified reference decoder, and the stream (hex) and the samples it
yields, or the exception class it raises and the samples yielded before that, are stored in
tests/golden/decode_cases.json.  Nothing in the repo imports the reference at test time.
"""
from __future__ import annotations

import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from flac import common as _common  # noqa: E402  (the reference's)
from flac.crc import crc8 as _crc8, crc16 as _crc16  # noqa: E402
from flac.decoder import decode  # noqa: E402


def crc8(bs: bytes) -> int:
    return _crc8(bs, _common.CRC8_POLYNOMIAL)


def crc16(bs: bytes) -> int:
    return _crc16(bs, _common.CRC16_POLYNOMIAL)

BS = 16           # block size of every crafted frame
BITS = 16


class W:
    """MSB-first bit writer"""

    def __init__(self):
        self.bits = []

    def u(self, v: int, n: int):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def s(self, v: int, n: int):
        self.u(v & ((1 << n) - 1), n)

    def unary(self, q: int):
        self.bits += [0] * q + [1]

    def pad(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def bytes(self) -> bytes:
        assert len(self.bits) % 8 == 0
        return bytes(int("".join(map(str, self.bits[i:i + 8])), 2) for i in range(0, len(self.bits), 8))


def prologue(total: int, channels: int = 2) -> bytes:
    w = W()
    for ch in b"fLaC":
        w.u(ch, 8)
    w.u(1, 1), w.u(0, 7), w.u(34, 24)
    w.u(BS, 16), w.u(BS, 16), w.u(0, 24), w.u(0, 24)
    w.u(44100, 20), w.u(channels - 1, 3), w.u(BITS - 1, 5), w.u(total, 36)
    for _ in range(16):
        w.u(0, 8)
    return w.bytes()


def rice(w: W, r: int, k: int):
    u = (r << 1) if r >= 0 else ((-r) << 1) - 1
    w.unary(u >> k)
    if k:
        w.u(u & ((1 << k) - 1), k)


def frame(number: int, chcode: int, subframes) -> bytes:
    """subframes: callables writing one subframe into a W"""
    w = W()
    w.u(0x7FFC, 15), w.u(0, 1)         # sync, blocking strategy
    w.u(0b0110, 4), w.u(0, 4)          # block size: 8-bit (bs - 1) follows; sample rate from STREAMINFO
    w.u(chcode, 4), w.u(0, 3), w.u(0, 1)
    w.u(number, 8)                     # coded number < 128: one byte
    w.u(BS - 1, 8)
    hdr = w.bytes()
    w.u(crc8(hdr), 8)
    for sf in subframes:
        sf(w)
    w.pad()
    body = w.bytes()
    return body + crc16(body).to_bytes(2, "big")


def sf_constant(value, bits=BITS):
    def f(w):
        w.u(0, 1), w.u(0b000000, 6), w.u(0, 1)
        w.s(value, bits)
    return f


def sf_verbatim(samples, bits=BITS):
    def f(w):
        w.u(0, 1), w.u(0b000001, 6), w.u(0, 1)
        for v in samples:
            w.s(v, bits)
    return f


def sf_fixed(order, warm, residuals, k=3, escape_bits=None, method5=False, wasted=None, bits=BITS, porder=0):
    def f(w):
        w.u(0, 1), w.u(0b001000 | order, 6)
        if wasted is None:
            w.u(0, 1)
        else:
            w.u(1, 1), w.unary(wasted)
        for v in warm:
            w.s(v, bits)
        w.u(1 if method5 else 0, 2), w.u(porder, 4)
        nparts = 1 << porder
        per = BS // nparts
        pos = 0
        for pi in range(nparts):
            cnt = per - order if pi == 0 else per
            part = residuals[pos:pos + cnt]
            pos += cnt
            if escape_bits is not None and pi == nparts - 1:
                w.u(31 if method5 else 15, 5 if method5 else 4)
                w.u(escape_bits, 5)
                for r in part:
                    w.s(r, escape_bits)
            else:
                w.u(k, 5 if method5 else 4)
                for r in part:
                    rice(w, r, k)
    return f


def sf_lpc(order, warm, precision, shift, coefs, residuals, k=2, bits=BITS):
    def f(w):
        w.u(0, 1), w.u(0b100000 | (order - 1), 6), w.u(0, 1)
        for v in warm:
            w.s(v, bits)
        w.u(precision - 1, 4), w.s(shift, 5)
        for c in coefs:
            w.s(c, precision)
        w.u(0, 2), w.u(0, 4)
        w.u(k, 4)
        for r in residuals:
            rice(w, r, k)
    return f


def sf_reserved():
    def f(w):
        w.u(0, 1), w.u(0b000100, 6), w.u(0, 1)
    return f


def res(n, seed, amp=9):
    out, x = [], seed
    for _ in range(n):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        out.append((x >> 8) % (2 * amp + 1) - amp)
    return out


def build():
    ramp = [100 * i - 700 for i in range(BS)]
    cases = {}
    cases["constant_verbatim"] = [frame(0, 1, [sf_constant(-1234), sf_verbatim(ramp)]),
                                  frame(1, 1, [sf_verbatim(ramp[::-1]), sf_constant(32767)])]
    cases["fixed_orders_partitions"] = [frame(0, 1, [sf_fixed(0, [], res(16, 1), k=4, porder=2),
                                                     sf_fixed(1, [55], res(15, 2), k=0, porder=1)]),
                                        frame(1, 1, [sf_fixed(3, [5, -6, 7], res(13, 3), porder=0),
                                                     sf_fixed(4, [1000, 990, 985, 970], res(12, 4), k=5, porder=2)])]
    cases["escaped_partitions"] = [frame(0, 1, [sf_fixed(2, [10, 12], res(14, 5, 100), escape_bits=9),
                                                sf_fixed(1, [-3], res(15, 6, 3000), escape_bits=13, porder=1)])]
    cases["method5_and_big_parameters"] = [frame(0, 1, [sf_fixed(2, [10, 12], res(14, 7, 40000), k=17, method5=True),
                                                        sf_fixed(0, [], res(16, 8, 5), k=1, method5=True, escape_bits=7, porder=2)])]
    cases["wasted_bits_flag"] = [frame(0, 1, [sf_fixed(1, [64], [2 * v for v in res(15, 9)], wasted=0, bits=BITS),
                                              sf_fixed(1, [32], res(15, 10), wasted=2, bits=BITS - 1)])]
    cases["left_side"] = [frame(0, 8, [sf_fixed(1, [500], res(15, 11)), sf_fixed(1, [-20], res(15, 12), bits=17)])]
    cases["side_right"] = [frame(0, 9, [sf_fixed(1, [33], res(15, 13), bits=17), sf_fixed(2, [7, 9], res(14, 14))])]
    cases["mid_side"] = [frame(0, 10, [sf_fixed(1, [401], res(15, 15)), sf_fixed(1, [-3], res(15, 16), bits=17)]),
                         frame(1, 10, [sf_verbatim(ramp), sf_verbatim([v % 7 - 3 for v in ramp], bits=17)])]
    cases["lpc_with_negative_coefficients"] = [frame(0, 1, [sf_lpc(3, [100, 120, 130], 7, 4, [31, -20, 9], res(13, 17)),
                                                            sf_lpc(8, [5, 9, -4, 3, 0, 1, 8, -9], 12, 9, [900, -700, 400, -200, 90, -40, 20, -3], res(8, 18))])]
    cases["lpc_negative_shift_raises"] = [frame(0, 1, [sf_lpc(1, [100], 5, -2, [7], res(15, 19)), sf_constant(1)])]
    cases["reserved_subframe_type_raises"] = [frame(0, 1, [sf_constant(5), sf_constant(6)]),
                                              frame(1, 1, [sf_reserved(), sf_constant(1)])]
    cases["three_channels"] = [frame(0, 2, [sf_constant(1), sf_fixed(1, [9], res(15, 20)), sf_verbatim(ramp)])]
    return cases


def main():
    out = {}
    for name, frames in build().items():
        ch = 3 if name == "three_channels" else 2
        stream = prologue(BS * len(frames), ch) + b"".join(frames)
        got, exc = [], None
        try:
            rate, bits, channels, total, it = decode(io.BufferedReader(io.BytesIO(stream)))
            for s in it:
                got.append(list(s))
        except Exception as e:  # noqa: BLE001
            exc = type(e).__name__
        out[name] = {"stream": stream.hex(), "samples": got, "raises": exc}
        print(f"{name}: {len(stream)} bytes -> {len(got)} sample frames, raises {exc}")
    import platform
    out["_meta"] = {"python": platform.python_version(), "how": "tests/golden/make_decode_golden.py, unmodified reference decoder"}
    with open(os.path.join(HERE, "decode_cases.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
