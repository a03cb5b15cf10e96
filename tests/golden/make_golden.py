#!/usr/bin/env python3
"""Mint the golden fixtures from the REAL reference (turlando/flac-py).

Run in the build container, where the reference is mounted at /root/reference:

    python tests/golden/make_golden.py

For each case the unmodified reference encoder (flac/encoder.py:48 encode) is run on
deterministic synthetic PCM (tools/synth/flacb200_synth.py) and the following is
stored under tests/golden/:

    cases.json          parameters, sha256 of PCM and of the reference stream, frame chunk
                        lengths, per-subframe decisions (type/order/shift/coefficients/
                        autocorrelation bits/Rice partition order + parameters), or the
                        exception class the reference raises on out-of-domain input
    <case>.flac         the reference's output stream (small cases only)

The reference decoder (flac/decoder.py:31 decode) is run on every stereo stream and
must return the input PCM; its result hash is stored too.  Nothing else in the repo
imports the reference at test time: the oracle and the CUDA path are checked against
these files.
"""
from __future__ import annotations

import hashlib
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "flac-py_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools", "synth"))

import numpy as np  # noqa: E402

import flacb200_synth as S  # noqa: E402
import refharness as R  # noqa: E402

# name, cls, n, channels, bits, rate, block, L, q, rmin, rmax, seed
CASES = [
    # config 1: 16-bit mono, fixed only (LPC stub), incl. the 3240 tail
    ("c1_mixed", "mixed", 4608 * 2 + 3240, 1, 16, 44100, 4608, 0, 5, 0, 5, 0xF1AC0001),
    ("c1_sine", "sine", 4608 + 3240, 1, 16, 44100, 4608, 0, 5, 0, 5, 0xF1AC0001),
    ("c1_white", "white", 4608 + 3240, 1, 16, 44100, 4608, 0, 5, 0, 5, 0xF1AC0001),
    # config 2: 16-bit stereo, l=12 q=5 r=0..5, incl. the 864 tail
    ("c2_sine", "sine", 4608 * 2 + 864, 2, 16, 44100, 4608, 12, 5, 0, 5, 0xF1AC0002),
    ("c2_lowpass", "lowpass", 4608 * 2 + 864, 2, 16, 44100, 4608, 12, 5, 0, 5, 0xF1AC0002),
    ("c2_highpass", "highpass", 4608 * 2 + 864, 2, 16, 44100, 4608, 12, 5, 0, 5, 0xF1AC0002),
    ("c2_white", "white", 4608 * 3 + 864, 2, 16, 44100, 4608, 12, 5, 0, 5, 0xF1AC0002),
    ("c2_silence", "silence", 4608 * 2 + 864, 2, 16, 44100, 4608, 12, 5, 0, 5, 0xF1AC0002),
    ("c2_clip", "clip", 4608 * 2 + 864, 2, 16, 44100, 4608, 12, 5, 0, 5, 0xF1AC0002),
    # config 3: 24-bit stereo 96 kHz, b=16384 l=32 r=0..8, incl. the 12288 tail
    ("c3_white", "white", 16384 + 12288, 2, 24, 96000, 16384, 32, 5, 0, 8, 0xF1AC0003),
    ("c3_sine", "sine", 16384 + 12288, 2, 24, 96000, 16384, 32, 5, 0, 8, 0xF1AC0003),
    ("c3_clip", "clip", 16384 + 12288, 2, 24, 96000, 16384, 32, 5, 0, 8, 0xF1AC0003),
    # config 4: 24-bit 8 channels 48 kHz, defaults
    ("c4_white", "white", 4608 * 2, 8, 24, 48000, 4608, 12, 5, 0, 5, 0xF1AC0004),
    ("c4_mixed", "mixed", 4608 * 2, 8, 24, 48000, 4608, 12, 5, 0, 5, 0xF1AC0004),
    # other precisions / ranges / block sizes
    ("q8_white", "white", 4096 + 1000, 2, 16, 44100, 4096, 12, 8, 2, 5, 11),
    ("q12_white24", "white", 4096 + 1000, 2, 24, 96000, 4096, 32, 12, 0, 6, 12),
    ("q15_white24", "white", 4096 + 1000, 2, 24, 96000, 4096, 32, 15, 0, 6, 13),
    ("bs1000", "white", 1000 + 255 + 20, 2, 16, 44100, 1000, 8, 7, 0, 3, 14),
    ("bs255", "lowpass", 255 + 100, 2, 16, 44100, 255, 8, 7, 0, 3, 15),
    ("bs192_r0", "sine", 192 * 5, 2, 16, 44100, 192, 4, 6, 0, 0, 16),
    ("many_frames", "white", 256 * 140, 2, 16, 44100, 256, 6, 5, 0, 4, 17),  # 2-byte coded numbers
]

# out-of-domain inputs: the reference raises
OOD = [
    ("ood_silence", lambda: np.zeros((4608, 2), np.int32), 2, 16, 44100, 4608, 12, 5, 0, 5),
    ("ood_dc", lambda: np.full((4608, 2), 1000, np.int32), 2, 16, 44100, 4608, 12, 5, 0, 5),
    ("ood_lsb", lambda: (S.pcm(4608, 2, 16, 44100, "silence", 6) // 4).astype(np.int32),
     2, 16, 44100, 4608, 12, 5, 0, 5),
    ("ood_tail1", lambda: S.pcm(4608 + 1, 2, 16, 44100, "white", 7), 2, 16, 44100, 4608, 12, 5, 0, 5),
    ("ood_tail3", lambda: S.pcm(4608 + 3, 2, 16, 44100, "white", 7), 2, 16, 44100, 4608, 12, 5, 0, 5),
    ("ood_tail5", lambda: S.pcm(4608 + 5, 2, 16, 44100, "white", 7), 2, 16, 44100, 4608, 12, 5, 0, 5),
    ("ood_tail8", lambda: S.pcm(4608 + 8, 2, 16, 44100, "white", 7), 2, 16, 44100, 4608, 12, 5, 0, 5),
]


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def f64_hex(x: float) -> str:
    return struct.pack(">d", float(x)).hex()


def main() -> None:
    out = {"python": sys.version.split()[0], "reference": "turlando/flac-py @ /root/reference",
           "cases": [], "ood": [], "kat": {}}
    for (name, cls, n, ch, bits, rate, bs, L, q, rmin, rmax, seed) in CASES:
        x = S.pcm(n, ch, bits, rate, cls, seed)
        chunks, recs = R.encode_introspect(x.tolist(), ch, rate, bits, bs, L, q, rmin, rmax)
        stream = b"".join(chunks)
        with open(os.path.join(HERE, name + ".flac"), "wb") as f:
            f.write(stream)
        subframes = []
        for r in recs:
            subframes.append({
                "type": r["type"],
                "order": (r["lpc_order"] if r["type"] == "lpc" else r["fixed_order"]),
                "fixed_order": r["fixed_order"],
                "fixed_err": r.get("fixed_err"),
                "lpc_order": r.get("lpc_order"),
                "lpc_shift": r.get("lpc_shift"),
                "lpc_coefs": r.get("lpc_coefs"),
                "lpc_total": r.get("lpc_total") if L > 0 else None,
                "quant": r["quant"] if L > 0 else None,
                "autoc": [f64_hex(a) for a in r["autoc"]] if L > 0 else None,
                "porder": r["porder"], "method": r["method"], "params": r["params"],
            })
        case = {
            "name": name, "cls": cls, "n": n, "channels": ch, "bits": bits, "rate": rate,
            "block_size": bs, "max_lpc_order": L, "qlp_precision": q, "rice_min": rmin,
            "rice_max": rmax, "seed": seed,
            "pcm_sha256": sha(x.tobytes()), "stream_sha256": sha(stream),
            "chunk_lengths": [len(c) for c in chunks], "subframes": subframes,
        }
        if ch == 2:
            rate_, bps_, ch_, total_, frames = R.decode(stream)
            dec = np.array(frames, dtype=np.int32).reshape(-1, 2)
            assert (dec == x).all(), name
            case["decode"] = {"rate": rate_, "bits": bps_, "channels": ch_, "total": total_,
                              "pcm_sha256": sha(dec.tobytes())}
        out["cases"].append(case)
        print(f"{name}: {len(stream)} bytes, {len(recs)} subframes, "
              f"{sum(1 for r in recs if r['type'] == 'lpc')} lpc")
    for (name, gen, ch, bits, rate, bs, L, q, rmin, rmax) in OOD:
        x = gen()
        try:
            R.encode_chunks(x.tolist(), x.shape[0], ch, rate, bits, bs, L, q, rmin, rmax)
            exc = None
        except Exception as e:  # noqa: BLE001
            exc = type(e).__name__
        np.save(os.path.join(HERE, name + ".npy"), x.astype(np.int32 if bits > 16 else np.int16))
        out["ood"].append({"name": name, "channels": ch, "bits": bits, "rate": rate,
                           "block_size": bs, "max_lpc_order": L, "qlp_precision": q,
                           "rice_min": rmin, "rice_max": rmax, "raises": exc})
        print(f"{name}: raises {exc}")
    # known answers for the primitives (computed by the reference)
    mods = R.load()
    crc, cn, enc = mods["flac.crc"], mods["flac.coded_number"], mods["flac.encoder"]
    out["kat"]["crc8_123456789"] = crc.crc8(b"123456789", 0x107)
    out["kat"]["crc16_123456789"] = crc.crc16(b"123456789", 0x18005)
    out["kat"]["coded_number"] = {str(v): cn.encode(v).hex() for v in
                                  (0, 1, 127, 128, 2047, 2048, 5742, 65535, 65536, 374999,
                                   2097151, 2097152, 67108863, 67108864, 2147483647)}
    out["kat"]["tukey"] = {str(n): [f64_hex(v) for v in enc.tukey(n, 0.5)] for n in (8, 27, 100)}
    rng = np.random.default_rng(5)
    sums = []
    for _ in range(8):
        v = (rng.standard_normal(777) * 10.0 ** rng.integers(-3, 12, 777)).tolist()
        sums.append({"x": [f64_hex(t) for t in v], "sum": f64_hex(sum(t for t in v))})
    out["kat"]["float_sum"] = sums
    hdrs = {}
    for (bs_, num) in ((4608, 0), (16384, 70000), (864, 200), (3618, 4), (255, 1), (192, 129)):
        p = enc.put_frame_header(mods["flac.common"].FrameHeader(
            mods["flac.common"].BlockingStrategy.Fixed, bs_, None, mods["flac.common"].Channels.L_R,
            None, num))
        hdrs[f"{bs_}:{num}"] = p.buffer.hex()
    out["kat"]["frame_header"] = hdrs
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote cases.json")


if __name__ == "__main__":
    main()
