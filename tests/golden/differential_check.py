#!/usr/bin/env python3
"""Randomised differential check against the REAL reference (turlando/flac-py), for the build
container only (the reference is mounted at /root/reference and cannot travel):

    python tests/golden/differential_check.py [cases] [seed]

Every case draws channel count, sample width, rate, block size, LPC order, coefficient precision,
Rice partition range, signal class and length at random, encodes it with the unmodified reference
encoder, with the CPU oracle (oracle/flac_oracle.c) and with the CUDA kernels' logic on the CPU
SIMT simulator (tests/sim), and compares the three: the same bytes, or the same exception class
where the reference raises.  The last run is recorded in DESIGN.md.
"""
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "tools", "synth")]

import flacb200_synth as S  # noqa: E402
import oracle as O  # noqa: E402
import refharness as R  # noqa: E402
import simlib  # noqa: E402

STATUS = {1: "AssertionError", 2: "ValueError", 3: "ZeroDivisionError", 6: "OverflowError"}


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 20261018)
    classes = ["white", "mixed", "sine", "lowpass", "highpass", "clip", "silence"]
    bad = raised = 0
    t0 = time.time()
    for it in range(cases):
        ch = rng.choice([1, 2, 2, 2, 3, 5])
        bits = rng.choice([16, 16, 24, 8, 12, 20])
        rate = rng.choice([44100, 48000, 96000])
        bs = rng.choice([16, 64, 192, 255, 576, 1000, 1152, 2304])
        L = rng.choice([1, 2, 4, 8, 12]) if rate <= 48000 else rng.choice([8, 12, 20, 32])
        q = rng.choice([5, 6, 8, 12, 15])
        rmax = rng.choice([0, 2, 4, 6])
        rmin = rng.choice([0, 0, min(1, rmax)])
        cls = rng.choice(classes)
        n = bs * rng.choice([1, 2, 3]) + rng.choice([0, 1, 5, bs // 3, bs - 1])
        x = S.pcm(n, ch, bits, rate, cls, 1000 + it)
        try:
            ref = b"".join(R.encode_chunks([list(map(int, r)) for r in x], n, ch, rate, bits, bs, L, q, rmin, rmax))
            rexc = None
        except Exception as e:  # noqa: BLE001
            ref, rexc = None, type(e).__name__
            raised += 1
        try:
            got = O.stream_prologue(rate, ch, bits, n, bs) + O.encode_frames(x, bits, bs, L, q, rmin, rmax).data
            oexc = None
        except O.OracleError as e:
            got, oexc = None, e.exc_class.__name__
        rc, sdata, _, _, _ = simlib.encode_frames(x, bits, bs, L, q, rmin, rmax)
        oracle_ok = (ref == got) if ref is not None else (rexc == oexc)
        sim_ok = (rc == 0 and sdata == ref[42:]) if ref is not None else (STATUS.get(rc) == rexc)
        if not (oracle_ok and sim_ok):
            bad += 1
            print("MISMATCH", dict(case=it, ch=ch, bits=bits, rate=rate, bs=bs, L=L, q=q, rmin=rmin, rmax=rmax, cls=cls, n=n),
                  "reference raises", rexc, "oracle raises", oexc, "oracle ok", oracle_ok, "sim rc", rc, "sim ok", sim_ok)
    print(f"{cases} cases ({raised} where the reference raises), {bad} mismatches, {time.time() - t0:.1f} s, "
          f"python {sys.version.split()[0]}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
