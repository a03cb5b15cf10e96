#!/usr/bin/env python3
"""Development aid: one end-to-end encode (pinned host buffers on both sides) with FLACB200_TIMELINE=1:
when every chunk's copies and kernels ran.

    python tools/e2e_timeline.py [--workload c2]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "tools", "synth"), ROOT]
import numpy as np  # noqa: E402

import bench  # noqa: E402
import flacb200_synth as S  # noqa: E402
from flac import _native as N  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    args = ap.parse_args()
    m = bench.workload_shape(args.workload, None)
    ctx = N.Context(0)
    n, ch, bits, bs = m["n"], m["channels"], m["bits"], m["block"]
    nbytes = n * ch * ((bits + 7) // 8)
    cap = N.encode_bound(n, ch, bits, bs, m["rmax"])
    d_in = ctx.device_alloc(nbytes)
    bench.synth_device(0, d_in, args.workload, 0, n)
    h_in, h_out = N.PinnedBuffer(nbytes), N.PinnedBuffer(cap)
    ctx.d2h(h_in.ptr, d_in, nbytes)
    enc_args = (n, ch, bits, bs, m["L"], m["q"], m["rmin"], m["rmax"], 0)
    for k in range(6):
        if k == 5:
            os.environ["FLACB200_TIMELINE"] = "1"
        t0 = time.perf_counter()
        rc, out_len, st = ctx.encode_raw(h_in.ptr, *enc_args, N.LENIENT, h_out.ptr, cap)
        dt = (time.perf_counter() - t0) * 1e3
        N.raise_for_status(rc)
        print(f"call {k}: wall {dt:.3f} ms kernels {st.kernel_ms:.3f} h2d {st.h2d_ms:.3f} d2h {st.d2h_ms:.3f} out {out_len}", flush=True)


if __name__ == "__main__":
    main()
