#!/usr/bin/env python3
"""Development aid: what does the encode cost when the serialiser stores the stream straight into pinned host
memory (zero copy over PCIe) instead of into HBM?  PCM resident in HBM in both cases.

    python tools/zero_copy_probe.py [--workload c2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "tools", "synth"), ROOT]
import numpy as np  # noqa: E402

import bench  # noqa: E402
from flac import _native as N  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--seconds", type=float, default=None)
    args = ap.parse_args()
    m = bench.workload_shape(args.workload, args.seconds)
    ctx = N.Context(0)
    n, ch, bits, bs = m["n"], m["channels"], m["bits"], m["block"]
    nbytes = n * ch * ((bits + 7) // 8)
    cap = N.encode_bound(n, ch, bits, bs, m["rmax"])
    d_in = ctx.device_alloc(nbytes)
    d_out = ctx.device_alloc(cap)
    h_out = N.PinnedBuffer(cap)
    h_ref = N.PinnedBuffer(cap)
    bench.synth_device(0, d_in, args.workload, 0, n)
    enc_args = (n, ch, bits, bs, m["L"], m["q"], m["rmin"], m["rmax"], 0)
    flags = N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT
    res = {}
    for name, out in (("hbm", d_out), ("pinned_host", h_out.ptr)):
        ks = []
        for k in range(7):
            rc, out_len, st = ctx.encode_raw(d_in, *enc_args, flags, out, cap)
            N.raise_for_status(rc)
            if k >= 2:
                ks.append(st.kernel_ms)
        rc, out_len, st = ctx.encode_raw(d_in, *enc_args, flags | N.STAGE_TIMES, out, cap)
        res[name] = {"kernel_ms": round(float(np.median(ks)), 4), "stages_ms": [round(float(v), 4) for v in st.stage_ms], "bytes_out": out_len}
    ctx.d2h(h_ref.ptr, d_out, out_len)
    res["same_bytes"] = bool((h_ref.array[:out_len] == h_out.array[:out_len]).all())
    res["GBps_to_host"] = round(out_len / res["pinned_host"]["kernel_ms"] / 1e6, 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
