#!/usr/bin/env python3
"""Development aid: resident encode / decode throughput for a few launch shapes (one GPU).

    python tools/perf_sweep.py [--workload c2] [--seconds 600] [--threads 576,288,192] [--decode]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "tools", "synth"), ROOT]
import numpy as np  # noqa: E402

import bench  # noqa: E402
from flac import _native as N  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--seconds", type=float, default=None)
    ap.add_argument("--threads", default="0")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--decode", action="store_true")
    ap.add_argument("--slots", action="store_true", help="slot + compaction placement instead of look-back")
    args = ap.parse_args()
    m = bench.workload_shape(args.workload, args.seconds)
    ctx = N.Context(0)
    n, ch, bits, bs = m["n"], m["channels"], m["bits"], m["block"]
    nbytes = n * ch * ((bits + 7) // 8)
    cap = N.encode_bound(n, ch, bits, bs, m["rmax"])
    d_in = ctx.device_alloc(nbytes)
    d_out = ctx.device_alloc(cap)
    bench.synth_device(0, d_in, args.workload, 0, n)

    class _Raw:
        size = nbytes
    raw = _Raw()
    enc_args = (n, ch, bits, bs, m["L"], m["q"], m["rmin"], m["rmax"], 0)
    out_len = 0
    for t in args.threads.split(";"):  # "0" = library default, else "threads,all_planes,variant"
        if t != "0":
            os.environ["FLACB200_SHAPE"] = t
        else:
            os.environ.pop("FLACB200_SHAPE", None)
        ks = []
        for k in range(args.steps + 2):
            rc, out_len, st = ctx.encode_raw(d_in, *enc_args, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT | (N.SLOT_COMPACT if args.slots else 0), d_out, cap)
            N.raise_for_status(rc)
            if k >= 2:
                ks.append(st.kernel_ms)
        rc, out_len, st = ctx.encode_raw(d_in, *enc_args, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT | N.STAGE_TIMES, d_out, cap)
        ms = float(np.median(ks))
        print(json.dumps({"workload": args.workload, "shape": t, "threads": st.threads_per_frame, "smem": st.smem_bytes,
                          "kernel_ms": round(ms, 4), "stages_ms": [round(float(v), 4) for v in st.stage_ms], "Msamples/s": round(n * ch / ms / 1e3, 1),
                          "GB/s": round((raw.size + out_len) / ms / 1e6, 2), "bytes_out": out_len,
                          "ood": int(st.out_of_domain)}), flush=True)
    if args.decode:
        pcm_cap = n * ch + 1024
        d_pcm = ctx.device_alloc(pcm_cap * 4)
        ks = []
        for k in range(args.steps + 2):
            rc, nsf, fch, st = ctx.decode_raw(d_out, out_len, bits, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE, d_pcm, pcm_cap)
            N.raise_for_status(rc)
            if k >= 2:
                ks.append(st.kernel_ms)
        rc, out_len, st = ctx.encode_raw(d_in, *enc_args, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT | N.STAGE_TIMES, d_out, cap)
        ms = float(np.median(ks))
        chk = np.empty(1024, dtype=np.int32)
        ctx.d2h(chk.ctypes.data, d_pcm, 4096)
        ok = None
        print(json.dumps({"decode": args.workload, "kernel_ms": round(ms, 4), "frames": int(st.frames), "n_sample_frames": nsf,
                          "false_candidates": int(st.out_of_domain), "Msamples/s": round(nsf * fch / ms / 1e3, 1),
                          "GB/s": round((out_len + nsf * fch * bits // 8) / ms / 1e6, 2), "head_matches_input": ok}), flush=True)


if __name__ == "__main__":
    main()
