#!/usr/bin/env python3
"""Development aid: which frames of a bench workload lie outside the reference's domain, and why
(strict encodes from the frame behind the last one found; run on the GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "tools", "synth"), ROOT]
import bench  # noqa: E402
import flacb200_synth as S  # noqa: E402
from flac import _native as N  # noqa: E402

name, want = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 10
m = bench.workload_shape(name)
n, ch, bits, bs = m["n"], m["channels"], m["bits"], m["block"]
ctx = N.Context(0)
cb = (bits + 7) // 8
d_in = ctx.device_alloc(n * ch * cb)
cap = N.encode_bound(n, ch, bits, bs, m["rmax"])
d_out = ctx.device_alloc(cap)
bench.synth_device(0, d_in, name, 0, n)
F = (n + bs - 1) // bs
f0, found = 0, []
while f0 < F and len(found) < want:
    rc, out_len, st = ctx.encode_raw(d_in + f0 * bs * ch * cb, n - f0 * bs, ch, bits, bs, m["L"], m["q"], m["rmin"], m["rmax"], f0,
                                     N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE, d_out, cap)
    if rc == 0:
        break
    bad = f0 + int(st.first_bad_frame)
    found.append((bad, rc, S.CLASSES[(bad * bs // S.SEGMENT) % 6], S.CLASSES[((bad + 1) * bs - 1) // S.SEGMENT % 6]))
    print(found[-1], flush=True)
    f0 = bad + 1
print("frames", F, "found", len(found))
