#!/usr/bin/env python3
"""Development aid: host<->device copy rates of this box through the library's own copy calls
(pinned host memory), to tell whether the end-to-end encode is bound by the copies."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "flac-py_b200"), ROOT]
from flac import _native as N  # noqa: E402

ctx = N.Context(0)
for mb in (16, 106, 512):
    n = mb << 20
    pin = N.PinnedBuffer(n)
    pin.array[:] = 1
    d = ctx.device_alloc(n)
    for name, fn in (("h2d", lambda: ctx.h2d(d, pin.ptr, n)), ("d2h", lambda: ctx.d2h(pin.ptr, d, n))):
        fn()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        dt = (time.perf_counter() - t0) / 10
        print(f"{name} {mb} MiB: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.3f} ms)")
