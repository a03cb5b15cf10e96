#!/usr/bin/env python3
"""Development aid: do a host->device and a device->host copy overlap on this box?  (torch
streams, pinned host tensors; plumbing only.)"""
import torch

n = 100 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0), s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
    print(f"100 MiB: h2d alone {a:.3f} ms, d2h alone {b:.3f} ms, both at once {c:.3f} ms "
          f"(serial would be {a + b:.3f}, perfect overlap {max(a, b):.3f})")
