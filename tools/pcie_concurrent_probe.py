#!/usr/bin/env python3
"""Aggregate host<->device copy ceiling of the box with N GPUs copying AT THE SAME TIME (one rank per
GPU, pinned host buffers, torch streams): H2D alone, D2H alone, both directions at once.  This is the
denominator of `e2e.frac_of_pcie_ceiling` in bench.py: the end-to-end encode moves its PCM in and its
stream out through the same host.

    torchrun --nproc-per-node N tools/pcie_concurrent_probe.py [--mb 128] > profiles/pcie_ceiling_<N>gpu.json
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.ones(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    sync_t = torch.zeros(1, device="cuda")

    def barrier():
        if world > 1:
            dist.all_reduce(sync_t)
        torch.cuda.synchronize()

    def timed(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.iters

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    res = {}
    for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("duplex", both, 2 * n)):
        dt = timed(fn)
        res[name] = {"aggregate_GBps": world * nbytes / dt / 1e9, "per_gpu_GBps": nbytes / dt / 1e9, "ms": dt * 1e3}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mib_per_copy": args.mb, "pinned": True, "concurrent": True, **res,
                          "how": "every rank copies at the same time (barrier, then iters copies per rank, max over ranks of the wall time)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
