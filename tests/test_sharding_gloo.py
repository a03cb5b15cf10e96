"""The N > 1 path on CPU: two gloo ranks each encode their frame range (kernels on the SIMT
simulator, since there is no GPU here), exchange one integer each, and the byte ranges placed at
the exchanged offsets are the single-rank stream."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank: int, world: int, port: int, out_dir: str):
    for p in (os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "oracle"), HERE):
        sys.path.insert(0, p)
    import torch
    import simlib
    from flac.shard import byte_offsets, sample_range
    from helpers import CASES, case_kwargs, case_pcm
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    case = CASES["many_frames"]
    x = case_pcm(case)
    bs = case["block_size"]
    s0, s1 = sample_range(x.shape[0], bs, rank, world)
    rc, data, offs, _, _ = simlib.encode_frames(x[s0:s1], case["bits"], first_frame_number=s0 // bs, **case_kwargs(case))
    assert rc == 0
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(data)], dtype=torch.int64))
    off = byte_offsets([int(s) for s in sizes])
    path = os.path.join(out_dir, "out.flac")
    dist.barrier()
    with open(path, "r+b") as f:  # every rank writes its own byte range, no data moves between ranks
        f.seek(off[rank])
        f.write(data)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_write_the_single_rank_stream(tmp_path):
    sys.path.insert(0, HERE)
    from helpers import CASES, case_stream
    case = CASES["many_frames"]
    ref = case_stream(case)
    path = tmp_path / "out.flac"
    path.write_bytes(ref[:42] + bytes(len(ref) - 42))
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert path.read_bytes() == ref


def test_frame_ranges_are_balanced_and_ordered():
    sys.path.insert(0, os.path.join(ROOT, "flac-py_b200"))
    from flac.shard import byte_offsets, frame_range, sample_range
    for n in (1, 7, 5743, 375000):
        for world in (1, 2, 4, 8):
            r = [frame_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert sample_range(4608 * 3 + 100, 4608, 1, 2) == (4608 * 2, 4608 * 3 + 100)
    assert byte_offsets([10, 20, 30]) == [42, 52, 72, 102]
