#!/usr/bin/env python3
"""Benchmark of the FLAC encode hot path on B200 (BASELINE.json metric: encode Msamples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c2fixed]
    python bench.py --impl reference ...          # the CPU arm (oracle port on the host cores)
    torchrun --nproc-per-node N bench.py --gpus N # one rank per GPU, frame-range sharded

A step is one pass of the whole encode path over one stream of the workload (default: config 2
of BASELINE.json, 16-bit stereo 44.1 kHz, 10 minutes, block 4608, LPC order 12, qlp precision 5,
Rice partition orders 0..5; "samples" are channel-samples).  `value` is measured with the PCM
already resident in HBM and the encoded stream left in HBM; `e2e` goes through the C ABI with
pinned HOST buffers (H2D of the PCM and D2H of the stream inside the timed region).  With N
ranks every rank encodes its own frame range of an N-times longer stream (weak scaling; frames
are independent, so there is no collective on the data path, only the barrier and the max over
ranks of the device time).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools", "synth")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, channels, bits, rate, seconds, block, L, q, rmin, rmax, seed)
    "c1": ("16-bit mono 44.1 kHz 10 s, block 4608, fixed predictors only", 1, 16, 44100, 10, 4608, 0, 5, 0, 5, 0xF1AC0001),
    "c2": ("16-bit stereo 44.1 kHz 10 min, block 4608, max LPC order 12, qlp precision 5, Rice partition order 0..5",
           2, 16, 44100, 600, 4608, 12, 5, 0, 5, 0xF1AC0002),
    "c2fixed": ("16-bit stereo 44.1 kHz 10 min, block 4608, fixed predictors only, Rice partition order 0..5",
                2, 16, 44100, 600, 4608, 0, 5, 0, 5, 0xF1AC0002),
    "c3": ("24-bit stereo 96 kHz 1 h, block 16384, max LPC order 32, qlp precision 5, Rice partition order 0..8",
           2, 24, 96000, 3600, 16384, 32, 5, 0, 8, 0xF1AC0003),
    "c4": ("24-bit 8-channel 48 kHz 10 h corpus (375 000 frames, 41.5 GB of PCM), block 4608, max LPC order 12, qlp precision 5, "
           "Rice partition order 0..5", 8, 24, 48000, 36000, 4608, 12, 5, 0, 5, 0xF1AC0004),
}


def synth(name: str, rank: int = 0, seconds: float | None = None) -> tuple[np.ndarray, dict]:
    import flacb200_synth as S
    desc, ch, bits, rate, secs, bs, L, q, rmin, rmax, seed = WORKLOADS[name]
    if seconds is not None:
        secs = seconds
    n = int(rate * secs)
    parts = []
    step = 1 << 21
    for s0 in range(0, n, step):  # rank r takes frames [r*n, (r+1)*n) of the longer stream
        parts.append(S.pack_le(S.pcm(min(step, n - s0), ch, bits, rate, "mixed", seed, start=rank * n + s0), bits))
    raw = np.frombuffer(b"".join(parts), dtype=np.uint8)
    meta = dict(desc=desc, channels=ch, bits=bits, rate=rate, n=n, block=bs, L=L, q=q, rmin=rmin, rmax=rmax)
    return raw, meta


def workload_shape(name: str, seconds: float | None = None) -> dict:
    desc, ch, bits, rate, secs, bs, L, q, rmin, rmax, seed = WORKLOADS[name]
    if seconds is not None:
        secs = seconds
    return dict(desc=desc, channels=ch, bits=bits, rate=rate, n=int(rate * secs), block=bs, L=L, q=q, rmin=rmin, rmax=rmax, seed=seed)


def synth_device(device: int, dst_ptr: int, name: str, start: int, n: int) -> int:
    """The workload's PCM, sample frames [start, start + n), written by the GPU itself into its HBM
    (tools/synth/flacb200_synth.cu: bit for bit what synth() returns; tests/test_synth_device.py)."""
    import flacb200_synth as S
    desc, ch, bits, rate, secs, bs, L, q, rmin, rmax, seed = WORKLOADS[name]
    return S.pcm_device(device, dst_ptr, n, ch, bits, rate, "mixed", seed, start)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines, self.first = device, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        self.first = len(self.lines)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        lines = self.lines[self.first:] if len(self.lines) - self.first >= 2 else self.lines[-3:]
        for ln in lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_near_gpu(local_rank: int) -> str:
    """Best effort: run this rank on the CPUs of the GPU's NUMA node, so that its pinned host
    buffers (first touch) and the copy threads sit next to the GPU's PCIe root."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"cpus {spec} (NUMA node of GPU {local_rank})"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"
    return "not bound"


def ncu_traffic(workload: str, kernel: str):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return int(json.load(f)[workload][kernel]["bytes"])
    except Exception:  # noqa: BLE001
        return None


def peak_hbm() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_arm(name: str, rank: int, target_s: float = 15.0, threads: int | None = None) -> dict:
    """The CPU port of the reference's algorithm (oracle/flac_oracle.c, one POSIX thread per host
    core, contiguous frame ranges) on a bounded slice of the same workload."""
    import oracle as O
    import flacb200_synth as S
    desc, ch, bits, rate, secs, bs, L, q, rmin, rmax, seed = WORKLOADS[name]
    cores = threads or os.cpu_count() or 1
    probe = S.pcm(bs * cores, ch, bits, rate, "mixed", seed)
    t0 = time.perf_counter()
    oracle_in_domain(probe, bits, bs, L, q, rmin, rmax, threads=cores)
    per_round = max(time.perf_counter() - t0, 1e-4)
    rounds = int(max(1, min(target_s / per_round, (rate * secs) // (bs * cores))))
    x = S.pcm(bs * cores * rounds, ch, bits, rate, "mixed", seed)
    dt, passes = 0.0, 0
    nfr_ok = cores * rounds
    while dt < target_s * 0.6 or passes == 0:  # repeat the slice until enough CPU time is on the clock
        t0 = time.perf_counter()
        enc, nfr_ok = oracle_in_domain(x, bits, bs, L, q, rmin, rmax, threads=cores)
        dt += time.perf_counter() - t0
        passes += 1
    ms = x.shape[0] * ch * passes / dt / 1e6
    return {"value": ms, "unit": "Msamples/s", "cores": cores, "kind": "port", "_enc": enc, "_frames": nfr_ok,
            "sample": f"first {x.shape[0]} sample frames ({cores * rounds} frames, {x.shape[0] / rate:.1f} s of audio) of the "
                      f"workload, {passes} passes, oracle/flac_oracle.c with {cores} threads, {dt:.2f} s wall, {len(enc.data)} bytes out per pass",
            "seconds": dt, "realtime_x": x.shape[0] * passes / rate / dt}


def oracle_in_domain(x, bits, bs, L, q, rmin, rmax, **kw):
    """Oracle encode of the frames in front of the first one the reference raises on (the oracle has no
    lenient mode: parity is only defined inside the reference's domain)."""
    import oracle as O
    try:
        return O.encode_frames(x, bits, bs, L, q, rmin, rmax, **kw), (x.shape[0] + bs - 1) // bs
    except O.OracleError as e:
        k = max(int(e.frame), 0)
        return O.encode_frames(x[:k * bs], bits, bs, L, q, rmin, rmax, **kw), k


def parity_check(ctx, N, name: str, start: int, first_frame: int, n: int, d_in: int, d_out: int, cap: int, enc_args: tuple,
                 oracle_enc=None, oracle_frames: int = 0, info_frames: int = 48) -> dict:
    """The stream this rank's GPU produced against the CPU oracle, on the same PCM:
    * bytes of the first `oracle_frames` frames (the slice the CPU arm encoded anyway; a smaller slice
      is encoded here when there is none, i.e. on ranks > 0 and with --no-cpu-baseline);
    * per subframe of the first `info_frames` frames: autocorrelation lags (bit patterns), quantised
      coefficients + shifts of every order, the decisions (type, order, partition order, parameters).
    north_star: "identical quantized coefficients wherever the fp64 autocorrelation agrees, and report a
    mismatch rate otherwise" -- coefficient sets are only compared where every lag agrees."""
    import oracle as O
    import flacb200_synth as S
    desc, ch, bits, rate, secs, bs, L, q, rmin, rmax, seed = WORKLOADS[name]
    nfr = (n + bs - 1) // bs
    offs = np.zeros(nfr + 1, dtype=np.uint64)
    flags = N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT
    rc, out_len, st = ctx.encode_raw(d_in, *enc_args, flags, d_out, cap, frame_offsets=offs)
    N.raise_for_status(rc, "encode (parity)")
    if oracle_enc is None:
        oracle_frames = min(nfr, 64)
        x = S.pcm(min(n, bs * oracle_frames), ch, bits, rate, "mixed", seed, start=start)
        oracle_enc, oracle_frames = oracle_in_domain(x, bits, bs, L, q, rmin, rmax, first_frame_number=first_frame,
                                                     threads=os.cpu_count() or 1)
    k = min(oracle_frames, nfr)
    want = np.frombuffer(oracle_enc.data, dtype=np.uint8)[:int(oracle_enc.frame_offsets[k])]
    got = np.empty(int(offs[k]), dtype=np.uint8)
    if got.size:
        ctx.d2h(got.ctypes.data, d_out, got.size)
    mlen = min(got.size, want.size)
    byte_mismatches = int(np.count_nonzero(got[:mlen] != want[:mlen])) + abs(got.size - want.size)
    frame_mismatches = int(np.count_nonzero(offs[:k + 1] != oracle_enc.frame_offsets[:k + 1]))
    out = {"frames_checked": int(k), "bytes_checked": int(want.size), "byte_mismatches": byte_mismatches,
           "frame_offset_mismatches": frame_mismatches}
    # stage values on a slice (needs the per-subframe records: a separate small encode on both sides)
    kf = min(info_frames, nfr)
    ns = min(n, kf * bs)
    if L > 0 and kf > 0:
        x = S.pcm(ns, ch, bits, rate, "mixed", seed, start=start)
        ref, kf = oracle_in_domain(x, bits, bs, L, q, rmin, rmax, first_frame_number=first_frame, want_info=True)
        ns = min(ns, kf * bs)
        info = (N.SubframeInfo * max(kf * ch, 1))()
        if kf:
            rc, _, _ = ctx.encode_raw(d_in, ns, *enc_args[1:], flags, d_out, cap, info=info)
            N.raise_for_status(rc, "encode (parity, stage values)")
        lags = lag_bad = sets = set_bad = dec_bad = 0
        for a, r in zip(list(info)[:kf * ch], ref.info or []):
            la = [a.autoc[i] for i in range(L + 1)]
            lr = [r.autoc[i] for i in range(L + 1)]
            bad = sum(1 for u, v in zip(la, lr) if np.float64(u).tobytes() != np.float64(v).tobytes())
            lags += L + 1
            lag_bad += bad
            if bad == 0:
                for o in range(L):
                    sets += 1
                    if (a.lpc_ncoef[o] != r.lpc_ncoef[o] or a.lpc_shift[o] != r.lpc_shift[o]
                            or list(a.lpc_q[o][:a.lpc_ncoef[o]]) != list(r.lpc_q[o][:r.lpc_ncoef[o]])):
                        set_bad += 1
            if (a.type, a.order, a.porder, a.method, a.bits) != (r.type, r.order, r.porder, r.method, r.bits):
                dec_bad += 1
        out.update({"subframes_checked": int(kf * ch), "lpc_lags_checked": lags, "lpc_lag_mismatches": lag_bad,
                    "lpc_coef_sets_checked": sets, "lpc_coef_mismatches": set_bad, "decision_mismatches": dec_bad,
                    "lpc_lag_mismatch_rate": lag_bad / lags if lags else 0.0})
    else:
        out.update({"subframes_checked": 0, "lpc_lag_mismatches": 0, "lpc_coef_mismatches": 0, "decision_mismatches": 0})
    return out


def api_and_cli(name: str, m: dict, raw: np.ndarray, flac_bytes: int) -> dict:
    """`e2e_api`: bytes in -> b"".join(flac.encoder.encode_pcm(...)) -> bytes out, the drop-in's bulk entry
    (what the command line calls); `cli`: `python -m flac encode / decode` on a WAV file in /dev/shm, the
    time the tool prints (the reference's own timing line, flac/__main__.py:108-109, :54-55) and the wall
    time of the process (interpreter start, CUDA context and library load included)."""
    import tempfile
    import wave
    from flac.encoder import EncoderParameters, encode_pcm
    ch, bits, bs, rate, n = m["channels"], m["bits"], m["block"], m["rate"], m["n"]
    params = EncoderParameters(block_size=bs, rice_partition_order=range(m["rmin"], m["rmax"] + 1),
                               lpc_order=range(m["L"] + 1), qlp_precision=m["q"])
    payload = raw.tobytes()  # a Python bytes object: pageable memory, like wave.readframes() returns
    ts, size = [], 0
    for _ in range(4):
        t0 = time.perf_counter()
        stream = b"".join(encode_pcm(rate, bits, ch, payload, params, strict=False))
        ts.append(time.perf_counter() - t0)
        size = len(stream)
    dt = float(np.median(ts[1:]))
    out = {"e2e_api": {"value": n * ch / dt / 1e6, "unit": "Msamples/s", "ms": dt * 1e3, "bytes_in": len(payload), "bytes_out": size,
                       "stream_matches_resident_encode": bool(size == 42 + flac_bytes),
                       "call": "b''.join(flac.encoder.encode_pcm(rate, bits, channels, payload_bytes, parameters))"}}
    if bits == 16 and ch == 2:  # the only layout the reference's command line parses correctly
        tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        wav, flac, back = (os.path.join(tmp, f) for f in ("in.wav", "out.flac", "back.wav"))
        try:
            with wave.open(wav, "wb") as w:
                w.setnchannels(ch), w.setsampwidth(2), w.setframerate(rate)
                w.writeframes(payload)
            env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "flac-py_b200"))
            rec = {}
            for action, a, b in (("encode", wav, flac), ("decode", flac, back)):
                cmd = [sys.executable, "-m", "flac", action, a, b]
                if action == "encode":
                    cmd += ["-b", str(bs), "-l", str(m["L"]), "-q", str(m["q"]), "-r", f"{m['rmin']},{m['rmax']}"]
                t0 = time.perf_counter()
                r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
                wall = time.perf_counter() - t0
                printed = [ln for ln in r.stdout.splitlines() if "completed in" in ln]
                secs = float(printed[-1].split("completed in")[1].split()[0]) if printed else None
                rec[action] = {"printed_s": secs, "process_wall_s": wall, "rc": r.returncode,
                               "Msamples_per_s_printed": n * ch / secs / 1e6 if secs else None}
                if r.returncode != 0:
                    rec[action]["stderr_tail"] = r.stderr[-400:]
            with wave.open(back, "rb") as w:
                same = w.readframes(w.getnframes()) == payload
            rec["flac_bytes"] = os.path.getsize(flac)
            rec["round_trip_bit_exact"] = bool(same)
            rec["stream_matches_resident_encode"] = bool(rec["flac_bytes"] == 42 + flac_bytes)
            out["cli"] = rec
        finally:
            for f in (wav, flac, back):
                if os.path.exists(f):
                    os.remove(f)
            os.rmdir(tmp)
    return out


def pcie_ceiling(world: int, local_rank: int, mb: int = 128, iters: int = 10) -> dict:
    """Aggregate host<->device copy rate of THIS box with all ranks copying at the same time (pinned host
    memory, plain copies, no kernels): H2D alone, D2H alone, both at once.  The end-to-end encode can go
    no faster than its bytes through these (tools/pcie_concurrent_probe.py is the same thing stand-alone)."""
    import torch
    import torch.distributed as dist
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device=f"cuda:{local_rank}")
    d_b = torch.ones(n, dtype=torch.uint8, device=f"cuda:{local_rank}")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    sync_t = torch.zeros(1, device=f"cuda:{local_rank}")

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    out = {}
    for name, fns, nbytes in (("h2d", (h2d,), n), ("d2h", (d2h,), n), ("duplex", (h2d, d2h), 2 * n)):
        for f in fns:
            f()
        if world > 1:
            dist.all_reduce(sync_t)
        best = 0.0
        for _ in range(3):  # a ceiling: the best of three rounds (a round now and then runs at two thirds of the rate)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                for f in fns:
                    f()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local_rank}")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = max(best, world * nbytes * iters / float(t.item()) / 1e9)
        out[name] = best
    return {"h2d_GBps": out["h2d"], "d2h_GBps": out["d2h"], "duplex_GBps": out["duplex"], "n_gpus": world,
            "how": f"{mb} MiB pinned copies, every rank at the same time, max over ranks; aggregate over the ranks; best of three rounds"}


def c4_strong(ctx, N, rank: int, world: int, local_rank: int, barrier, steps: int = 2) -> dict:
    """BASELINE.json config 4 as it is stated: the 8-channel 24-bit 48 kHz 10 h corpus (375 000 frames of
    4608, 41.5 GB of PCM), frame-sharded over the ranks -- STRONG scaling: rank r generates frames
    [r F / N, (r + 1) F / N) on its own GPU and encodes them with their global frame numbers; no data
    moves between ranks (the streams concatenate: `shard_check`).  Device-resident, like `value`."""
    import torch
    import torch.distributed as dist
    desc, ch, bits, rate, secs, bs, L, q, rmin, rmax, seed = WORKLOADS["c4"]
    F = rate * secs // bs
    f0, f1 = F * rank // world, F * (rank + 1) // world
    nfr, n = f1 - f0, (f1 - f0) * bs
    nbytes = n * ch * 3
    cap = N.encode_bound(n, ch, bits, bs, rmax)
    d_in, d_out = ctx.device_alloc(nbytes), ctx.device_alloc(cap)
    try:
        t0 = time.perf_counter()
        synth_device(local_rank, d_in, "c4", f0 * bs, n)
        gen_s = time.perf_counter() - t0
        enc_args = (n, ch, bits, bs, L, q, rmin, rmax, f0)
        flags = N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT
        rc, out_len, st = ctx.encode_raw(d_in, *enc_args, flags, d_out, cap)  # warm-up: buffers, first launch
        N.raise_for_status(rc, "encode (c4_strong)")
        barrier()
        ctx.timer_begin()
        for _ in range(steps):
            rc, out_len, st = ctx.encode_raw(d_in, *enc_args, flags, d_out, cap)
            N.raise_for_status(rc, "encode (c4_strong)")
        dev_ms = ctx.timer_end() / steps
        barrier()
        try:
            par = parity_check(ctx, N, "c4", f0 * bs, f0, n, d_in, d_out, cap, enc_args, info_frames=8)
            par = {k: par[k] for k in ("frames_checked", "byte_mismatches", "lpc_lag_mismatches", "lpc_coef_mismatches", "decision_mismatches")}
        except Exception as e:  # noqa: BLE001
            par = {"error": f"{type(e).__name__}: {e}"}
    finally:
        ctx.device_free(d_in)
        ctx.device_free(d_out)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    tot = torch.tensor([float(out_len), float(st.out_of_domain), float(par.get("byte_mismatches", 1)),
                        float(par.get("frames_checked", 0)), float(par.get("lpc_coef_mismatches", 1))], dtype=torch.float64,
                       device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms = float(t.item())
    samples = F * bs * ch
    return {"workload": desc, "scaling": "strong", "n_gpus": world, "frames": int(F), "frames_per_rank": int(nfr),
            "value": samples / ms / 1e3, "unit": "Msamples/s", "ms_per_step": ms, "realtime_x": F * bs / rate / (ms / 1e3),
            "pcm_bytes": int(F * bs * ch * 3), "stream_bytes": int(tot[0].item()),
            "frames_outside_reference_domain": int(tot[1].item()),
            "parity": {"frames_checked": int(tot[3].item()), "byte_mismatches": int(tot[2].item()),
                       "lpc_coef_mismatches": int(tot[4].item())},
            "generated_on_device_s_rank0": gen_s, "timing": f"CUDA events, max over ranks, mean of {steps} steps, PCM and stream resident in HBM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--copies", type=int, default=3, help="distinct device copies rotated between steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seconds", type=float, default=None, help="override the stream length (debugging)")
    ap.add_argument("--no-c4-strong", action="store_true", help="skip the config-4 strong-scaling record")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    desc = WORKLOADS[args.workload][0]

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        vals = []
        info = None
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_arm(args.workload, 0, target_s=1.0)
        for _ in range(steps):
            info = cpu_arm(args.workload, 0, target_s=max(2.0, 30.0 / steps))
            vals.append(info)
        tot_samples = sum(v["value"] * v["seconds"] for v in vals)
        tot_s = sum(v["seconds"] for v in vals)
        v = tot_samples / tot_s
        rate = WORKLOADS[args.workload][3] * WORKLOADS[args.workload][1]
        line = {"impl": "reference", "metric": "encode_throughput", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64 + f64 (LPC analysis)", "data": "synthetic",
                "config": {"workload": desc, "note": "bounded slice per step; CPU port of the reference algorithm "
                           "(the reference itself is pure Python, about 1000x slower: 0.03 Msamples/s, BASELINE.md)"},
                "realtime_x": v * 1e6 / rate,
                "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")} | {"value": v},
                "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # Everything but the final JSON line goes to stderr: libraries print banners on fd 1
    # (NCCL's version line, for one), and the driver reads exactly one JSON line from stdout.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from flac import _native as N

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    host_binding = bind_near_gpu(local_rank) if world > 1 else "single rank, not bound"
    sync_t = torch.zeros(1, device=f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.all_reduce(sync_t)
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # nvidia-smi takes about a second to come up: start it before the synthesis
    m = workload_shape(args.workload, args.seconds)
    ch, bits, bs = m["channels"], m["bits"], m["block"]
    n = m["n"]
    cbytes = (bits + 7) // 8
    ctx = N.Context(local_rank)
    pcm_bytes = n * ch * cbytes
    cap = N.encode_bound(n, ch, bits, bs, m["rmax"])
    nfr = (n + bs - 1) // bs
    # The PCM is generated on the device (rank r: sample frames [r n, (r + 1) n) of the N times longer stream).
    # Inputs of more than 3 GiB (config 4: 41.5 GB) get one device copy (they are far larger than L2 anyway) and
    # no host copy: their end-to-end leg runs on the first GiB of PCM.
    big = pcm_bytes > (3 << 30)
    copies = 1 if big else max(1, args.copies)
    d_in = [ctx.device_alloc(pcm_bytes) for _ in range(copies)]
    d_out = [ctx.device_alloc(cap) for _ in range(copies)]
    for d in d_in:
        synth_device(local_rank, d, args.workload, rank * n, n)
    e2e_frames = nfr if not big else max(1, (1 << 30) // (bs * ch * cbytes))
    e2e_n = n if not big else e2e_frames * bs
    e2e_bytes = e2e_n * ch * cbytes
    e2e_cap = N.encode_bound(e2e_n, ch, bits, bs, m["rmax"])
    pin_in = N.PinnedBuffer(e2e_bytes)
    ctx.d2h(pin_in.ptr, d_in[0], e2e_bytes)
    raw = pin_in.array  # the (first GiB of the) PCM on the host
    pin_out = N.PinnedBuffer(e2e_cap)
    enc_args = (n, ch, bits, bs, m["L"], m["q"], m["rmin"], m["rmax"], rank * nfr)
    e2e_args = (e2e_n,) + enc_args[1:]

    def step_resident(k: int, extra_flags: int = 0):
        rc, out_len, st = ctx.encode_raw(d_in[k % copies], *enc_args, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT | extra_flags,
                                         d_out[k % copies], cap)
        N.raise_for_status(rc, "encode (resident)")
        return out_len, st

    def step_e2e():
        rc, out_len, st = ctx.encode_raw(pin_in.ptr, *e2e_args, N.LENIENT, pin_out.ptr, e2e_cap)
        N.raise_for_status(rc, "encode (host buffers)")
        return out_len, st

    # ---- resident (kernel) measurement ----
    for k in range(args.warmup):
        out_len, st = step_resident(k)
    barrier()
    if rank == 0:
        sampler.mark()   # samples from here on fall into the timed region
    launches0 = N.lib().flacb_kernel_launches()
    kernel_ms, stage_ms = [], []
    ctx.timer_begin()
    t0 = time.perf_counter()
    for k in range(args.steps):
        out_len, st = step_resident(args.warmup + k)
        kernel_ms.append(st.kernel_ms)
    dev_ms = ctx.timer_end()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = N.lib().flacb_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    flac_bytes = out_len
    # The timed steps launch their stages back to back (programmatic dependent launch: the start of a stage
    # overlaps the end of the one in front), so no event can sit between two stages there.  The stages' own
    # times, for the roofline of the dominant kernel, come from a few extra steps with an event between the
    # stages (FLACB_STAGE_TIMES), outside the timed region.
    for k in range(max(3, min(args.steps, 8))):
        _, st_s = step_resident(k, N.STAGE_TIMES)
        stage_ms.append([float(v) for v in st_s.stage_ms])
    stage_ms = stage_ms[1:]

    # ---- end to end through the C ABI with host buffers ----
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e_len, e_st = step_e2e()
    ctx.sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    assert big or e_len == flac_bytes
    try:
        pcie = pcie_ceiling(world, local_rank)
    except Exception as e:  # noqa: BLE001
        pcie = {"error": f"{type(e).__name__}: {e}"}

    # ---- the path users call: the package's bulk entry with Python objects (pageable memory) on both sides,
    # and the reference's own command line on a WAV file (rank 0, one GPU) ----
    api_info = None
    if rank == 0 and world == 1 and not big:
        try:
            api_info = api_and_cli(args.workload, m, raw, flac_bytes)
        except Exception as e:  # noqa: BLE001
            api_info = {"error": f"{type(e).__name__}: {e}"}

    # ---- the CPU arm (rank 0, one GPU): the oracle port on a bounded slice of the same PCM; its bytes are
    # the reference for the parity check below ----
    cb_arm = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb_arm = cpu_arm(args.workload, 0)

    # ---- parity of the benchmarked stream itself (every rank checks its own frame range) ----
    last = (args.warmup + args.steps - 1) % copies
    try:
        parity = parity_check(ctx, N, args.workload, rank * n, rank * nfr, n, d_in[last], d_out[last], cap, enc_args,
                              cb_arm["_enc"] if cb_arm else None, cb_arm["_frames"] if cb_arm else 0,
                              info_frames=48 if bs <= 8192 else 12)
    except Exception as e:  # noqa: BLE001
        parity = {"error": f"{type(e).__name__}: {e}"}
    shard_check = None
    if world > 1:
        import hashlib
        mine = np.empty(flac_bytes, dtype=np.uint8)
        ctx.d2h(mine.ctypes.data, d_out[last], flac_bytes)
        digests = [None] * world
        dist.all_gather_object(digests, (int(flac_bytes), hashlib.sha256(mine.tobytes()).hexdigest(),
                                         {k: v for k, v in parity.items() if isinstance(v, (int, float))}))
        if rank == 0:
            # rank 0 encodes the last rank's frame range on its own GPU: a sharded encode must give the bytes a
            # single GPU gives for that range (frame numbers included), so the ranks' streams concatenate to the
            # single stream
            r = world - 1
            synth_device(local_rank, d_in[0], args.workload, r * n, n)
            rc, len_r, _ = ctx.encode_raw(d_in[0], *enc_args[:-1], r * nfr, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT, d_out[0], cap)
            N.raise_for_status(rc, "encode (shard check)")
            other = np.empty(len_r, dtype=np.uint8)
            ctx.d2h(other.ctypes.data, d_out[0], len_r)
            shard_check = {"rank_reencoded_on_rank0": r, "bytes": int(len_r),
                           "identical": bool(len_r == digests[r][0] and hashlib.sha256(other.tobytes()).hexdigest() == digests[r][1]),
                           "stream_bytes_per_rank": [d[0] for d in digests]}
            tot = {}
            for d in digests:
                for k, v in d[2].items():
                    if k != "lpc_lag_mismatch_rate":
                        tot[k] = tot.get(k, 0) + v
            if tot.get("lpc_lags_checked"):
                tot["lpc_lag_mismatch_rate"] = tot["lpc_lag_mismatches"] / tot["lpc_lags_checked"]
            parity = tot | {"ranks": world}

    # ---- config 5 on the side (rank 0, stereo streams: the reference decoder's domain): decode the stream
    # just produced back to PCM; the WHOLE result is compared with the input; not part of `value` ----
    decode_info = None
    if rank == 0 and ch == 2 and not big:
        try:
            import oracle as O
            pcm_cap = n * ch + 1024
            d_pcm = ctx.device_alloc(pcm_cap * 4)
            dms = []
            for _ in range(4):
                rc, nsf, fch, dst = ctx.decode_raw(d_out[last], flac_bytes, bits, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE, d_pcm, pcm_cap)
                N.raise_for_status(rc, "decode (resident)")
                dms.append(dst.kernel_ms)
            src = raw.reshape(-1, cbytes).astype(np.int32)
            want = src[:, 0] | (src[:, 1] << 8) | ((src[:, 2] << 16) if cbytes == 3 else 0)
            want = (want ^ (1 << (8 * cbytes - 1))) - (1 << (8 * cbytes - 1))
            # end to end: stream in pinned host memory -> interleaved int32 PCM in pinned host memory
            pin_pcm = N.PinnedBuffer(pcm_cap * 4)
            ctx.d2h(pin_out.ptr, d_out[last], flac_bytes)
            e_ms = []
            for _ in range(3):
                t0 = time.perf_counter()
                rc, nsf2, fch2, dst2 = ctx.decode_raw(pin_out.ptr, flac_bytes, bits, 0, pin_pcm.ptr, pcm_cap)
                e_ms.append((time.perf_counter() - t0) * 1e3)
                N.raise_for_status(rc, "decode (host buffers)")
            got = pin_pcm.array.view(np.int32)[:nsf2 * fch2]
            exact = bool(nsf2 == n and fch2 == ch and got.size == want.size and (got == want).all())
            dk = float(np.median(dms[1:]))
            de = float(np.median(e_ms[1:]))
            # CPU arm of the decode: the oracle's decoder (one thread, like the reference) on the first frames
            k = min(nfr, 256)
            offs = np.zeros(nfr + 1, dtype=np.uint64)
            rc, _, _ = ctx.encode_raw(d_in[last], *enc_args, N.INPUT_ON_DEVICE | N.OUTPUT_ON_DEVICE | N.LENIENT, d_out[last], cap, frame_offsets=offs)
            head_stream = O.stream_prologue(m["rate"], ch, bits, k * bs, bs) + pin_out.array[:int(offs[k])].tobytes()
            t0 = time.perf_counter()
            dec = O.decode_stream(head_stream)
            cpu_s = time.perf_counter() - t0
            algo_d = flac_bytes + n * ch * cbytes
            decode_info = {"value": nsf * fch / dk / 1e3, "unit": "Msamples/s", "kernel_ms": dk, "sample_frames": int(nsf),
                           "bit_exact": exact, "samples_compared": int(want.size),
                           "e2e": {"value": nsf2 * fch2 / de / 1e3, "unit": "Msamples/s", "ms": de, "h2d_bytes": int(flac_bytes),
                                   "d2h_bytes": int(nsf2 * fch2 * 4), "h2d_ms": dst2.h2d_ms, "d2h_ms": dst2.d2h_ms},
                           "roofline": {"bound": "hbm", "achieved": algo_d / dk / 1e6, "peak": peak_hbm()[0], "unit": "GB/s",
                                        "frac": algo_d / dk / 1e6 / peak_hbm()[0], "algorithmic_bytes": int(algo_d)},
                           "cpu_baseline": {"value": k * bs * ch / cpu_s / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
                                            "sample": f"first {k} frames of the stream, oracle/flac_oracle.c decoder, one thread, {cpu_s:.2f} s",
                                            "bit_exact": bool((np.asarray(dec[4]).reshape(-1) == want[:k * bs * ch]).all())},
                           "note": "BASELINE.json config 5 when the workload is c2: GPU decode of the stream encoded above"}
        except Exception as e:  # noqa: BLE001  (never fail the encode bench over the side measurement)
            decode_info = {"error": f"{type(e).__name__}: {e}"}

    # ---- BASELINE.json config 4, strong scaling over the ranks (a record of its own on the default line) ----
    c4_info = None
    if args.workload == "c2" and not args.no_c4_strong and args.seconds is None:
        try:
            c4_info = c4_strong(ctx, N, rank, world, local_rank, barrier)
        except Exception as e:  # noqa: BLE001
            c4_info = {"error": f"{type(e).__name__}: {e}"}

    # max over ranks of the device time; totals over ranks
    t = torch.tensor([dev_ms, e2e_ms, wall_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = (float(v) for v in t.tolist())
    samples_per_step = n * ch * world
    value = samples_per_step * args.steps / (dev_ms / 1e3) / 1e6
    e2e_value = e2e_n * ch * world * e2e_steps / (e2e_ms / 1e3) / 1e6

    if rank == 0:
        peak, peak_src = peak_hbm()
        k_ms = float(np.mean(kernel_ms))
        stages = [float(v) for v in np.mean(np.array(stage_ms), axis=0)]
        stage_names = ["analyse_kernel", "lpc_coefficients_kernel", "choose_kernel", "serialise_kernel"]
        dom = int(np.argmax(stages))
        # every stage but the coefficient kernel reads the step's PCM once; the serialising kernel
        # also writes the stream, the others a few hundred bytes per subframe (not counted)
        algo = pcm_bytes + flac_bytes
        dom_algo = algo if dom == 3 else pcm_bytes
        achieved = dom_algo / (stages[dom] / 1e3) / 1e9
        pipeline = algo / (k_ms / 1e3) / 1e9
        line = {
            "metric": "encode_throughput", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32/int64 + f64 (LPC analysis)", "data": "synthetic",
            "config": {"workload": desc, "samples_per_step_per_gpu": n * ch, "frames_per_step_per_gpu": nfr,
                       "l2": f"{copies} distinct device cop{'ies' if copies > 1 else 'y'} of input and output rotated between steps "
                             f"({copies * (pcm_bytes + cap) / 1e6:.0f} MB > 126 MB L2)",
                       "data_generated": "on the device (tools/synth/flacb200_synth.cu), bit-identical to the numpy generator",
                       "sharding": "frame ranges, one rank per GPU, no data-path collective",
                       "host_binding_rank0": host_binding,
                       "threads_per_frame": st.threads_per_frame, "smem_bytes_per_frame": st.smem_bytes},
            "realtime_x": value * 1e6 / (m["rate"] * ch),
            "compression_ratio": flac_bytes / pcm_bytes,
            "subframes": {"fixed": [int(v) for v in st.subframes_fixed], "lpc": int(st.subframes_lpc)},
            "frames_outside_reference_domain": int(st.out_of_domain),
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload, stage_names[dom]) if args.seconds is None else None, "peak_source": peak_src,
                         "kernel": stage_names[dom], "kernel_ms": stages[dom], "algorithmic_bytes": int(dom_algo),
                         "stages_ms": dict(zip(stage_names, stages)), "stages_sum_ms": float(sum(stages)),
                         "stages_measured": "separate steps with an event between the stages (FLACB_STAGE_TIMES); the timed steps overlap them",
                         "pipeline": {"kernel_ms": k_ms, "algorithmic_bytes": int(algo), "achieved": pipeline,
                                      "frac": pipeline / peak},
                         "note": "dominant kernel of the step (CUDA events around each stage on the launching stream); "
                                 "pipeline = all stages of the step together.  LPC configs are integer/fp64-ALU bound by "
                                 "the reference's semantics (all orders 1..L evaluated on every sample): see DESIGN.md"},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(e2e_bytes * world),
                    "d2h_bytes_per_step": int(e_len * world), "ms_per_step": e2e_ms / e2e_steps,
                    "sample": "the whole workload" if not big else f"the first {e2e_frames} frames ({e2e_bytes / 1e9:.2f} GB of PCM) of the workload",
                    "h2d_ms": e_st.h2d_ms, "kernel_ms": e_st.kernel_ms, "d2h_ms": e_st.d2h_ms,
                    "realtime_x": e2e_value * 1e6 / (m["rate"] * ch)},
            "gpu_launches": int(launches), "clocks": clocks, "parity": parity, "decode": decode_info,
        }
        if "error" not in pcie:
            # lower bound of an end-to-end step on this box: the PCM in and the stream out at the measured ceilings,
            # both directions at once
            h_ms = e2e_bytes * world / pcie["h2d_GBps"] / 1e6
            d_ms = e_len * world / pcie["d2h_GBps"] / 1e6
            x_ms = (e2e_bytes + e_len) * world / pcie["duplex_GBps"] / 1e6
            bound = max(h_ms, d_ms, x_ms)
            line["e2e"]["pcie_ceiling"] = pcie
            line["e2e"]["pcie_bound_ms"] = bound
            line["e2e"]["frac_of_pcie_ceiling"] = bound / (e2e_ms / e2e_steps)
            line["e2e"]["saturates_first"] = "h2d" if h_ms >= max(d_ms, x_ms) else ("d2h" if d_ms >= x_ms else "both directions together")
        else:
            line["e2e"]["pcie_ceiling"] = pcie
        if api_info is not None:
            line.update(api_info)
            if "e2e_api" in api_info:
                line["e2e_api"]["frac_of_e2e"] = api_info["e2e_api"]["value"] / e2e_value
        if shard_check is not None:
            line["shard_check"] = shard_check
        if c4_info is not None:
            line["c4_strong"] = c4_info
        line["cpu_baseline"] = {k: cb_arm[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb_arm else None
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
