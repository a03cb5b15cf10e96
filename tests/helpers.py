"""Shared helpers for the parity tests (golden fixtures, case tables)."""
from __future__ import annotations

import hashlib
import json
import os
import struct

import numpy as np

import flacb200_synth as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN, "cases.json")) as _f:
    GOLD = json.load(_f)

CASES = {c["name"]: c for c in GOLD["cases"]}
OOD = {c["name"]: c for c in GOLD["ood"]}


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def hex_f64(h: str) -> float:
    return struct.unpack(">d", bytes.fromhex(h))[0]


def case_pcm(case: dict) -> np.ndarray:
    x = S.pcm(case["n"], case["channels"], case["bits"], case["rate"], case["cls"], case["seed"])
    assert sha(x.tobytes()) == case["pcm_sha256"], "synthetic generator drifted from the fixtures"
    return x


def case_stream(case: dict) -> bytes:
    with open(os.path.join(GOLDEN, case["name"] + ".flac"), "rb") as f:
        b = f.read()
    assert sha(b) == case["stream_sha256"]
    return b


def case_kwargs(case: dict) -> dict:
    return dict(block_size=case["block_size"], max_lpc_order=case["max_lpc_order"],
                qlp_precision=case["qlp_precision"], rice_min=case["rice_min"],
                rice_max=case["rice_max"])


def ood_pcm(case: dict) -> np.ndarray:
    return np.load(os.path.join(GOLDEN, case["name"] + ".npy")).astype(np.int32)


with open(os.path.join(GOLDEN, "decode_cases.json")) as _f:
    DECODE_CASES = {k: v for k, v in json.load(_f).items() if not k.startswith("_")}

STATUS_NAMES = {0: None, 1: "AssertionError", 2: "ValueError", 3: "ZeroDivisionError", 4: "EOFError",
                6: "OverflowError", 7: "KeyError"}


def decode_case(case: dict) -> tuple[bytes, np.ndarray]:
    """-> (stream, samples the reference decoder yields before it returns or raises)"""
    rows = case["samples"]
    want = np.array(rows, dtype=np.int64).reshape(len(rows), -1) if rows else np.zeros((0, 0), dtype=np.int64)
    return bytes.fromhex(case["stream"]), want
