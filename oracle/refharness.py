"""Drive the REAL reference (turlando/flac-py) when it is importable.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (to mint the golden
fixtures the oracle is pinned against) and by the optional live-reference tests.
The reference is looked up in $FLAC_REF, /root/reference, then baseline/_ref/; it
is imported in isolation so that it can coexist with the drop-in `flac` package of
this repo in one process.  Nothing here is available on the GPU box.
"""
from __future__ import annotations

import io
import os
import sys
from typing import Iterable

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("FLAC_REF", ""), "/root/reference",
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]

_ref = None


def reference_root() -> str | None:
    for p in _CANDIDATES:
        if p and os.path.isfile(os.path.join(p, "flac", "encoder.py")):
            return p
    return None


def available() -> bool:
    return reference_root() is not None


def load():
    """Import the reference's modules under private names; returns a dict of modules."""
    global _ref
    if _ref is not None:
        return _ref
    root = reference_root()
    if root is None:
        raise ImportError("reference flac-py not found (set FLAC_REF)")
    if sys.version_info < (3, 12):
        raise RuntimeError("the oracle is pinned to Python >= 3.12 (compensated float sum())")
    saved = {k: v for k, v in sys.modules.items() if k == "flac" or k.startswith("flac.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        import flac.encoder  # noqa: F401
        import flac.decoder  # noqa: F401
        import flac.crc  # noqa: F401
        mods = {k: v for k, v in sys.modules.items() if k == "flac" or k.startswith("flac.")}
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "flac" or k.startswith("flac.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _ref = mods
    return mods


def _params(block_size, max_lpc_order, qlp_precision, rice_min, rice_max):
    enc = load()["flac.encoder"]
    return enc.EncoderParameters(block_size=block_size,
                                 rice_partition_order=range(rice_min, rice_max + 1),
                                 lpc_order=range(max(max_lpc_order, 1) + 1),
                                 qlp_precision=qlp_precision)


def encode_chunks(samples: Iterable[list[int]], n: int, channels: int, sample_rate: int, bps: int,
                  block_size: int, max_lpc_order: int = 12, qlp_precision: int = 5,
                  rice_min: int = 0, rice_max: int = 5) -> list[bytes]:
    """Run the unmodified reference encode().  max_lpc_order == 0 installs the LPC stub of
    SURVEY.md 8(c): a residual so large that fixed always wins (encoder.py:138)."""
    mods = load()
    enc = mods["flac.encoder"]
    common = mods["flac.common"]
    params = _params(block_size, max_lpc_order, qlp_precision, rice_min, rice_max)
    orig = enc.encode_subframe_lpc
    if max_lpc_order == 0:
        enc.encode_subframe_lpc = lambda s, o, p: (
            common.SubframeHeader(common.SubframeTypeLPC(1), 0),
            common.SubframeLPC([], p, 0, [], [1 << 62]))
    try:
        return list(enc.encode(sample_rate, bps, channels, n, iter(samples), params))
    finally:
        enc.encode_subframe_lpc = orig


def encode_introspect(samples: list[list[int]], channels: int, sample_rate: int, bps: int,
                      block_size: int, max_lpc_order: int = 12, qlp_precision: int = 5,
                      rice_min: int = 0, rice_max: int = 5):
    """Reference encode plus what it decided per subframe (by wrapping its functions)."""
    mods = load()
    enc = mods["flac.encoder"]
    records: list[dict] = []
    cur: dict = {}
    o_fixed, o_lpc, o_res, o_auto, o_quant = (enc.encode_subframe_fixed, enc.encode_subframe_lpc,
                                              enc.encode_residual, enc.autocorrelation,
                                              enc.quantize_lpc_coefficients)

    def w_fixed(s):
        nonlocal cur
        cur = {"autoc": [], "quant": []}
        records.append(cur)
        h, sf = o_fixed(s)
        cur["fixed_order"] = sf.order
        if len(s) > 4:
            cur["fixed_err"] = [sum(abs(r) for r in enc.prediction_residual(s, c))
                                for c in mods["flac.common"].FIXED_PREDICTOR_COEFFICIENTS]
        cur["fixed_total"] = sum(abs(x) for x in sf.residual)
        return h, sf

    def w_lpc(s, o, p):
        h, sf = o_lpc(s, o, p)
        cur["lpc_order"] = h.type_.order
        cur["lpc_shift"] = sf.shift
        cur["lpc_coefs"] = list(sf.coefficients)
        cur["lpc_total"] = sum(abs(x) for x in sf.residual)
        return h, sf

    def w_auto(s, lag):
        r = o_auto(s, lag)
        cur["autoc"].append(r)
        return r

    def w_quant(c, p):
        r = o_quant(c, p)
        cur["quant"].append((list(r[0]), r[1]))
        return r

    def w_res(samples_, bs, ss, po, rng):
        r = o_res(samples_, bs, ss, po, rng)
        cur["porder"] = r.partition_order
        cur["params"] = [p.parameter for p in r.partitions]
        cur["method"] = r.coding_method.value
        cur["pred_order"] = po
        return r

    enc.encode_subframe_fixed, enc.encode_subframe_lpc = w_fixed, w_lpc
    enc.encode_residual, enc.autocorrelation, enc.quantize_lpc_coefficients = w_res, w_auto, w_quant
    try:
        chunks = encode_chunks(samples, len(samples), channels, sample_rate, bps, block_size,
                               max_lpc_order, qlp_precision, rice_min, rice_max)
    finally:
        enc.encode_subframe_fixed, enc.encode_subframe_lpc = o_fixed, o_lpc
        enc.encode_residual, enc.autocorrelation, enc.quantize_lpc_coefficients = \
            o_res, o_auto, o_quant
    for r in records:
        if "lpc_total" in r and max_lpc_order > 0:
            r["type"] = "lpc" if r["lpc_total"] < r["fixed_total"] else "fixed"
        else:
            r["type"] = "fixed"
    return chunks, records


def decode(stream: bytes):
    """-> (sample_rate, bps, channels, total, list of sample frames) via the reference decoder."""
    dec = load()["flac.decoder"]
    rate, bps, ch, total, it = dec.decode(io.BufferedReader(io.BytesIO(stream)))
    return rate, bps, ch, total, [list(x) for x in it]
