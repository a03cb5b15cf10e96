"""The C-ABI library loads here (no GPU) and exports every symbol include/flacb200.h declares;
without a device the compute entry points fail loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "flacb200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(flacb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from flac import _native as N
    L = N.lib()
    declared = _declared()
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(L, sym), sym
    assert sorted(N.EXPORTS) == declared
    assert L.flacb_version() == 0x000100
    assert b"AssertionError" in L.flacb_strerror(1) and b"ZeroDivisionError" in L.flacb_strerror(3)


def test_host_only_entry_points_work_without_a_gpu():
    from flac import _native as N
    assert N.stream_prologue(44100, 2, 16, 22050, 4608).hex() == (
        "664c6143" "80000022" "120012000000000000000ac442f000005622" + "00" * 16)
    assert N.encode_bound(4608 * 3, 2, 16, 4608, 5) >= 3 * 2 * 4608 * 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from flac import _native as N
    with pytest.raises(N.NativeError):
        N.Context(0)
    from flac.encoder import EncoderParameters, encode_pcm
    p = EncoderParameters(4608, range(6), range(13), 5)
    with pytest.raises(N.NativeError):
        b"".join(encode_pcm(44100, 16, 2, bytes(4608 * 4), p))


def test_product_never_imports_the_oracle_or_the_simulator():
    pkg = os.path.join(ROOT, "flac-py_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    t = f.read()
                if re.search(r"import oracle|from oracle|flac_oracle|simlib|libflacb_sim|refharness", t):
                    bad.append(fn)
    assert not bad, bad
