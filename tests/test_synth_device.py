"""The synthetic generator on the device (tools/synth/flacb200_synth.cu, benchmark tooling) writes bit for
bit what the numpy generator returns: the benchmark corpora of configs 3 and 4 only exist on the device."""
import hashlib

import numpy as np
import pytest

import flacb200_synth as S
from helpers import CASES

pytestmark = pytest.mark.gpu


def _device_pcm(n, ch, bits, rate, cls, seed, start):
    from flac import _native as N
    ctx = N.default_context()
    nb = n * ch * ((bits + 7) // 8)
    d = ctx.device_alloc(nb + 16)
    try:
        assert S.pcm_device(ctx.device, d, n, ch, bits, rate, cls, seed, start) == nb
        out = np.empty(nb, dtype=np.uint8)
        ctx.d2h(out.ctypes.data, d, nb)
    finally:
        ctx.device_free(d)
    return out.tobytes()


@pytest.mark.parametrize("bits,rate", [(16, 44100), (24, 96000), (24, 48000), (12, 32000)])
@pytest.mark.parametrize("cls", list(S.CLASSES) + ["mixed"])
def test_device_generator_is_bit_identical(cls, bits, rate):
    for ch, n, start, seed in ((2, 5000, 0, 0xF1AC0002), (8, 3000, 65536 * 4 - 1500, 0xF1AC0004), (1, 4097, 123456789, 7)):
        want = S.pack_le(S.pcm(n, ch, bits, rate, cls, seed, start), bits)
        assert _device_pcm(n, ch, bits, rate, cls, seed, start) == want, (cls, bits, ch, start)


def test_device_generator_across_every_class_of_a_mixed_stream():
    n, start = 65536 * 6 + 777, 65536 * 5 - 300  # every class once, both ends inside a segment
    want = S.pack_le(S.pcm(n, 2, 24, 96000, "mixed", 0xF1AC0003, start), 24)
    assert _device_pcm(n, 2, 24, 96000, "mixed", 0xF1AC0003, start) == want


@pytest.mark.parametrize("name", ["c4_mixed", "c4_white", "c3_clip", "c2_silence"])
def test_device_generator_reproduces_the_golden_inputs(name):
    case = CASES[name]
    raw = _device_pcm(case["n"], case["channels"], case["bits"], case["rate"], case["cls"], case["seed"], 0)
    x = S.unpack_le(raw, case["channels"], case["bits"])
    assert hashlib.sha256(x.astype(np.int32).tobytes()).hexdigest() == case["pcm_sha256"]
