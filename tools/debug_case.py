#!/usr/bin/env python3
"""Development aid: encode one synthetic case on the GPU and report where it differs from the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "flac-py_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools", "synth"), ROOT]
import numpy as np  # noqa: E402
import flacb200_synth as S  # noqa: E402
import oracle as O  # noqa: E402
from flac.encoder import EncoderParameters, encode_frames  # noqa: E402

ch, bits, L, q, bs, cls = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6])
frames = int(sys.argv[7]) if len(sys.argv) > 7 else 40
x = S.pcm(bs * frames + bs // 3 + 5, ch, bits, 48000, cls, 99 + ch + bits)
p = EncoderParameters(block_size=bs, rice_partition_order=range(0, 5), lpc_order=range(L + 1), qlp_precision=q)
ref = O.encode_frames(x, bits, bs, L, q, 0, 4, threads=8)
for rep in range(3):
    sh = encode_frames(bits, ch, S.pack_le(x, bits), p, want_info=True)[0]
    got = sh.out.tobytes()
    o, ro = np.asarray(sh.offs), np.asarray(ref.frame_offsets)
    bad = [f for f in range(len(ro) - 1) if f + 1 >= len(o) or got[int(o[f]):int(o[f + 1])] != ref.data[int(ro[f]):int(ro[f + 1])]]
    print("rep", rep, "rc", sh.rc, "equal", got == ref.data, "len", len(got), len(ref.data), "bad frames", bad[:12], "of", len(ro) - 1)
    if bad and sh.info is not None:
        f = bad[0]
        for c in range(ch):
            i = sh.info[f * ch + c]
            print("  frame", f, "ch", c, "type", i.type, "order", i.order, "porder", i.porder, "bits", i.bits, "fixed_err", list(i.fixed_err), "params", list(i.params)[:i.nparams])
        a, b = got[int(o[f]):int(o[f + 1])], ref.data[int(ro[f]):int(ro[f + 1])]
        d = [k for k in range(min(len(a), len(b))) if a[k] != b[k]]
        print("  frame bytes", len(a), len(b), "first diff at", d[:5], "n diffs", len(d))
