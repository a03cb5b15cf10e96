#!/usr/bin/env python3
"""Randomised differential check of the DECODE path against the REAL reference decoder
(turlando/flac-py, flac/decoder.py), for the build container only (the reference is mounted at
/root/reference and cannot travel):

    python tests/golden/decode_differential_check.py [cases] [seed]

Small valid streams (made with the oracle encoder) get one to three random bit flips behind the
prologue, some are truncated; each is decoded by the unmodified reference decoder (lazily: the
samples it yields before it returns or raises) and by the CUDA decode kernels' logic on the CPU
SIMT simulator (tests/sim).  Compared: the samples and the exception class.  Excused: streams on
which the reference's unbounded integers leave int32, and streams whose (corrupted) frame headers
change the channel count from frame to frame (the C ABI returns one interleaved array with the
last frame's channel count).  The last run is recorded in DESIGN.md.
"""
import io
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'flac-py_b200'), os.path.join(ROOT, 'tools', 'synth')]
import numpy as np, oracle as O, refharness as R, flacb200_synth as S, simlib
STATUS={0:None,1:"AssertionError",2:"ValueError",3:"ZeroDivisionError",4:"EOFError",6:"OverflowError",7:"KeyError"}
rng=random.Random(int(sys.argv[2]) if len(sys.argv)>2 else 1)
cases=int(sys.argv[1]) if len(sys.argv)>1 else 50
bad=0; big=0; t0=time.time()
for it in range(cases):
    bs=rng.choice([64,192,576]); L=rng.choice([4,8,12]); q=rng.choice([5,8,12]); cls=rng.choice(["white","mixed","sine"])
    x=S.pcm(bs*3+rng.choice([0,17]),2,16,44100,cls,500+it)
    try:
        enc=O.encode_frames(x,16,bs,L,q,0,3)
    except O.OracleError:
        continue  # the clean input is outside the reference's domain: nothing to corrupt
    stream=bytearray(O.stream_prologue(44100,2,16,x.shape[0],bs)+enc.data)
    for _ in range(rng.choice([1,1,2,3])):
        pos=rng.randrange(42*8,len(stream)*8); stream[pos>>3]^=1<<(pos&7)
    if rng.random()<0.2: stream=stream[:rng.randrange(60,len(stream))]
    stream=bytes(stream)
    got=[]; rexc=None
    try:
        dec=R.load()["flac.decoder"]
        rate_,bits_,ch_,total_,rows=dec.decode(io.BufferedReader(io.BytesIO(stream)))
        for row in rows: got.append(list(row))
    except Exception as e:
        rexc=type(e).__name__
    rc,pcm,fch,nc=simlib.decode_frames(stream[42:],16)
    want=np.array(got,dtype=object).reshape(len(got),-1) if got else None
    too_big = any(abs(v)>=2**31 for row in got for v in row) or len({len(row) for row in got}) > 1
    same_exc = STATUS.get(rc,"?")==rexc
    same_n = pcm.shape[0]==len(got)
    same_v = same_n and (len(got)==0 or (pcm.shape[1]==len(got[0]) and (pcm.astype(object)==want).all()))
    if not (same_exc and same_n and same_v):
        if too_big: big+=1
        else:
            bad+=1; print("MISMATCH",it,dict(bs=bs,L=L,q=q,cls=cls),"ref",rexc,len(got),"sim rc",rc,STATUS.get(rc),pcm.shape)
print(f"{cases} corrupted streams, {bad} mismatches, {big} excused (reference values beyond int32 or frames of different channel counts), {time.time()-t0:.1f} s, python {sys.version.split()[0]}")
sys.exit(1 if bad else 0)
