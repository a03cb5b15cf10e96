import sys, os, time, cProfile, pstats
ROOT=os.environ.get("GRAFT_REPO_ROOT","/root/repo")
sys.path[:0]=[os.path.join(ROOT,'flac-py_b200'), os.path.join(ROOT,'tools','synth'), ROOT]
import numpy as np
import bench
import flacb200_synth as S
from flac.encoder import EncoderParameters, encode_pcm
m=bench.workload_shape("c2", None)
ch,bits,bs,rate,n=m["channels"],m["bits"],m["block"],m["rate"],m["n"]
x=S.pcm(n,ch,bits,rate,"mixed",1234,start=0)
payload=x.astype("<i2").tobytes()
params=EncoderParameters(block_size=bs, rice_partition_order=range(m["rmin"], m["rmax"]+1), lpc_order=range(m["L"]+1), qlp_precision=m["q"])
for k in range(3):
    t0=time.perf_counter(); s=b"".join(encode_pcm(rate,bits,ch,payload,params,strict=False)); print("join ms",(time.perf_counter()-t0)*1e3,len(s))
pr=cProfile.Profile(); pr.enable()
s=b"".join(encode_pcm(rate,bits,ch,payload,params,strict=False))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
