#!/bin/bash
# kernel time of the bench workloads under the serialising kernels (run on the GPU box):
#   tools/ab_serialise.sh "1 2 2,4" c2 c2fixed ...   (values of FLACB200_SERIALISE)
mkdir -p gpurun_out
sers=$1; shift
for w in "$@"; do
  for s in $sers; do
    echo "== $w FLACB200_SERIALISE=$s" | tee -a gpurun_out/ab.log
    FLACB200_SERIALISE=$s timeout 300 python tools/perf_sweep.py --workload $w --steps 5 2>&1 | tail -1 | cut -c1-300 | tee -a gpurun_out/ab.log
  done
done
