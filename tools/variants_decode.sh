#!/bin/bash
# development aid: decode kernel time of a workload under several builds of the library (build/libflacb200_<name>.so)
w=$1; shift
for v in "$@"; do
  lib=flac-py_b200/csrc/libflacb200.so
  [ "$v" != default ] && lib=build/libflacb200_$v.so
  echo "== $w $v" | tee -a gpurun_out/variants.log
  FLACB200_LIB=$PWD/$lib timeout 90 python tools/perf_sweep.py --workload $w --steps 5 --decode 2>&1 | grep '"decode"' | cut -c1-400 | tee -a gpurun_out/variants.log
done
