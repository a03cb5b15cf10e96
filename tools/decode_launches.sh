#!/bin/bash
# development aid: per-kernel times of one decode under a build of the library (ncu launch list; cold cache, serialised)
w=$1; v=$2
lib=flac-py_b200/csrc/libflacb200.so
[ "$v" != default ] && lib=build/libflacb200_$v.so
FLACB200_LIB=$PWD/$lib timeout 100 python tools/perf_sweep.py --workload $w --steps 1 --decode > /dev/null 2>&1 || exit 1
FLACB200_LIB=$PWD/$lib timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"find_sync|parse_headers|decode_frames|restore|chain|interleave" -c 7 --csv --log-file gpurun_out/dl_${w}_$v.csv python tools/perf_sweep.py --workload $w --steps 1 --decode > /dev/null 2>&1
grep -v "^==" gpurun_out/dl_${w}_$v.csv | awk -F'","' -v t="$w $v" 'NR>1{n=$5; sub(/\(.*/,"",n); printf "%s %s %s\n", t, n, $NF}' | tr -d '"'
