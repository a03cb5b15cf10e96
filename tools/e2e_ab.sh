# development aid: end-to-end wall time per call under two settings of an environment knob, alternating, on one box
#   bash tools/e2e_ab.sh FLACB200_EDGE_CHUNK_FRAMES 444 148,296
k=$1; a=$2; b=$3
for i in 1 2 3; do for v in "$a" "$b"; do export $k="$v"; echo -n "$k=$v: "; timeout 120 python tools/e2e_timeline.py 2>&1 | grep "^call [1-4]" | awk '{printf "%s ", $4}'; echo; done; done
