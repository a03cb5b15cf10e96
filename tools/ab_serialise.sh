#!/bin/bash
# A/B of the two serialising kernels on the bench workloads (run on the GPU box)
mkdir -p gpurun_out
for w in "$@"; do
  for s in 1 2; do
    echo "== $w FLACB200_SERIALISE=$s" >> gpurun_out/ab.log
    FLACB200_SERIALISE=$s python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/ab.err | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    r = d['roofline']
    print(json.dumps({'value': d['value'], 'ms': d['ms_per_step'], 'stages': r['stages_ms'], 'frac': r['frac'], 'e2e': d['e2e']['value'], 'threads': d['config']['threads_per_frame'], 'smem': d['config']['smem_bytes_per_frame'], 'decode': d.get('decode')}))
" >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
