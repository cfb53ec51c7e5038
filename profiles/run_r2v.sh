#!/bin/bash
# round-2 GPU session V: chunk-size sweep
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-gpu-eager --steps 10 --warmup 3"
for cfg in "16 32" "32 32" "8 32" "16 16" "32 64" "16 64"; do
set -- $cfg
SS2_SPATIAL_CHUNK=$1 SS2_TEMPORAL_CHUNK=$2 timeout 300 $B 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('spatial chunk $1 temporal chunk $2: fps %.1f ms %.3f'%(d['value'],d['ms_per_step']))"
done
