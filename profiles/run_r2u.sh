#!/bin/bash
# round-2 GPU session U: SpatialNet's two mesh regressors on two streams
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or build_spatial or stream_host" 2>&1 | tail -n 6 ) > gpurun_out/r2u_test.log 2>&1
tail -n 3 gpurun_out/r2u_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
B="python bench.py --no-e2e --no-cpu-baseline --no-gpu-eager --steps 10 --warmup 3"
for v in 1 0; do
SS2_SIDE_STREAM=$v timeout 300 $B 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('side=$v bench fps %.1f ms %.3f'%(d['value'],d['ms_per_step']))"
done
