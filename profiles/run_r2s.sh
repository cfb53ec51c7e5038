#!/bin/bash
# round-2 GPU session S: cost volume with a pixel row per thread (FMA-bound), conv_dc tile-plan parity test
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cost_volume or conv_dc_every_tile_plan or spatial_forward or temporal_golden" 2>&1 | tail -n 8 ) > gpurun_out/r2s_test.log 2>&1
tail -n 4 gpurun_out/r2s_test.log
if grep -q "passed" gpurun_out/r2s_test.log && ! grep -q "failed" gpurun_out/r2s_test.log; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager > gpurun_out/r2s_ncu_bench.log 2>&1
  python profiles/launch_summary.py gpurun_out/r2s_launches.csv > gpurun_out/r2s_launches_summary.txt 2>&1
  grep -i "cost_volume\|total" gpurun_out/r2s_launches_summary.txt
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-gpu-eager --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench fps %.1f ms %.3f'%(d['value'],d['ms_per_step']))"
fi
