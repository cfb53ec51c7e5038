#!/bin/bash
# round-2 GPU session ZZc (the last 90 GPU-seconds): packed fp16 conversions in the plane stores: tests of every writer + bench
mkdir -p gpurun_out
( timeout -s KILL 45 python -m pytest tests -m gpu -q -x -k "f16 or stream_golden or spatial_forward_golden or ccl_c256 or smooth_window_golden" 2>&1 | tail -n 4 ) > gpurun_out/r2zzc_pytest.log 2>&1
tail -n 2 gpurun_out/r2zzc_pytest.log
timeout -s KILL 40 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2zzc_bench.json 2> gpurun_out/r2zzc_bench.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2zzc_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('value %.1f ms %.3f e2e %.1f convms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_tensor']['kernel_ms_per_step']), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('FAILED', e)
PY
