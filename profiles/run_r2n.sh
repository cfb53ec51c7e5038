#!/bin/bash
# round-2 GPU session N: coalesced conv_dc epilogue, tile planner; plan sweep per layer
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv_kernels or spatial_forward or stream_golden" 2>&1 | tail -n 8 ) > gpurun_out/r2n_conv_test.log 2>&1
tail -n 4 gpurun_out/r2n_conv_test.log
timeout 600 python profiles/conv_bench.py plans 5 > gpurun_out/r2n_plans.jsonl 2> gpurun_out/r2n_plans.err
cat gpurun_out/r2n_plans.jsonl
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-gpu-eager --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench fps %.1f ms %.3f conv_ms %.3f'%(d['value'],d['ms_per_step'],d['roofline_tensor']['kernel_ms_per_step']))"
